"""A/B driver (test infrastructure): times the Generator step (B=16 x T=861, BASELINE configs[1]) and its per-layer
CUDA-event profile with two builds of the library on the SAME box, alternating, in fresh processes.
    python tests/ab_step.py <libA.so> <libB.so> [<libC.so> ...] [rounds]          (child: python tests/ab_step.py --child)
Box-to-box spread under the power cap is ~5 %, larger than most kernel changes: only same-box pairs are comparable."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))
    import torch
    import b200voc
    from b200voc import _lib
    lib = _lib.load()
    torch.manual_seed(1234)
    B, T = int(os.environ.get("AB_B", 16)), int(os.environ.get("AB_T", 861))
    gen = b200voc.Generator(b200voc.GANConfig(use_attention=False)).eval().cuda()
    g = torch.Generator().manual_seed(1)
    ins = [torch.randn(B, 80, T, generator=g).cuda(), torch.randn(B, T, 18, generator=g).cuda(),
           torch.randn(B, 128, generator=g).cuda(), torch.softmax(torch.randn(B, 6, generator=g), -1).cuda()]
    out = torch.empty(B, 1, 256 * T, device="cuda")
    with torch.no_grad():
        for _ in range(5):
            gen(*ins, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 30
        e0.record()
        for _ in range(steps):
            gen(*ins, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        lib.b200voc_gen_profile_enable(gen._handle, 1)
        for _ in range(3):
            gen(*ins, out=out)
        torch.cuda.synchronize()
        layers = {}
        for i in range(lib.b200voc_gen_profile_count(gen._handle)):
            layers[lib.b200voc_gen_profile_name(gen._handle, i).decode()] = round(float(lib.b200voc_gen_profile_ms(gen._handle, i)), 4)
    print(json.dumps(dict(ms=ms, layers=layers, checksum=float(out.double().abs().sum()))))


def main():
    libs = [a for a in sys.argv[1:] if a.endswith(".so")]          # two or more builds
    rest = [a for a in sys.argv[1:] if not a.endswith(".so")]
    rounds = int(rest[0]) if rest else 3
    res = {l: [] for l in libs}
    for r in range(rounds):
        for l in libs:
            env = dict(os.environ, B200VOC_LIB=os.path.abspath(l))
            o = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
            if o.returncode != 0:
                print(l, "FAILED", o.stderr[-2000:])
                continue
            res[l].append(json.loads(o.stdout.strip().splitlines()[-1]))
    for l in libs:
        rs = res[l]
        if not rs:
            continue
        print(f"== {l}: ms/step " + " ".join(f"{x['ms']:.3f}" for x in rs) + f"   checksum {rs[0]['checksum']:.6e}")
        keys = rs[0]["layers"].keys()
        print("   " + " ".join(f"{k}={min(x['layers'][k] for x in rs):.3f}" for k in keys))


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
