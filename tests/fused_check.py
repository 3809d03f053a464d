"""Bring-up / A-B driver for the fused narrow-stage kernels (csrc/stage_fused.cu); run on a B200:
    python tests/fused_check.py [B T]
Compares taps res2.2 / res3.2 and the waveform of the fused path with the CPU oracle, prints where the
largest error sits (row within the 512/256-row strip), then times fused vs layer-by-layer."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tts-core-remastered-1_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from oracle import vocoder7_oracle as O  # noqa: E402


def main():
    from b200voc import GANConfig, Generator
    B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2, 20)
    ocfg = O.OracleConfig(use_attention=False)
    ora = O.make_generator(ocfg, seed=1234)
    gen = Generator(GANConfig(use_attention=False)).eval()
    gen.load_state_dict(ora.state_dict())
    gen = gen.cuda()
    mel, pros, sty, emo = O.synthetic_inputs(B, T, seed=8)
    taps = {}
    with torch.no_grad():
        ref = O.generator_forward(ora.state_dict(), ocfg, mel, pros, sty, emo, taps=taps)
        args = [x.cuda() for x in (mel, pros, sty, emo)]
        for k in ["res2.2", "res3.2"]:
            _, t = gen(*args, _tap=k)
            torch.cuda.synchronize()
            want = torch.stack(taps[k], 1)
            want = want.reshape(-1, want.shape[2], want.shape[3])
            got = t.view(want.shape).cpu()
            err = (got - want).abs()
            n, c, l = [int(v) for v in torch.nonzero(err == err.max())[0]]
            print(f"tap {k}: max-abs {float(err.max()):.3e} (ref max {float(want.abs().max()):.2f}) at seq {n} ch {c} l {l}; "
                  f"finite {bool(torch.isfinite(got).all())}; launches {gen.launch_count()}", flush=True)
            rowerr = err.amax(dim=(0, 1))
            bad = torch.nonzero(rowerr > 4e-3 * max(1.0, float(want.abs().max()))).flatten()
            if bad.numel():
                print(f"   rows over tolerance: {bad.numel()} of {rowerr.numel()}, first {bad[:12].tolist()} last {bad[-6:].tolist()}")
        wav = gen(*args).cpu()
        torch.cuda.synchronize()
        err = (wav - ref).abs()
        print(f"wav: max-abs {float(err.max()):.3e} SNR {O.snr_db(ref, wav):.1f} dB at l {int(err.flatten().argmax()) % wav.shape[-1]}; "
              f"launches {gen.launch_count()}", flush=True)
        rowerr = err.amax(dim=(0, 1))
        bad = torch.nonzero(rowerr > 1e-3).flatten()
        if bad.numel():
            print(f"   samples over 1e-3: {bad.numel()} of {rowerr.numel()}, first {bad[:12].tolist()} last {bad[-6:].tolist()}")
        # timing
        Bt, Tt = 16, 861
        big = [x.cuda() for x in O.synthetic_inputs(Bt, Tt, seed=3)]
        for _ in range(3):
            gen(*big)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gen(*big)
        e1.record()
        torch.cuda.synchronize()
        print(f"B=16 T=861: {e0.elapsed_time(e1) / 10:.3f} ms/step (B200VOC_FUSED={os.environ.get('B200VOC_FUSED', '1')})", flush=True)
        lib = __import__("b200voc")._lib.load()
        lib.b200voc_gen_profile_enable(gen._handle, 1)
        gen(*big)
        torch.cuda.synchronize()
        n = lib.b200voc_gen_profile_count(gen._handle)
        print("  ".join(f"{lib.b200voc_gen_profile_name(gen._handle, i).decode()}={lib.b200voc_gen_profile_ms(gen._handle, i):.3f}"
                        for i in range(n)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ab":
        for v in ("1", "0"):
            env = dict(os.environ, B200VOC_FUSED=v)
            subprocess.run([sys.executable, __file__] + sys.argv[2:], env=env, timeout=300)
    else:
        main()
