"""GPU bring-up probe (test infrastructure).  Each case runs in its own subprocess so that a
faulting kernel (trap / illegal address poisons the CUDA context) cannot hide the other results.

    python tests/gpu_probe.py            # run every case, write gpurun_out/probe.jsonl
    python tests/gpu_probe.py <case>     # run one case in-process
"""
from __future__ import annotations

import json
import os
os.environ.setdefault("B200VOC_LIB", "dev")   # experiment / trace exports live in libb200voc_dev.so
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))

CASES = ["rowshift", "convt_up3", "convt_up2", "convt_up1", "convt_up0", "res32", "res64", "res128", "res256",
         "gen_noattn", "gen_taps"]


def _imports():
    import numpy as np
    import torch
    from b200voc import _lib
    return np, torch, _lib


def case_rowshift():
    np, torch, _lib = _imports()
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    a = torch.randn(144, 64, generator=g).half().cuda()
    b = torch.randn(64, 64, generator=g).half().cuda()
    out = torch.zeros(2, 16, 128, 64, device="cuda")
    _lib.check(lib.b200voc_exp_rowshift(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), _lib.current_stream()))
    torch.cuda.synchronize()
    res = {}
    for v in range(2):
        errs = []
        for s in range(16):
            ref = a[s:s + 128].float() @ b.float().t()
            errs.append(float((out[v, s] - ref).abs().max()))
        res[f"variant{v}_maxerr_by_shift"] = [round(e, 4) for e in errs]
    return res


def _convt(Cin, Cout, s, fmt=0, N=3, Lin=150):
    np, torch, _lib = _imports()
    lib = _lib.load()
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(N, Cin, Lin, generator=g) * 0.5).to(dt)
    w = (torch.randn(Cin, Cout, 2 * s, generator=g) / (2 * Cin) ** 0.5)
    bias = torch.randn(Cout, generator=g) * 0.1
    ref = torch.nn.functional.conv_transpose1d(x.double(), w.to(dt).double(), bias.double(), stride=s, padding=s // 2)
    x_cl = x.transpose(1, 2).contiguous().cuda()          # [N, L, C]
    wp = torch.empty(lib.b200voc_convt_packed_elems(Cin, Cout, s), dtype=dt, device="cuda")
    wd, bd = w.cuda(), bias.cuda()
    st = _lib.current_stream()
    _lib.check(lib.b200voc_pack_convt_weight(_lib.ptr(wd), Cin, Cout, s, fmt, _lib.ptr(wp), st))
    out = torch.full((N, s * Lin, Cout), float("nan"), dtype=dt, device="cuda")
    res = {}
    for lre in (0, 1):
        _lib.check(lib.b200voc_convt1d(_lib.ptr(x_cl), _lib.ptr(wp), _lib.ptr(bd), N, Lin, Cin, Cout, s, fmt, lre,
                                       _lib.ptr(out), st))
        torch.cuda.synchronize()
        r = torch.nn.functional.leaky_relu(ref, 0.1) if lre else ref
        got = out.float().cpu().transpose(1, 2).double()
        err = (got - r).abs()
        res[f"lrelu{lre}_maxerr"] = float(err.max())
        res[f"lrelu{lre}_nan"] = int(torch.isnan(got).sum())
        res[f"lrelu{lre}_refmax"] = float(r.abs().max())
    return res


def _res(C, fmt=0, B=2, nb=4, T=6, P=55):
    np, torch, _lib = _imports()
    from oracle import vocoder7_oracle as O
    lib = _lib.load()
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    N, L = B * nb, T * P
    out_res = {}
    for d in (1, 3, 5):
        g = torch.Generator().manual_seed(10 + d)
        x = torch.randn(N, C, L, generator=g) * 0.5
        a = torch.nn.functional.leaky_relu(x, 0.1).to(dt)                 # stored form
        xr = torch.where(a.float() >= 0, a.float(), a.float() * 10.0)      # what the kernel recovers
        cond = torch.randn(B, 128, T, generator=g)
        wc = torch.randn(2 * C, C, 3, generator=g) / (3 * C) ** 0.5
        bc = torch.randn(2 * C, generator=g) * 0.1
        wf = torch.randn(2 * C, 128, 1, generator=g) / 128 ** 0.5
        bf = torch.randn(2 * C, generator=g) * 0.1
        wp_ = torch.randn(C, C, 1, generator=g) / C ** 0.5
        bp = torch.randn(C, generator=g) * 0.1
        # reference on the kernel's operands (weights rounded to dt), per band -> batch B
        refs = []
        xq = xr.double().view(B, nb, C, L)
        for band in range(nb):
            # residual_block_forward applies leaky_relu itself: lrelu(xr) == a exactly
            refs.append(O.residual_block_forward(xq[:, band], cond.double(), wc.to(dt).double(), bc.double(), wf.double(),
                                                 bf.double(), wp_.to(dt).double(), bp.double(), d))
        ref = torch.stack(refs, 1).reshape(N, C, L)
        film = torch.nn.functional.conv1d(cond, wf, bf)                    # [B, 2C, T]
        film[:, :C] += 1.0
        film_cl = film.transpose(1, 2).contiguous().cuda()                 # [B, T, 2C]
        a_cl = a.transpose(1, 2).contiguous().cuda()
        wpk = torch.empty(lib.b200voc_resblock_packed_elems(C), dtype=dt, device="cuda")
        st = _lib.current_stream()
        wcd, wpd, bcd, bpd = wc.cuda(), wp_.cuda(), bc.cuda(), bp.cuda()
        _lib.check(lib.b200voc_pack_resblock_weights(_lib.ptr(wcd), _lib.ptr(wpd), C, fmt, _lib.ptr(wpk), st))
        out = torch.full((N, L, C), float("nan"), dtype=dt, device="cuda")
        _lib.check(lib.b200voc_resblock(_lib.ptr(a_cl), _lib.ptr(wpk), _lib.ptr(bcd), _lib.ptr(bpd), _lib.ptr(film_cl),
                                        N, L, C, d, T, nb, fmt, 0, _lib.ptr(out), st))
        torch.cuda.synchronize()
        got = out.float().cpu().transpose(1, 2).double()
        out_res[f"d{d}_maxerr"] = float((got - ref).abs().max())
        out_res[f"d{d}_nan"] = int(torch.isnan(got).sum())
        out_res[f"d{d}_refmax"] = float(ref.abs().max())
    return out_res


def _gen(name, taps=False):
    np, torch, _lib = _imports()
    from oracle import vocoder7_oracle as O
    from b200voc import GANConfig, Generator
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    cfg = GANConfig(use_attention=False)
    torch.manual_seed(1234)
    gen = Generator(cfg).eval()
    ora = O.make_generator(O.OracleConfig(use_attention=False))
    gen.load_state_dict(ora.state_dict())
    gen = gen.cuda()
    mel, pros, sty, emo = (torch.from_numpy(gold[k]) for k in ("mel", "prosody", "style", "emotion"))
    res = {}
    with torch.no_grad():
        wav = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda())
        torch.cuda.synchronize()
        ref = torch.from_numpy(gold["wav"])
        res["wav_maxerr"] = float((wav.cpu() - ref).abs().max())
        res["wav_snr_db"] = O.snr_db(ref, wav.cpu())
        res["launches"] = gen.launch_count()
        if taps:
            otaps = {}
            O.generator_forward(ora.state_dict(), ora.cfg, mel, pros, sty, emo, taps=otaps)
            B, T = mel.shape[0], mel.shape[2]
            for k in ["cond", "split", "up0", "res0.0", "res0.1", "res0.2", "up1", "res1.2", "up2", "res2.2", "up3",
                      "res3.0", "res3.2"]:
                _, t = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda(), _tap=k)
                torch.cuda.synchronize()
                if k == "cond":
                    got = t.view(B, T, 128).transpose(1, 2).cpu()
                    want = otaps["cond"]
                else:
                    want = torch.stack(otaps[k], 1)            # [B, nb, C, L]
                    want = want.reshape(-1, want.shape[2], want.shape[3])
                    got = t.view(want.shape).cpu()
                res[f"tap_{k}_maxerr"] = float((got - want).abs().max())
                res[f"tap_{k}_refmax"] = float(want.abs().max())
    return res


def run_case(name):
    if name == "rowshift":
        return case_rowshift()
    if name.startswith("convt_up"):
        shapes = {"convt_up0": (512, 256, 8), "convt_up1": (256, 128, 8), "convt_up2": (128, 64, 2),
                  "convt_up3": (64, 32, 2)}
        return _convt(*shapes[name])
    if name.startswith("res"):
        return _res(int(name[3:]))
    if name == "gen_noattn":
        return _gen("gen_b2_t9_noattn")
    if name == "gen_taps":
        return _gen("gen_b2_t9_noattn", taps=True)
    raise SystemExit(f"unknown case {name}")


def main():
    if len(sys.argv) > 1 and sys.argv[1] != "--all":
        print(json.dumps({"case": sys.argv[1], **run_case(sys.argv[1])}))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out_path = os.path.join(ROOT, "gpurun_out", "probe.jsonl")
    with open(out_path, "w") as f:
        for c in CASES:
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), c], capture_output=True, text=True,
                                   timeout=240)
                line = {"case": c, "rc": p.returncode, "secs": round(time.time() - t0, 1),
                        "stdout": p.stdout.strip().splitlines()[-1:] if p.stdout.strip() else [],
                        "stderr_tail": p.stderr.strip().splitlines()[-6:]}
            except subprocess.TimeoutExpired:
                line = {"case": c, "rc": "timeout"}
            f.write(json.dumps(line) + "\n")
            f.flush()
            print(json.dumps(line))


if __name__ == "__main__":
    main()
