"""Profiling driver (test infrastructure): a few Generator forwards at a chosen size, for ncu.
    python tests/prof_step.py --B 16 --T 861 --iters 2 [--stft]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))
import torch  # noqa: E402
import b200voc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=16)
ap.add_argument("--T", type=int, default=861)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--stft", action="store_true")
ap.add_argument("--attn", action="store_true")
ap.add_argument("--critics", action="store_true", help="also one MultiScaleDiscriminator forward on 4 x 1 s")
a = ap.parse_args()
torch.manual_seed(1234)
if a.stft:
    x = torch.rand(1024, 88200, device="cuda") * 2 - 1
    for _ in range(a.iters):
        lm = b200voc.log_mel(x)
        s = b200voc.stft(x, 1024, 256)
        y = b200voc.istft(s, 1024, 256, 88200)
        m = b200voc.stft_magnitude(x, 1024, 256)
    torch.cuda.synchronize()
    print("stft ok", float(lm.mean()), float((y - x).abs().max()))
else:
    gen = b200voc.Generator(b200voc.GANConfig(use_attention=a.attn)).eval().cuda()
    g = torch.Generator().manual_seed(1)
    ins = [torch.randn(a.B, 80, a.T, generator=g).cuda(), torch.randn(a.B, a.T, 18, generator=g).cuda(),
           torch.randn(a.B, 128, generator=g).cuda(), torch.softmax(torch.randn(a.B, 6, generator=g), -1).cuda()]
    with torch.no_grad():
        w = gen(*ins)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            w = gen(*ins)
        e1.record()
    torch.cuda.synchronize()
    print("gen ok", tuple(w.shape), float(w.abs().max()), f"{e0.elapsed_time(e1) / a.iters:.3f} ms per forward", flush=True)
    if a.critics:
        msd = b200voc.MultiScaleDiscriminator(b200voc.GANConfig()).eval().cuda()
        outs, _ = msd(torch.rand(4, 1, 22050, device="cuda") * 2 - 1)
        torch.cuda.synchronize()
        print("msd ok", [tuple(o.shape) for o in outs])
