"""Development script (GPU): the critic half of a training step -- .train() forward + backward of a loss over every score
and feature map -- for ncu launch lists (`ncu --metrics gpu__time_duration.sum ... python tests/prof_critics.py msd`) and
CUDA-event timing."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tts-core-remastered-1_b200")]
import b200voc  # noqa: E402


def main():
    kinds = sys.argv[1:] or ["mpd", "msd", "mbd"]
    B, T = 4, 22050
    x = (torch.rand(B, 1, T, device="cuda") * 2 - 1).requires_grad_(True)
    for kind in kinds:
        cls = {"mpd": b200voc.MultiPeriodDiscriminator, "msd": b200voc.MultiScaleDiscriminator,
               "mbd": b200voc.MultiBandDiscriminator}[kind]
        torch.manual_seed(1234)
        crit = cls(b200voc.GANConfig()).cuda().train()

        def step():
            crit.zero_grad(set_to_none=True)
            o, f = crit(x)
            (sum((s ** 2).mean() for s in o) + sum(m.abs().mean() for fs in f for m in fs)).backward()
        iters = int(os.environ.get("PROF_ITERS", "3"))
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            step()
        b.record()
        torch.cuda.synchronize()
        print(f"{kind}: train fwd+bwd {a.elapsed_time(b) / iters:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
