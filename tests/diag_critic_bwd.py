"""Development script (GPU): per-tensor relative error of the critic backward against the fp64 CPU oracle, for the
abs-feature loss and for a smooth loss, with the tensor-core paths on and off.  Prints a table; no assertions."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tts-core-remastered-1_b200")]
from oracle import vocoder7_oracle as O  # noqa: E402
import b200voc  # noqa: E402


def loss_abs(outs, feats):
    loss = 0.0
    for o in outs:
        loss = loss + ((o - 1.0) ** 2).mean()
    for fs in feats:
        for j, f in enumerate(fs):
            loss = loss + (0.5 + 0.1 * j) * f.abs().mean()
    return loss


def loss_smooth(outs, feats):
    loss = 0.0
    for o in outs:
        loss = loss + ((o - 1.0) ** 2).mean()
    for fs in feats:
        for j, f in enumerate(fs):
            loss = loss + (0.5 + 0.1 * j) * (f ** 2).mean() / (1.0 + float((f.detach() ** 2).mean()))
    return loss


def run(kind, training, loss_fn, label):
    cls = {"mpd": b200voc.MultiPeriodDiscriminator, "msd": b200voc.MultiScaleDiscriminator,
           "mbd": b200voc.MultiBandDiscriminator}[kind]
    torch.manual_seed(1234)
    mod = cls(b200voc.GANConfig()).cuda()
    if os.environ.get("DIAG_WARM", "0") == "1":      # let the power iteration converge (sigma -> spectral norm)
        mod.train()
        with torch.no_grad():
            for _ in range(6):
                mod(torch.rand(1, 1, 600, device="cuda"))
    mod.train(training)
    sd = {k: v.detach().cpu().double() for k, v in mod.state_dict().items()}
    x = torch.rand(2, 1, 2403, generator=torch.Generator().manual_seed(8)) * 2 - 1
    xg = x.cuda().requires_grad_(True)
    outs, feats = mod(xg)
    loss_fn(outs, feats).backward()
    torch.cuda.synchronize()
    for k in sd:
        if k.endswith("weight_orig") or k.endswith("bias"):
            sd[k].requires_grad_(True)
    xr = x.double().requires_grad_(True)
    r_outs, r_feats = O.critic_forward(kind, sd, O.OracleConfig(), xr, training=training)
    loss_fn(r_outs, r_feats).backward()
    fwd = max(float((a.detach().cpu().double() - b.detach()).abs().max() / b.detach().abs().max())
              for fa, fb in zip(feats, r_feats) for a, b in zip(fa, fb))
    rows = [("x", float((xg.grad.cpu().double() - xr.grad).abs().max() / xr.grad.abs().max()))]
    for n, p in mod.named_parameters():
        r = sd[n].grad
        rows.append((n.replace("discriminators.", ""), float((p.grad.cpu().double() - r).abs().max() / r.abs().max())))
    worst = sorted(rows, key=lambda t: -t[1])[:4]
    print(f"{label:28s} {kind} train={int(training)} fwd-feat {fwd:.2e}  worst: " + "  ".join(f"{n} {e:.2e}" for n, e in worst), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "tc"
    for kind in ("mpd", "msd", "mbd"):
        for training in (False, True):
            run(kind, training, loss_abs, f"[{mode}] abs-feature loss")
            run(kind, training, loss_smooth, f"[{mode}] smooth loss")
