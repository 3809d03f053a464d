"""Experiment driver (test infrastructure): raw tcgen05.mma SS-mode issue rate vs N, and the cost of
switching accumulator / shape between short groups of MMAs."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))
import torch
from b200voc import _lib
lib = _lib.load()
res = {}
blocks = 148
def run(n, iters):
    out = torch.zeros(blocks, dtype=torch.int64, device="cuda")
    for _ in range(2):
        _lib.check(lib.b200voc_exp_mma_rate(n, iters, blocks, out.data_ptr(), _lib.current_stream()))
    torch.cuda.synchronize()
    return float(out.float().mean())
for shape, name in enumerate(["32x32b.x32", "16x256b.x8", "16x128b.x16", "16x64b.x32"]):
    for nw in (1, 4, 16):
        cyc = run(20000 + 100 * shape + nw, 2000)
        res[f"tmem_ld_{name}_{nw}warps"] = dict(cycles_per_ld=round(cyc / 2000, 1), bytes_per_clk_per_sm=round(nw * 4096 * 2000 / cyc, 1))
print(json.dumps(res, indent=0))
