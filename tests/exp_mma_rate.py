"""Experiment driver (test infrastructure): raw tcgen05.mma SS-mode issue rate vs N, and the cost of
switching accumulator / shape between short groups of MMAs."""
import os, sys, json
os.environ.setdefault("B200VOC_LIB", "dev")   # experiment / trace exports live in libb200voc_dev.so
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
for variant, name in enumerate(["pingpong_try_wait_suspend", "pingpong_poll", "mma_commit_poll"]):
    cyc = run(30000 + variant, 2000)
    res[name] = dict(cycles_per_round_trip=round(cyc / 2000, 1))
print(json.dumps(res, indent=0))
