"""Experiment driver (test infrastructure): raw tcgen05.mma SS-mode issue rate vs N."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))
import torch
from b200voc import _lib
lib = _lib.load()
res = {}
for blocks in (1, 148):
    for n in (64, 128, 256):
        out = torch.zeros(blocks, dtype=torch.int64, device="cuda")
        iters = 2000
        for _ in range(2):
            _lib.check(lib.b200voc_exp_mma_rate(n, iters, blocks, out.data_ptr(), _lib.current_stream()))
        torch.cuda.synchronize()
        cyc = float(out.float().mean())
        per_mma = cyc / (iters * 4)
        ideal = 128 * n * 16 / 4096.0          # 8192 flop/clk/SM -> 4096 MAC/clk
        res[f"blocks{blocks}_N{n}"] = dict(cycles_per_mma=round(per_mma, 1), ideal=ideal, frac_of_peak=round(ideal / per_mma, 3))
print(json.dumps(res, indent=1))
