"""Experiment driver (test infrastructure): a CTA pair issuing tcgen05.mma.cta_group::2 (M=256, N=128)."""
import os, sys
os.environ.setdefault("B200VOC_LIB", "dev")   # experiment / trace exports live in libb200voc_dev.so
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))
import torch
from b200voc import _lib
lib = _lib.load()
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
A = torch.randn(256 * pairs, 64, device="cuda").half()
B = torch.randn(128, 64, device="cuda").half()
out = torch.full((256 * pairs, 128), float("nan"), device="cuda")
cyc = torch.zeros(pairs, dtype=torch.int64, device="cuda")
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0    # 1: both CTAs' TMA loads signal the leader's mbarrier
_lib.check(lib.b200voc_exp_cta2(A.data_ptr(), B.data_ptr(), pairs + 1000 * variant, out.data_ptr(), cyc.data_ptr(), _lib.current_stream()))
torch.cuda.synchronize()
ref = A.float() @ B.float().t()
err = (out - ref).abs().max().item()
print("cta_group::2 MMA: max abs err", err, "cycles (issue->done)", cyc.tolist())
# which half of B does each CTA supply?  report the error per (row half, column half)
for rh in range(2):
    for ch in range(2):
        e = (out[rh * 128:(rh + 1) * 128, ch * 64:(ch + 1) * 64] - ref[rh * 128:(rh + 1) * 128, ch * 64:(ch + 1) * 64]).abs().max().item()
        print(f"  rows {rh*128}-{rh*128+127} cols {ch*64}-{ch*64+63}: err {e:.3e}")
