"""Debug driver (test infrastructure): kernel time of the wide-stage residual block with parts of the
pipeline knocked out (B200VOC_DBG bits, -DB200VOC_TRACE builds only): 1 = GLU epilogue does nothing,
2 = store epilogue does nothing (C=128), 4 = weight ring not refilled, 8 = no GEMM2 MMAs, 16 = input
tiles not refilled, 32 = no GEMM1 MMAs."""
import os, sys
os.environ.setdefault("B200VOC_LIB", "dev")   # experiment / trace exports live in libb200voc_dev.so
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))
import torch
from b200voc import _lib
lib = _lib.load()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N, T = 64, 861
P = 64 if C == 128 else 8
L = T * P
dt = torch.float16
a = (torch.randn(N, L, C, device="cuda") * 0.3).to(dt)
film = torch.randn(N // 4, T, 2 * C, device="cuda")
wc, wp = torch.randn(2 * C, C, 3, device="cuda") * 0.05, torch.randn(C, C, 1, device="cuda") * 0.05
bc, bp = torch.zeros(2 * C, device="cuda"), torch.zeros(C, device="cuda")
wpk = torch.empty(lib.b200voc_resblock_packed_elems(C), dtype=dt, device="cuda")
st = _lib.current_stream()
_lib.check(lib.b200voc_pack_resblock_weights(wc.data_ptr(), wp.data_ptr(), C, 0, wpk.data_ptr(), st))
out = torch.empty_like(a)
run = lambda: _lib.check(lib.b200voc_resblock(a.data_ptr(), wpk.data_ptr(), bc.data_ptr(), bp.data_ptr(), film.data_ptr(), N, L, C, 3, T, 4, 0, 1, out.data_ptr(), st))
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
flops = 14.0 * N * L * C * C
print(f"C={C} dbg={os.environ.get('B200VOC_DBG', '0')}: {ms:.4f} ms  {flops / ms / 1e9:.0f} TFLOP/s")
