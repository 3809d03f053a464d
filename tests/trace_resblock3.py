"""Debug driver (test infrastructure, -DB200VOC_TRACE builds): clock64 timeline of CTA 0 of the wide-stage
residual block (C=128): per 64-channel chunk, when the MMA issuer passed its waits / finished issuing, when
the GLU epilogue saw the accumulator and finished, when GEMM2's k-block was issued; per tile, the store epilogue."""
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
NCH = C // 64
dt = torch.float16
a = (torch.randn(N, L, C, device="cuda") * 0.3).to(dt)
film = torch.randn(N // 4, T, 2 * C, device="cuda")
wc, wp = torch.randn(2 * C, C, 3, device="cuda") * 0.05, torch.randn(C, C, 1, device="cuda") * 0.05
bc, bp = torch.zeros(2 * C, device="cuda"), torch.zeros(C, device="cuda")
wpk = torch.empty(lib.b200voc_resblock_packed_elems(C), dtype=dt, device="cuda")
st = _lib.current_stream()
_lib.check(lib.b200voc_pack_resblock_weights(wc.data_ptr(), wp.data_ptr(), C, 0, wpk.data_ptr(), st))
out = torch.empty_like(a)
trace = torch.zeros(7 * 64 * 4, dtype=torch.int64, device="cuda")
for it in range(2):
    if it == 1:
        lib.b200voc_debug_set_trace(trace.data_ptr())
    _lib.check(lib.b200voc_resblock(a.data_ptr(), wpk.data_ptr(), bc.data_ptr(), bp.data_ptr(), film.data_ptr(), N, L, C, 3, T, 4, 0, 1, out.data_ptr(), st))
    torch.cuda.synchronize()
lib.b200voc_debug_set_trace(0)
t = trace.cpu().view(7, 64, 4)
t0 = int(t[1, 0, 0])
r = lambda s, i, k: (int(t[s, i, k]) - t0) if int(t[s, i, k]) else -1
print(f"C={C}: cycles relative to the first chunk; chunk gc = tile*{NCH} + j")
print(" gc | MMA: a_full d1_empty | issued(+d1_full commit) | after g2(prev) || E1: d1_full h_empty done || G2 of this chunk: h_full d2_empty")
for gc in range(16, 40):
    print(f"{gc:3d} | {r(1,gc,0):7d} {r(1,gc,1):7d} | {r(1,gc,2):7d} | {r(1,gc,3):7d} || {r(3,gc,0):7d} {r(3,gc,1):7d} {r(3,gc,2):7d} || {r(2,gc,0):7d} {r(2,gc,1):7d}")
print("tile | A load issued | E2: d2_full done")
for it in range(8, 20):
    print(f"{it:3d} | {r(0,it,0):7d} | {r(4,it,0):7d} {r(4,it,1):7d}")
per_chunk = (int(t[1, 60, 2]) - int(t[1, 20, 2])) / 40
print("steady-state cycles per chunk:", per_chunk, " (MMA floor 1536 + 256 for GEMM2's k-block)")
