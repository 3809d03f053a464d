"""Debug driver (test infrastructure): clock64 timeline of CTA 0 of the narrow-stage residual block."""
import os, sys
os.environ.setdefault("B200VOC_LIB", "dev")   # experiment / trace exports live in libb200voc_dev.so
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))
import torch
from b200voc import _lib
lib = _lib.load()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N, T = 64, 861
P = 128 if C == 64 else 256
L = T * P
dt = torch.float16
a = (torch.randn(N, L, C, device="cuda") * 0.3).to(dt)   # raw x (narrow-stage convention)
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
    _lib.check(lib.b200voc_resblock(a.data_ptr(), wpk.data_ptr(), bc.data_ptr(), bp.data_ptr(), film.data_ptr(), N, L, C, 3, T, 4, 0, 0, out.data_ptr(), st))
    torch.cuda.synchronize()
lib.b200voc_debug_set_trace(0)
t = trace.cpu().view(7, 64, 4)
t0 = int(t[0, 0, 0])
print("C", C, "cycles relative to first TMA issue; rows = tile index")
print("tile | TMA issue | X: landed d2_empty | prep: x_done done | G1: p_full d1_empty issued | E1: d1_full h_empty done | G2: h_full | E2: d2_full done")
for i in range(20, 44):
    r = lambda s, k: int(t[s, i, k]) - t0 if int(t[s, i, k]) else -1
    print(f"{i:3d} | {r(0,0):7d} | {r(5,0):7d} {r(5,1):7d} | {r(6,0):7d} {r(6,1):7d} | {r(1,0):7d} {r(1,1):7d} {r(1,2):7d} | {r(3,0):7d} {r(3,1):7d} {r(3,2):7d} | {r(2,0):7d} | {r(4,0):7d} {r(4,1):7d}")
per_tile = (int(t[4, 60, 1]) - int(t[4, 20, 1])) / 40
print("steady-state cycles per tile:", per_tile)
print("E1 detail: tile | d1_full->h_empty | ->first LDTM done | ->first 16 ch stored | ->all stored | ->arrived")
for i in range(30, 38):
    a, b, c, d, e, f = int(t[3, i, 0]), int(t[3, i, 1]), int(t[3, i, 3]), int(t[6, i, 2]), int(t[6, i, 3]), int(t[3, i, 2])
    print(i, b - a, c - b, d - c, e - d, f - e)
print("G1 issued -> issuer sees d1_full (dbg&4):", [int(t[1, i, 3]) - int(t[1, i, 2]) for i in range(30, 40)])
print("G1 d1_empty pass -> issued:", [int(t[1, i, 2]) - int(t[1, i, 1]) for i in range(30, 40)])
