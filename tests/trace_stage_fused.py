"""Debug driver (test infrastructure, dev build): clock64 timeline of CTA 0 of a fused narrow-stage kernel
(csrc/stage_fused.cu) inside a full B=16 x T=861 step.   python tests/trace_stage_fused.py [32|64a|64b]"""
import os, sys
os.environ.setdefault("B200VOC_LIB", "dev")   # experiment / trace exports live in libb200voc_dev.so
which = sys.argv[1] if len(sys.argv) > 1 else "32"
os.environ["B200VOC_SF_TRACE"] = which
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))
import torch
from b200voc import _lib, GANConfig, Generator
from oracle import vocoder7_oracle as O
lib = _lib.load()
ocfg = O.OracleConfig(use_attention=False)
ora = O.make_generator(ocfg, seed=1234)
gen = Generator(GANConfig(use_attention=False)).eval()
gen.load_state_dict(ora.state_dict())
gen = gen.cuda()
args = [x.cuda() for x in O.synthetic_inputs(16, 861, seed=3)]
trace = torch.zeros(16 * 5 * 32, dtype=torch.int64, device="cuda")
with torch.no_grad():
    gen(*args)
    torch.cuda.synchronize()
    lib.b200voc_debug_set_trace(trace.data_ptr())
    gen(*args)
    torch.cuda.synchronize()
    lib.b200voc_debug_set_trace(0)
t = trace.cpu().view(16, 5, 32)
nblk = {"32": 3, "64a": 1, "64b": 2}[which]
nt = 4 if which == "32" else 2
for bs in range(4, 10):
    t0 = int(t[bs, 0, 0])
    r = lambda role, ev: (int(t[bs, role, ev]) - t0) if int(t[bs, role, ev]) else -1
    print(f"--- band-strip {bs}: total to next start {int(t[bs + 1, 0, 0]) - t0} clk")
    print(f" issuer: start 0 | in_full {r(0,1)} | CT issued {r(0,2)}")
    for blk in range(nblk):
        print(f"  blk {blk}: G1 issue {[r(0, 3 + blk * 8 + k) for k in range(nt)]}  G2 issue {[r(0, 3 + blk * 8 + 4 + k) for k in range(nt)]}")
    if which == "32":
        print(f"  Z issue {[r(0, 27 + k) for k in range(nt)]}")
    for eg in range(4):
        line = f" EG{eg}: ct_full {r(1 + eg, 0)} ct_done {r(1 + eg, 1)} |"
        for blk in range(nblk):
            line += f" b{blk}: d1 {r(1 + eg, 2 + blk * 4)} glu {r(1 + eg, 3 + blk * 4)} d2 {r(1 + eg, 4 + blk * 4)} e2 {r(1 + eg, 5 + blk * 4)} |"
        if which == "32":
            line += f" z {r(1 + eg, 14)} done {r(1 + eg, 15)}"
        print(line)
        if nblk > 1:
            print(f"      b1 detail: d1 {r(1+eg,6)} | ld {r(1+eg,16)} glu-math {r(1+eg,17)} st-wait {r(1+eg,18)} arrive {r(1+eg,7)} || d2 {r(1+eg,8)} | ld {r(1+eg,19)} math {r(1+eg,20)} sts {r(1+eg,21)} fence {r(1+eg,22)} arrive {r(1+eg,9)}")
