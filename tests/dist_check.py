"""Multi-GPU check (test infrastructure), launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
Every rank synthesizes its shard with the CUDA Generator (no collective in the math); rank 0
gathers the waveforms (the single NCCL all_gather) and checks them against its own single-GPU run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))
import torch
import torch.distributed as dist
from b200voc import GANConfig, Generator, scheduler as S
from oracle import vocoder7_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ora = O.make_generator(O.OracleConfig(use_attention=False), seed=1234)
gen = Generator(GANConfig(use_attention=False)).eval()
gen.load_state_dict(ora.state_dict())
gen = gen.cuda()
lengths = [120, 64, 120, 200, 64, 33, 200, 120, 64, 90]
items = [O.synthetic_inputs(1, T, seed=1000 + k) for k, T in enumerate(lengths)]
args = ([m[0].cuda() for m, _, _, _ in items], [p[0].cuda() for _, p, _, _ in items],
        [s[0].cuda() for _, _, s, _ in items], [e[0].cuda() for _, _, _, e in items])
with torch.no_grad():
    got = S.sharded_synthesize(gen, *args, max_batch=4, gather_to=0)
    torch.cuda.synchronize()
    if rank == 0:
        assert sorted(got) == list(range(len(lengths))), sorted(got)
        worst = 0.0
        for i, (m, p, s, e) in enumerate(items):
            one = gen(m.cuda(), p.cuda(), s.cuda(), e.cuda())[0]
            assert torch.equal(got[i].reshape(-1), one.reshape(-1)), f"utterance {i} differs from the single-GPU result"
            ref = O.generator_forward(ora.state_dict(), ora.cfg, m, p, s, e)[0]
            worst = max(worst, float((one.cpu() - ref).abs().max()))
        print(f"dist_check ok: world={world} sharded == single-GPU bit for bit; max |gpu - oracle| = {worst:.2e}")
dist.barrier()
dist.destroy_process_group()
