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
    # fewer utterances than ranks: ranks with an empty shard still take part in the ragged gather (round-1 ADVICE)
    few = S.sharded_synthesize(gen, args[0][:1], args[1][:1], args[2][:1], args[3][:1], gather_to=0)
    if rank == 0:
        assert sorted(few) == [0] and torch.equal(few[0].reshape(-1), got[0].reshape(-1))
    # long-form (BASELINE configs[4] in small): chunk + halo units spread over the ranks, one ragged gather, stitched
    # on rank 0 == the single-GPU synthesize_long == the un-chunked forward, bit for bit
    mel, pros, sty, emo = [x.cuda() for x in O.synthetic_inputs(3, 700, seed=60)]
    long_ = S.sharded_synthesize_long(gen, mel, pros, sty, emo, chunk_frames=128, halo=8, max_batch=8, gather_to=0)
    torch.cuda.synchronize()
    if rank == 0:
        single = S.synthesize_long(gen, mel, pros, sty, emo, chunk_frames=128, halo=8, max_batch=8)
        full = gen(mel, pros, sty, emo)
        assert torch.equal(long_, single), "sharded long-form differs from the single-GPU chunked result"
        assert float((long_ - full).abs().max()) <= 1e-6, "chunked long-form differs from the un-chunked forward"
        print(f"dist_check ok: world={world} long-form sharded == single-GPU == un-chunked, bit for bit")
    # host-in / host-out streaming shards (BASELINE configs[3] in small)
    hm, hp, hs, he = [x.pin_memory() for x in O.synthetic_inputs(10, 64, seed=9)]
    lo, wavs = S.sharded_synthesize_streaming(gen, torch.device("cuda", local), hm, hp, hs, he, max_batch=4)
    n = wavs.shape[0]
    want = gen(hm[lo:lo + n].cuda(), hp[lo:lo + n].cuda(), hs[lo:lo + n].cuda(), he[lo:lo + n].cuda()).cpu()
    assert torch.equal(wavs, want)
    if rank == 0:
        print(f"dist_check ok: world={world} streaming shards == direct calls, bit for bit")
dist.barrier()
dist.destroy_process_group()
