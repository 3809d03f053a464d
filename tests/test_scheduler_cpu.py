"""CPU: host-side scheduler logic, including the N>1 path on gloo (world_size 2) with the CPU
oracle standing in for the CUDA Generator (same call signature)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200voc import scheduler as S
from oracle import vocoder7_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_shards_balanced_and_complete():
    lengths = [861, 120, 500, 500, 30, 861, 77, 400, 860, 1]
    for world in (1, 2, 4, 8):
        shards = S.plan_shards(lengths, world)
        assert sorted(i for s in shards for i in s) == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)
    assert S.plan_shards([], 4) == [[], [], [], []]
    with pytest.raises(ValueError):
        S.plan_shards([1], 0)


def test_group_by_length_exact_and_bounded():
    lengths = [10, 12, 10, 10, 12, 7]
    batches = S.group_by_length(lengths, max_batch=2)
    assert sorted(i for b in batches for i in b) == list(range(6))
    for b in batches:
        assert len(b) <= 2 and len({lengths[i] for i in b}) == 1


def test_streaming_synthesizer_host_logic():
    """CPU device: the streaming loop degenerates to the sequential one (stand-in synthesizer)."""
    calls = []

    def synth(mel, pros, sty, emo, scale=1.0):
        calls.append(mel.shape)
        return (mel.sum(1, keepdim=True) * scale).repeat_interleave(4, -1)
    batches = [O.synthetic_inputs(2, 5, seed=k) for k in range(3)]
    outs = [torch.empty(2, 1, 20) for _ in batches]
    S.StreamingSynthesizer(synth, torch.device("cpu"), depth=2, scale=2.0).run(batches, outs)
    assert len(calls) == 3
    for b, o in zip(batches, outs):
        assert torch.equal(o, (b[0].sum(1, keepdim=True) * 2.0).repeat_interleave(4, -1))
    with pytest.raises(ValueError):
        S.StreamingSynthesizer(synth, torch.device("cpu")).run(batches, outs[:2])
    with pytest.raises(ValueError):
        S.StreamingSynthesizer(synth, torch.device("cpu"), depth=0)


def test_chunk_plan_tiles_the_utterance():
    for T, chunk, halo in [(5167, 512, 8), (100, 512, 8), (513, 512, 6), (1024, 256, 0)]:
        plan = S.chunk_plan(T, chunk, halo)
        assert plan[0][2] == 0 and plan[-1][3] == T
        for (a, b, ks, ke), nxt in zip(plan, plan[1:] + [None]):
            assert 0 <= a <= ks < ke <= b <= T
            assert ks - a <= halo and b - ke <= halo
            if nxt is not None:
                assert ke == nxt[2]


def _oracle_synth():
    cfg = O.OracleConfig(use_attention=False)
    sd = {k: v.double() for k, v in O.make_generator(cfg, seed=1234).state_dict().items()}

    def synth(mel, pros, sty, emo, **kw):
        with torch.no_grad():
            return O.generator_forward(sd, cfg, mel.double(), pros.double(), sty.double(), emo.double(), **kw).float()
    return synth


def test_synthesize_long_equals_full_on_oracle():
    synth = _oracle_synth()
    mel, pros, sty, emo = O.synthetic_inputs(1, 70, seed=11)
    full = synth(mel, pros, sty, emo)
    got = S.synthesize_long(synth, mel, pros, sty, emo, chunk_frames=24, halo=6, max_batch=4)
    assert got.shape == full.shape
    assert float((got - full).abs().max()) <= 1e-6
    short = S.synthesize_long(synth, mel, pros, sty, emo, chunk_frames=24, halo=2)
    assert float((short - full).abs().max()) > 1e-5      # a halo below the receptive field is NOT exact


def test_synthesize_batch_ragged_matches_individual():
    synth = _oracle_synth()
    items = [O.synthetic_inputs(1, T, seed=T) for T in (9, 14, 9, 5)]
    outs = S.synthesize_batch(synth, [m[0] for m, _, _, _ in items], [p[0] for _, p, _, _ in items],
                              [s[0] for _, _, s, _ in items], [e[0] for _, _, _, e in items], max_batch=2)
    for (m, p, s, e), w in zip(items, outs):
        ref = synth(m, p, s, e)[0]
        assert w.shape == ref.shape and float((w - ref).abs().max()) <= 1e-6


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    synth = _oracle_synth()
    items = [O.synthetic_inputs(1, T, seed=100 + k) for k, T in enumerate((6, 11, 6, 8, 11))]
    args = ([m[0] for m, _, _, _ in items], [p[0] for _, p, _, _ in items], [s[0] for _, _, s, _ in items],
            [e[0] for _, _, _, e in items])
    got = S.sharded_synthesize(synth, *args, max_batch=2, gather_to=0)
    if rank == 0:
        ok = sorted(got) == list(range(5))
        for i, (m, p, s, e) in enumerate(items):
            ok = ok and bool(torch.equal(got[i], synth(m, p, s, e)[0]))     # sharded == unsharded, bit for bit
        ret[0] = ok
    else:
        ret[rank] = set(got) == set(S.plan_shards([6, 11, 6, 8, 11], world)[rank])
    dist.destroy_process_group()


def _worker_long(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    synth = _oracle_synth()
    mel, pros, sty, emo = O.synthetic_inputs(2, 50, seed=77)
    got = S.sharded_synthesize_long(synth, mel, pros, sty, emo, chunk_frames=16, halo=6, max_batch=3, gather_to=0)
    ok = True
    if rank == 0:
        single = S.synthesize_long(synth, mel, pros, sty, emo, chunk_frames=16, halo=6, max_batch=3)
        ok = got.shape == single.shape and float((got - single).abs().max()) <= 1e-6
        ok = ok and float((got - synth(mel, pros, sty, emo)).abs().max()) <= 1e-6     # == the un-chunked forward
    else:
        units, pieces = got
        ok = len(units) == len(pieces) > 0
    # fewer utterances than ranks: one shard is empty and still takes part in the gather
    one = [O.synthetic_inputs(1, 7, seed=5)]
    g1 = S.sharded_synthesize(synth, [one[0][0][0]], [one[0][1][0]], [one[0][2][0]], [one[0][3][0]], gather_to=0)
    if rank == 0:
        ok = ok and sorted(g1) == [0] and bool(torch.equal(g1[0], synth(*one[0])[0]))
    else:
        ok = ok and g1 == {}
    # streaming shard of equally long host utterances
    mels, pr, st, em = O.synthetic_inputs(5, 6, seed=9)
    lo, wavs = S.sharded_synthesize_streaming(synth, torch.device("cpu"), mels, pr, st, em, max_batch=2)
    ok = ok and lo == (0 if rank == 0 else 3) and wavs.shape[0] == (3 if rank == 0 else 2)
    ok = ok and float((wavs - synth(mels[lo:lo + wavs.shape[0]], pr[lo:lo + wavs.shape[0]], st[lo:lo + wavs.shape[0]],
                                    em[lo:lo + wavs.shape[0]])).abs().max()) <= 1e-6
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_long_form_and_ragged_gather_gloo_world2():
    """configs[4] host logic: chunk units spread over 2 ranks + one ragged gather == single-device synthesize_long ==
    the un-chunked forward; a rank with an empty shard takes part in the gather; streaming shards cover the list."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker_long, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0] is True and ret[1] is True


def test_long_units_cover_every_utterance():
    units = S.long_units(3, 5167, 512, 8)
    assert len(units) == 3 * 11
    for b in range(3):
        mine = [u for u in units if u[0] == b]
        assert mine[0][3] == 0 and mine[-1][4] == 5167
        assert all(x[4] == y[3] for x, y in zip(mine, mine[1:]))


def test_sharded_synthesis_gloo_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0] is True and ret[1] is True


def test_graphed_synthesizer_refuses_cpu_and_caller_buffers():
    """GraphedSynthesizer (one CUDA graph per request shape) has no CPU path and owns its output buffer."""
    import pytest
    from b200voc.scheduler import GraphedSynthesizer
    gs = GraphedSynthesizer(lambda *a, **k: None)
    x = [torch.zeros(1, 80, 4), torch.zeros(1, 4, 18), torch.zeros(1, 128), torch.zeros(1, 6)]
    with pytest.raises(ValueError):
        gs(*x)
    assert gs.captures == 0
