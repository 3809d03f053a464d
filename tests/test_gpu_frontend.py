"""GPU parity tests for the callers / wire formats either side of the hot path (SURVEY.md 8f ranks
1-2): GlobalStyleTokens (vocoder7/gst.py), time-major mel input, 16-bit PCM output and the padded
batch length mask -- all through the C ABI."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vocoder7_oracle as O  # noqa: E402


def _gst():
    from b200voc import GANConfig, GlobalStyleTokens
    gst = GlobalStyleTokens(GANConfig()).eval()
    gst.load_state_dict(O.make_gst_state(seed=1234))
    return gst.cuda()


def test_gst_matches_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "gst_b3_t150.npz"))
    gst = _gst()
    mel = torch.from_numpy(gold["mel"]).cuda()
    want = torch.from_numpy(gold["style"])
    with torch.no_grad():
        got = gst(mel).cpu()
        got_t = gst(mel.transpose(1, 2).contiguous(), mel_layout="BTC").cpu()
    assert got.shape == want.shape
    # fp32 everywhere; the only differences are summation order and expf vs torch's exp
    assert float((got - want).abs().max()) <= 1e-5
    assert torch.equal(got, got_t)


@pytest.mark.parametrize("B,T", [(1, 1), (2, 7), (1, 64), (3, 65), (2, 861)])
def test_gst_matches_oracle_ragged(B, T):
    gst = _gst()
    sd = O.make_gst_state(seed=1234)
    g = torch.Generator().manual_seed(T)
    mel = torch.randn(B, 80, T, generator=g) * 3.0
    with torch.no_grad():
        got = gst(mel.cuda()).cpu()
    want = O.gst_forward(sd, mel)
    assert float((got - want).abs().max()) <= 1e-5


@pytest.fixture(scope="module")
def models():
    from b200voc import GANConfig, Generator
    ocfg = O.OracleConfig(use_attention=False)
    ora = O.make_generator(ocfg, seed=1234)
    gen = Generator(GANConfig(use_attention=False)).eval()
    gen.load_state_dict(ora.state_dict())
    return ocfg, ora, gen.cuda()


def test_time_major_mel_is_bit_identical(models):
    _, _, gen = models
    mel, pros, sty, emo = [x.cuda() for x in O.synthetic_inputs(3, 37, seed=5)]
    with torch.no_grad():
        a = gen(mel, pros, sty, emo).clone()
        b = gen(mel.transpose(1, 2).contiguous(), pros, sty, emo, mel_layout="BTC").clone()
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        gen(mel, pros, sty, emo, mel_layout="BTC")          # channels on the wrong axis


def test_pcm16_output_and_length_mask(models):
    ocfg, ora, gen = models
    B, T = 4, 40
    mel, pros, sty, emo = O.synthetic_inputs(B, T, seed=6)
    lens = torch.tensor([40, 17, 1, 33])
    with torch.no_grad():
        f32 = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda()).clone()
        pcm = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda(), out_dtype=torch.int16).clone()
        msk = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda(), frame_lengths=lens.cuda()).clone()
        both = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda(), out_dtype=torch.int16, frame_lengths=lens).clone()
        ref = O.generator_forward(ora.state_dict(), ocfg, mel, pros, sty, emo)
    assert pcm.dtype == torch.int16 and pcm.shape == f32.shape
    assert torch.equal(pcm.cpu(), O.pcm16(f32.cpu()))                       # same rounding as the oracle's wire format
    assert int((pcm.cpu().int() - O.pcm16(ref).int()).abs().max()) <= 34     # 1e-3 * 32767 + rounding
    for b in range(B):
        n = int(lens[b]) * 256
        assert torch.equal(msk[b, 0, :n], f32[b, 0, :n])
        assert not bool(msk[b, 0, n:].any())
        assert torch.equal(both[b, 0, :n], pcm[b, 0, :n]) and not bool(both[b, 0, n:].any())


def test_weight_updates_through_data_need_invalidate_and_copies_are_independent():
    """round-1 ADVICE: (1) in-place updates through .data (the reference trainer's EMA update, vocoder7/trainer.py:53-55)
    do not bump the version counter the weight pack is keyed on -- Generator.invalidate() forces the re-pack;
    (2) copy.deepcopy (the EMA / eval copy idiom) must not share the native handle."""
    import copy
    from b200voc import GANConfig, Generator
    torch.manual_seed(3)
    gen = Generator(GANConfig(use_attention=False)).eval().cuda()
    ins = [x.cuda() for x in O.synthetic_inputs(1, 12, seed=5)]
    with torch.no_grad():
        a = gen(*ins).clone()
        for p in gen.parameters():
            p.data.mul_(0.5)
        gen.invalidate()
        b = gen(*ins).clone()
        assert not torch.equal(a, b)
        twin = copy.deepcopy(gen)
        assert twin._handle is None and gen._handle is not None
        c = twin(*ins).clone()
        assert twin._handle != gen._handle
        assert torch.equal(b, c)
        for p in twin.parameters():
            p.data.mul_(2.0)
        twin.invalidate()
        assert torch.equal(gen(*ins), b)                   # the original still runs ITS weights
        assert torch.equal(twin(*ins), a)                  # x0.5 then x2.0 is exact in binary floating point
    del twin


def test_learnable_stft_is_differentiable_like_the_reference():
    """vocoder7/stft.py:22-34 is differentiable; so is the drop-in (round-1 ADVICE: it used to return a tensor without
    grad_fn, then raised): gradients reach the waveform and the gains (values: tests/test_gpu_stft.py)."""
    import b200voc
    m = b200voc.LearnableSTFT(1024, 256).cuda().train()
    wav = torch.rand(1, 1, 4000, device="cuda", requires_grad=True)
    out = m(wav)
    assert out.shape == (1, 513, 16) and out.grad_fn is not None
    out.sum().backward()
    assert wav.grad is not None and m.filterbank.grad is not None and bool(torch.isfinite(wav.grad).all())
    with torch.no_grad():
        assert m(wav).shape == (1, 513, 16)


def test_forward_and_stft_are_cuda_graph_capturable():
    """the header's promise: all launches go to the given stream, nothing allocates or synchronises in the forward
    calls (after one warm-up call, and b200voc.stft_prepare for the STFT tables) -> both the Generator forward and the
    STFT -> log-mel transform can be captured in a CUDA graph and replayed on new inputs.  This is what makes small
    requests (BASELINE configs[0]: B = 1, T = 172, ~0.1 ms of tensor work behind 16 launches) launch-latency free."""
    import b200voc
    from b200voc import GANConfig, Generator
    ora = O.make_generator(O.OracleConfig(use_attention=False), seed=1234)
    gen = Generator(GANConfig(use_attention=False)).eval()
    gen.load_state_dict(ora.state_dict())
    gen = gen.cuda()
    B, T = 1, 172
    static = [x.cuda() for x in O.synthetic_inputs(B, T, seed=1)]
    out = torch.empty(B, 1, 256 * T, device="cuda")
    b200voc.stft_prepare(1024, 80, 22050)
    lm_out = None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.no_grad(), torch.cuda.stream(side):
        for _ in range(2):                                  # warm-up on the capture stream (packing, func attributes, workspace)
            gen(*static, out=out)
            b200voc.log_mel(out[:, 0])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(graph):
        gen(*static, out=out)
        lm_out = b200voc.log_mel(out[:, 0])
    fresh = [x.cuda() for x in O.synthetic_inputs(B, T, seed=2)]
    for s, f in zip(static, fresh):
        s.copy_(f)
    graph.replay()
    torch.cuda.synchronize()
    got_wav, got_lm = out.clone(), lm_out.clone()
    with torch.no_grad():
        want_wav = gen(*fresh)
        want_lm = b200voc.log_mel(want_wav[:, 0])
    assert torch.equal(got_wav, want_wav)
    assert torch.equal(got_lm, want_lm)
    # replay latency vs eager launch latency for the small request
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        graph.replay()
    e1.record()
    with torch.no_grad():
        for _ in range(20):
            gen(*static, out=out)
            b200voc.log_mel(out[:, 0])
    e2.record()
    torch.cuda.synchronize()
    print(f"B=1 T=172 generator + log-mel: graph replay {e0.elapsed_time(e1) / 20:.3f} ms, eager {e1.elapsed_time(e2) / 20:.3f} ms")


def test_graphed_synthesizer_matches_eager_across_shapes():
    """scheduler.GraphedSynthesizer: one CUDA graph per exact request shape, replayed on static buffers -- bit-identical
    to the eager forward, also when shapes alternate (each graph keeps its own workspace alive) and when the least
    recently used graph is evicted and re-captured."""
    from b200voc import GANConfig, Generator
    from b200voc.scheduler import GraphedSynthesizer
    ora = O.make_generator(O.OracleConfig(use_attention=False), seed=1234)
    gen = Generator(GANConfig(use_attention=False)).eval()
    gen.load_state_dict(ora.state_dict())
    gen = gen.cuda()
    gs = GraphedSynthesizer(gen, max_graphs=2)
    shapes = [(1, 20), (2, 33), (1, 20), (3, 9), (2, 33), (1, 20)]
    for i, (B, T) in enumerate(shapes):
        ins = [x.cuda() for x in O.synthetic_inputs(B, T, seed=10 + i)]
        got = gs(*ins).clone()
        with torch.no_grad():
            want = gen(*ins)
        assert torch.equal(got, want), (i, B, T)
    assert gs.captures == 5 and len(gs._graphs) == 2       # (1,20) was evicted by (3,9) and captured again
    ins = [x.cuda() for x in O.synthetic_inputs(1, 20, seed=99)]
    got16 = gs(*ins, out_dtype=torch.int16).clone()         # flags / formats are part of the key
    with torch.no_grad():
        assert torch.equal(got16, gen(*ins, out_dtype=torch.int16))
    with pytest.raises(ValueError):
        gs(*[x.cpu() for x in ins])


@pytest.mark.parametrize("B,T,dtype", [(1, 9, torch.float32), (3, 23, torch.float32), (2, 16, torch.int16)])
def test_forward_writes_nothing_outside_the_output(B, T, dtype):
    """the waveform is written by the fused last-stage kernel strip by strip (halo rows are recomputed, never stored):
    sentinel guards on both sides of ``out=`` must survive, for float and PCM16 outputs and ragged strip counts"""
    from b200voc import GANConfig, Generator
    ora = O.make_generator(O.OracleConfig(use_attention=False), seed=1234)
    gen = Generator(GANConfig(use_attention=False)).eval()
    gen.load_state_dict(ora.state_dict())
    gen = gen.cuda()
    ins = [x.cuda() for x in O.synthetic_inputs(B, T, seed=3)]
    n, G = B * 256 * T, 8192
    sentinel = 12345 if dtype == torch.int16 else -777.0
    big = torch.full((n + 2 * G,), sentinel, device="cuda", dtype=dtype)
    out = big[G:G + n].view(B, 1, 256 * T)
    with torch.no_grad():
        got = gen(*ins, out=out, out_dtype=dtype)
        want = gen(*ins, out_dtype=dtype)
    torch.cuda.synchronize()
    assert bool((big[:G] == sentinel).all()) and bool((big[G + n:] == sentinel).all())
    assert torch.equal(got, want)
