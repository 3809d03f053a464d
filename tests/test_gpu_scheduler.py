"""GPU: scheduler drivers on the real CUDA Generator (single device).  Chunked long-form synthesis
with halo >= 6 frames must reproduce the full-utterance output (every output sample's arithmetic is
position independent: same MMA K order, per-element epilogues), ragged batches must equal the
per-utterance results."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vocoder7_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def gen():
    from b200voc import GANConfig, Generator
    ora = O.make_generator(O.OracleConfig(use_attention=False), seed=1234)
    g = Generator(GANConfig(use_attention=False)).eval()
    g.load_state_dict(ora.state_dict())
    return g.cuda()


def test_synthesize_long_equals_full(gen):
    from b200voc import scheduler as S
    mel, pros, sty, emo = [x.cuda() for x in O.synthetic_inputs(2, 300, seed=3)]
    with torch.no_grad():
        full = gen(mel, pros, sty, emo).clone()
        chunked = S.synthesize_long(gen, mel, pros, sty, emo, chunk_frames=64, halo=8, max_batch=8)
    assert chunked.shape == full.shape
    assert float((chunked - full).abs().max()) <= 1e-6
    with torch.no_grad():
        ref = O.generator_forward(O.make_generator(O.OracleConfig(use_attention=False)).state_dict(),
                                  O.OracleConfig(use_attention=False), *[x.cpu() for x in (mel, pros, sty, emo)])
    assert float((chunked.cpu() - ref).abs().max()) <= 1e-3


def test_long_form_60s_properties(gen):
    """BASELINE configs[4] shape per utterance (60 s, T=5167) at batch 2: finite, in range, and the
    chunked result equals the full forward."""
    from b200voc import scheduler as S
    mel, pros, sty, emo = [x.cuda() for x in O.synthetic_inputs(2, 5167, seed=60)]
    with torch.no_grad():
        chunked = S.synthesize_long(gen, mel, pros, sty, emo, chunk_frames=512, halo=8, max_batch=16)
        full = gen(mel, pros, sty, emo)
    assert chunked.shape == (2, 1, 256 * 5167)
    assert bool(torch.isfinite(chunked).all()) and float(chunked.abs().max()) < 1.0
    assert float((chunked - full).abs().max()) <= 1e-6


def test_synthesize_batch_ragged(gen):
    from b200voc import scheduler as S
    items = [O.synthetic_inputs(1, T, seed=T) for T in (40, 77, 40, 13, 77)]
    mels = [m[0].cuda() for m, _, _, _ in items]
    pros = [p[0].cuda() for _, p, _, _ in items]
    stys = [s[0].cuda() for _, _, s, _ in items]
    emos = [e[0].cuda() for _, _, _, e in items]
    with torch.no_grad():
        outs = S.synthesize_batch(gen, mels, pros, stys, emos, max_batch=2)
        for (m, p, s, e), w in zip(items, outs):
            one = gen(m.cuda(), p.cuda(), s.cuda(), e.cuda())[0]
            assert w.shape == one.shape and torch.equal(w, one)


def test_streaming_synthesizer_matches_direct_calls(gen):
    """host-in / host-out pipelined loop (uploads and downloads overlapping the kernels): every batch
    equals the direct call bit for bit, including when slots and host buffers are recycled."""
    from b200voc import scheduler as S
    batches = [tuple(t.pin_memory() for t in O.synthetic_inputs(3, 50, seed=200 + k)) for k in range(5)]
    outs = [torch.empty(3, 1, 256 * 50).pin_memory() for _ in batches]
    with torch.no_grad():
        S.StreamingSynthesizer(gen, torch.device("cuda"), depth=2).run(batches, outs)
        torch.cuda.synchronize()
        for b, o in zip(batches, outs):
            want = gen(*[t.cuda() for t in b]).cpu()
            assert torch.equal(o, want)
