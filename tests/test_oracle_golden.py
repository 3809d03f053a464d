"""CPU: the oracle restatement against the committed golden vectors (made by oracle/make_golden.py
from the reference's own generator.py / stft.py run in the authoring container)."""
import os

import numpy as np
import pytest
import torch

from oracle import vocoder7_oracle as O

GEN_CASES = {
    "gen_b1_t12_attn": dict(attn=True, window=None, kw={}),
    "gen_b2_t9_noattn": dict(attn=False, window=None, kw={}),
    "gen_b2_t9_noattn_drop": dict(attn=False, window=None, kw=dict(style_drop=True, emo_drop=False, w_style=0.7, w_emo=1.3)),
    "gen_b1_t16_attnwin": dict(attn=True, window=512, kw={}),
}


@pytest.fixture(scope="module")
def weights():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    return O.make_generator(O.OracleConfig(), seed=1234).state_dict()


def test_seeded_weights_match_fingerprint(weights, golden_dir):
    fp = np.load(os.path.join(golden_dir, "weights_seed1234_fingerprint.npz"))
    assert set(fp.files) == set(weights.keys())
    for k, v in weights.items():
        v = v.double()
        got = np.array([float(v.sum()), float(v.abs().sum()), float((v * v).sum())])
        np.testing.assert_allclose(got, fp[k], rtol=1e-9, atol=1e-9, err_msg=k)


@pytest.mark.parametrize("name", sorted(GEN_CASES))
def test_generator_restatement_matches_golden(name, weights, golden_dir):
    c = GEN_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = O.OracleConfig(use_attention=c["attn"], attn_window=c["window"])
    t = lambda k: torch.from_numpy(gold[k])
    taps = {}
    with torch.no_grad():
        y = O.generator_forward(weights, cfg, t("mel"), t("prosody"), t("style"), t("emotion"), taps=taps, **c["kw"])
    assert y.shape == gold["wav"].shape
    assert float((y - t("wav")).abs().max()) <= 2e-6          # fp32 round-off across BLAS thread counts
    assert float((y.double() - t("wav_fp64")).abs().max()) <= 5e-6
    assert float((taps["cond"] - t("cond")).abs().max()) <= 1e-5
    for k in gold.files:
        if k.startswith("tap_"):
            got = taps[k[4:]][0][0, :8, :64]
            assert float((got - t(k)).abs().max()) <= 1e-5, k
    assert float(y.abs().max()) < 1.0                           # tanh range (generator.py:98)


def test_generator_output_length_and_flops():
    cfg = O.OracleConfig(use_attention=False, hidden_dim=512)
    # 473.6 MFLOP per mel frame, 40.80 GFLOP per audio second at H=512
    per_frame = O.generator_flops(cfg, 1, 1) / 1e6
    assert abs(per_frame - 473.6) < 0.5
    assert abs(O.generator_flops(cfg, 1, 861) / (256 * 861 / 22050) / 1e9 - 40.80) < 0.05


def test_chunked_equals_full_with_halo():
    """Receptive field is < 6 mel frames per side: chunk + halo + discard reproduces the full output."""
    torch.manual_seed(0)
    cfg = O.OracleConfig(use_attention=False)
    sd = {k: v.double() for k, v in O.make_generator(cfg).state_dict().items()}
    mel, pros, sty, emo = [x.double() for x in O.synthetic_inputs(1, 40, seed=5)]
    with torch.no_grad():
        full = O.generator_forward(sd, cfg, mel, pros, sty, emo)
        halo, lo, hi = 6, 12, 28
        s, e = lo - halo, hi + halo
        part = O.generator_forward(sd, cfg, mel[:, :, s:e], pros[:, s:e], sty, emo)
    got = part[..., halo * 256:(halo + hi - lo) * 256]
    assert float((got - full[..., lo * 256:hi * 256]).abs().max()) < 1e-12


def test_stft_family_matches_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "stft_b3_n4000.npz"))
    wav = torch.from_numpy(gold["wav"])
    for n in (512, 1024, 2048):
        mag = O.learnable_stft_forward(wav, torch.from_numpy(gold[f"gain_{n}"]), n, 256)
        assert mag.shape == (3, n // 2 + 1, 1 + 4000 // 256)
        np.testing.assert_allclose(mag.numpy(), gold[f"mag_{n}"], rtol=0, atol=2e-4)
    lm = O.log_mel(wav.squeeze(1))
    assert float((lm - torch.from_numpy(gold["logmel"])).abs().mean()) <= 1e-5
    assert float((lm.double() - torch.from_numpy(gold["logmel_fp64"])).abs().mean()) <= 1e-4
    spec = torch.complex(torch.from_numpy(gold["spec_re"]), torch.from_numpy(gold["spec_im"]))
    rt = O.istft(spec, 1024, 256, 4000)
    assert float((rt - torch.from_numpy(gold["istft"])).abs().max()) <= 1e-5
    assert float((rt - wav.squeeze(1)).abs().max()) <= 1e-5    # iSTFT o STFT = id
    loss = O.stft_loss_forward(wav, torch.from_numpy(gold["wav2"]),
                               [torch.from_numpy(gold[f"gain_{n}"]) for n in (512, 1024, 2048)], [512, 1024, 2048], 256, 2.0)
    assert abs(float(loss) - float(gold["stft_loss"])) <= 1e-4 * abs(float(gold["stft_loss"]))


def test_mel_filterbank_properties():
    fb = O.mel_filterbank()
    assert fb.shape == (513, 80) and float(fb.min()) >= 0.0
    nz = (fb > 0).float().mean()
    assert 0.015 < float(nz) < 0.035          # ~2.4 % dense
    # each frequency bin feeds at most two adjacent mel bins
    assert int((fb > 0).sum(1).max()) <= 2


def test_emulated_fp16_plan_meets_gate():
    """The kernel numerics plan (oracle/emulate.py) predicts the gate is met with fp16 operands."""
    from oracle import emulate as E
    cfg = O.OracleConfig(use_attention=False)
    sd = O.make_generator(cfg).state_dict()
    mel, pros, sty, emo = O.synthetic_inputs(1, 96, seed=3)
    with torch.no_grad():
        ref = O.generator_forward(sd, cfg, mel, pros, sty, emo)
        y = E.emulated_forward(sd, cfg, mel, pros, sty, emo, E.PLANS["fp16"])
    assert float((y - ref).abs().max()) <= 1e-3 and O.snr_db(ref, y) >= 40.0


def test_gst_restatement_matches_reference_golden(golden_dir):
    """GlobalStyleTokens (vocoder7/gst.py): golden made by the reference class itself."""
    gold = np.load(os.path.join(golden_dir, "gst_b3_t150.npz"))
    sd = {k[3:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("sd.")}
    seeded = O.make_gst_state(seed=1234)
    assert list(seeded.keys()) == ["tokens", "attn_conv.0.weight", "attn_conv.0.bias", "attn_conv.2.weight",
                                   "attn_conv.2.bias"]
    for k in seeded:
        assert torch.equal(seeded[k], sd[k]), k
    y = O.gst_forward(sd, torch.from_numpy(gold["mel"]))
    assert float((y - torch.from_numpy(gold["style"])).abs().max()) <= 1e-6
    # the reference's softmax runs over the axis the einsum sums over: the style is sum_n tokens[n]
    assert float((y - sd["tokens"].sum(0)).abs().max()) <= 1e-5


def test_pcm16_wire_format():
    w = torch.tensor([-1.5, -1.0, -0.5, 0.0, 0.25, 0.99999, 1.0, 2.0, 1.5 / 32767.0, 2.5 / 32767.0])
    assert O.pcm16(w).tolist() == [-32767, -32767, -16384, 0, 8192, 32767, 32767, 32767, 2, 2]
