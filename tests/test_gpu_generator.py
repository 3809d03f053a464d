"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the C ABI against the CPU
oracle and the committed golden vectors.

Tolerances: the Generator computes with fp16 tensor-core operands / fp32 accumulation and 16-bit
inter-layer storage; north_star's gate is waveform max-abs <= 1e-3 and SNR >= 40 dB vs the fp32
oracle.  Layer-level tests compare against an fp64 evaluation on the SAME rounded operands, so the
only error left is the output rounding to the 16-bit storage format (<= 2^-11 relative for fp16,
2^-8 for bf16) plus fp32 accumulation order.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import vocoder7_oracle as O  # noqa: E402


def _lib():
    from b200voc import _lib
    return _lib, _lib.load()


@pytest.fixture(scope="module")
def models():
    from b200voc import GANConfig, Generator
    ocfg = O.OracleConfig(use_attention=False)
    ora = O.make_generator(ocfg, seed=1234)
    gen = Generator(GANConfig(use_attention=False)).eval()
    gen.load_state_dict(ora.state_dict())
    return ocfg, ora, gen.cuda()


@pytest.mark.parametrize("name,kw", [
    ("gen_b2_t9_noattn", {}),
    ("gen_b2_t9_noattn_drop", dict(style_drop=True, emo_drop=False, w_style=0.7, w_emo=1.3)),
])
def test_generator_matches_golden(models, golden_dir, name, kw):
    _, _, gen = models
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    t = lambda k: torch.from_numpy(gold[k]).cuda()
    with torch.no_grad():
        wav = gen(t("mel"), t("prosody"), t("style"), t("emotion"), **kw).cpu()
    ref = torch.from_numpy(gold["wav"])
    assert wav.shape == ref.shape
    assert float((wav - ref).abs().max()) <= 1e-3
    assert O.snr_db(ref, wav) >= 40.0
    assert gen.launch_count() >= 16          # the CUDA path actually ran (16 launches with the fused narrow stages, 22 without)


@pytest.mark.parametrize("B,T", [(1, 1), (1, 7), (3, 33), (2, 172), (5, 61)])
def test_generator_matches_oracle_ragged_sizes(models, B, T):
    """edge cases: single frame, sizes that are not multiples of the 128-row tile, odd batch."""
    ocfg, ora, gen = models
    mel, pros, sty, emo = O.synthetic_inputs(B, T, seed=100 + T)
    with torch.no_grad():
        ref = O.generator_forward(ora.state_dict(), ocfg, mel, pros, sty, emo)
        wav = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda()).cpu()
    assert wav.shape == (B, 1, 256 * T)
    assert float((wav - ref).abs().max()) <= 1e-3
    assert O.snr_db(ref, wav) >= 40.0


def test_generator_intermediates(models):
    ocfg, ora, gen = models
    B, T = 2, 20
    mel, pros, sty, emo = O.synthetic_inputs(B, T, seed=8)
    taps = {}
    with torch.no_grad():
        O.generator_forward(ora.state_dict(), ocfg, mel, pros, sty, emo, taps=taps)
        for k in ["cond", "split", "up0", "res0.2", "up1", "res1.0", "res2.2", "up3", "res3.2"]:
            _, t = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda(), _tap=k)
            if k == "cond":
                got, want = t.view(B, T, 128).transpose(1, 2).cpu(), taps["cond"]
                tol = 1e-5
            else:
                want = torch.stack(taps[k], 1)
                want = want.reshape(-1, want.shape[2], want.shape[3])
                got = t.view(want.shape).cpu()
                tol = 4e-3 * max(1.0, float(want.abs().max()))
            assert float((got - want).abs().max()) <= tol, k


def test_generator_batch_independence_and_determinism(models):
    """utterances are independent units: a batch equals its rows run alone, bit for bit (this is
    what makes the multi-GPU sharding exact), and repeated runs are bit-identical."""
    _, _, gen = models
    mel, pros, sty, emo = [x.cuda() for x in O.synthetic_inputs(4, 50, seed=77)]
    with torch.no_grad():
        full = gen(mel, pros, sty, emo).clone()
        again = gen(mel, pros, sty, emo).clone()
        rows = torch.cat([gen(mel[i:i + 1], pros[i:i + 1], sty[i:i + 1], emo[i:i + 1]).clone() for i in range(4)])
    assert torch.equal(full, again)
    assert torch.equal(full, rows)


@pytest.mark.parametrize("plan,maxabs,snr", [("fp16", 1e-3, 40.0), ("mixed", 1.5e-3, 40.0), ("bf16", 7e-3, 40.0)])
def test_precision_plans(plan, maxabs, snr):
    """fp16 (default) meets both gates; bf16 meets the SNR gate only (plain bf16 operands
    cannot reach max-abs 1e-3 on this network, measured 5.7e-3 / 42.2 dB on these inputs, 4.3e-3 / 49 dB in the bench line: the max-abs bound
    sits just above so that a regression of the bf16 path shows)."""
    from b200voc import GANConfig, Generator
    ocfg = O.OracleConfig(use_attention=False)
    ora = O.make_generator(ocfg, seed=1234)
    gen = Generator(GANConfig(use_attention=False, precision=plan)).eval()
    gen.load_state_dict(ora.state_dict())
    gen = gen.cuda()
    mel, pros, sty, emo = O.synthetic_inputs(2, 64, seed=21)
    with torch.no_grad():
        ref = O.generator_forward(ora.state_dict(), ocfg, mel, pros, sty, emo)
        wav = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda()).cpu()
    err, got_snr = float((wav - ref).abs().max()), O.snr_db(ref, wav)
    assert err <= maxabs and got_snr >= snr, f"{plan}: max-abs {err:.3e} (bound {maxabs}), SNR {got_snr:.1f} dB (bound {snr})"


def test_generator_full_size_properties(models):
    """BASELINE configs[1] size (B=16, T=861): size-independent properties instead of a 16x CPU
    oracle run -- output range, finite, and one utterance of the batch equals the same utterance
    synthesised alone (which IS checked against the oracle)."""
    ocfg, ora, gen = models
    B, T = 16, 861
    mel, pros, sty, emo = O.synthetic_inputs(B, T, seed=4321)
    with torch.no_grad():
        wav = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda()).clone()
        assert wav.shape == (B, 1, 220416)
        assert bool(torch.isfinite(wav).all()) and float(wav.abs().max()) < 1.0
        i = 11
        one = gen(mel[i:i + 1].cuda(), pros[i:i + 1].cuda(), sty[i:i + 1].cuda(), emo[i:i + 1].cuda())
        assert torch.equal(one[0], wav[i])
        ref = O.generator_forward(ora.state_dict(), ocfg, mel[i:i + 1], pros[i:i + 1], sty[i:i + 1], emo[i:i + 1])
    assert float((one.cpu() - ref).abs().max()) <= 1e-3
    assert O.snr_db(ref, one.cpu()) >= 40.0


def test_errors_are_loud(models):
    _, _, gen = models
    with pytest.raises(ValueError):
        gen(torch.zeros(1, 79, 4).cuda(), torch.zeros(1, 4, 18).cuda(), torch.zeros(1, 128).cuda(), torch.zeros(1, 6).cuda())
    with pytest.raises(ValueError):
        gen(torch.zeros(1, 80, 4).cuda(), torch.zeros(1, 5, 18).cuda(), torch.zeros(1, 128).cuda(), torch.zeros(1, 6).cuda())


# ------------------------------------------------------------------ layer level, via the C ABI
@pytest.mark.parametrize("Cin,Cout,s", [(512, 256, 8), (256, 128, 8), (128, 64, 2), (64, 32, 2)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_convt1d_layer(Cin, Cout, s, fmt):
    _l, lib = _lib()
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    N, Lin = 3, 150
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(N, Cin, Lin, generator=g) * 0.5).to(dt)
    w = torch.randn(Cin, Cout, 2 * s, generator=g) / (2 * Cin) ** 0.5
    bias = torch.randn(Cout, generator=g) * 0.1
    ref = F.conv_transpose1d(x.double(), w.to(dt).double(), bias.double(), stride=s, padding=s // 2)
    x_cl = x.transpose(1, 2).contiguous().cuda()
    wp = torch.empty(lib.b200voc_convt_packed_elems(Cin, Cout, s), dtype=dt, device="cuda")
    wd, bd = w.cuda(), bias.cuda()
    st = _l.current_stream()
    _l.check(lib.b200voc_pack_convt_weight(_l.ptr(wd), Cin, Cout, s, fmt, _l.ptr(wp), st))
    out = torch.full((N, s * Lin, Cout), float("nan"), dtype=dt, device="cuda")
    _l.check(lib.b200voc_convt1d(_l.ptr(x_cl), _l.ptr(wp), _l.ptr(bd), N, Lin, Cin, Cout, s, fmt, 0, _l.ptr(out), st))
    torch.cuda.synchronize()
    got = out.float().cpu().transpose(1, 2).double()
    ulp = 2.0 ** -11 if fmt == 0 else 2.0 ** -8
    assert not bool(torch.isnan(got).any())
    assert float(((got - ref).abs() / (ref.abs() + 1.0)).max()) <= 1.2 * ulp


@pytest.mark.parametrize("C,d,fmt,store_lrelu", [(C, d, 0, 0) for C in (32, 64, 128, 256) for d in (1, 3, 5)] +
                         [(32, 3, 1, 0), (64, 5, 1, 0), (128, 3, 1, 0), (32, 5, 0, 1), (64, 1, 0, 1), (128, 5, 0, 1),
                          (256, 3, 0, 1)])
def test_resblock_layer(C, d, fmt, store_lrelu):
    """every channel width / dilation in fp16, plus bf16 operands and the leaky_relu-stored output form"""
    _l, lib = _lib()
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    B, nb, T, P = 2, 4, 6, 55
    N, L = B * nb, T * P
    g = torch.Generator().manual_seed(10 + d)
    x = torch.randn(N, C, L, generator=g) * 0.5
    if lib.b200voc_resblock_input_is_lrelu(C):     # wide stages carry leaky_relu(x), narrow ones raw x
        a = F.leaky_relu(x, 0.1).to(dt)
        xr = torch.where(a.float() >= 0, a.float(), a.float() * 10.0)
    else:
        a = x.to(dt)
        xr = a.float()
    cond = torch.randn(B, 128, T, generator=g)
    wc = torch.randn(2 * C, C, 3, generator=g) / (3 * C) ** 0.5
    bc = torch.randn(2 * C, generator=g) * 0.1
    wf = torch.randn(2 * C, 128, 1, generator=g) / 128 ** 0.5
    bf = torch.randn(2 * C, generator=g) * 0.1
    wp_ = torch.randn(C, C, 1, generator=g) / C ** 0.5
    bp = torch.randn(C, generator=g) * 0.1
    xq = xr.double().view(B, nb, C, L)
    ref = torch.stack([O.residual_block_forward(xq[:, k], cond.double(), wc.to(dt).double(), bc.double(), wf.double(),
                                                bf.double(), wp_.to(dt).double(), bp.double(), d) for k in range(nb)], 1)
    ref = ref.reshape(N, C, L)
    film = F.conv1d(cond, wf, bf)
    film[:, :C] += 1.0
    film_cl = film.transpose(1, 2).contiguous().cuda()
    a_cl = a.transpose(1, 2).contiguous().cuda()
    wpk = torch.empty(lib.b200voc_resblock_packed_elems(C), dtype=dt, device="cuda")
    st = _l.current_stream()
    wcd, wpd, bcd, bpd = wc.cuda(), wp_.cuda(), bc.cuda(), bp.cuda()
    _l.check(lib.b200voc_pack_resblock_weights(_l.ptr(wcd), _l.ptr(wpd), C, fmt, _l.ptr(wpk), st))
    out = torch.full((N, L, C), float("nan"), dtype=dt, device="cuda")
    _l.check(lib.b200voc_resblock(_l.ptr(a_cl), _l.ptr(wpk), _l.ptr(bcd), _l.ptr(bpd), _l.ptr(film_cl), N, L, C, d, T,
                                  nb, fmt, store_lrelu, _l.ptr(out), st))
    torch.cuda.synchronize()
    got = out.float().cpu().transpose(1, 2).double()
    assert not bool(torch.isnan(got).any())
    if store_lrelu:
        ref = F.leaky_relu(ref, 0.1)
    # h is rounded to 16 bits before GEMM2 and the output to 16 bits: a few ulps of the output scale
    ulp = 2.0 ** -11 if fmt == 0 else 2.0 ** -8
    assert float(((got - ref).abs() / (ref.abs() + 1.0)).max()) <= 3 * ulp


@pytest.mark.parametrize("C,fmt", [(128, 0), (256, 0), (128, 1), (64, 0)])
def test_resblock_layer_large_biases(C, fmt):
    """the biases are added on the tensor core as [b/2 hi, b/2 lo] columns against an all-ones operand (resblock3.cu,
    stage_fused.cu): biases of magnitude ~5-8 with awkward mantissas must come through (a missing or misplaced bias
    column is an O(1) error).  With |h| ~ 5-10 the 16-bit rounding of h ahead of GEMM2 is itself several output ulps
    at |out| ~ 1, hence 8 ulps here instead of the 3 of the unit-scale test."""
    _l, lib = _lib()
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    B, nb, T, P, d = 1, 4, 3, 128, 3
    N, L = B * nb, T * P
    g = torch.Generator().manual_seed(77 + C)
    x = torch.randn(N, C, L, generator=g) * 0.5
    if lib.b200voc_resblock_input_is_lrelu(C):
        a = F.leaky_relu(x, 0.1).to(dt)
        xr = torch.where(a.float() >= 0, a.float(), a.float() * 10.0)
    else:
        a = x.to(dt)
        xr = a.float()
    cond = torch.randn(B, 128, T, generator=g)
    wc = torch.randn(2 * C, C, 3, generator=g) / (3 * C) ** 0.5
    bc = torch.randn(2 * C, generator=g) * 2.0 + 1.2345678      # value biases ~ +-5, gate biases keep the gates open / shut
    wf = torch.randn(2 * C, 128, 1, generator=g) / 128 ** 0.5
    bf = torch.randn(2 * C, generator=g) * 0.1
    wp_ = torch.randn(C, C, 1, generator=g) / C ** 0.5
    bp = torch.randn(C, generator=g) * 8.0 + 0.0123456
    xq = xr.double().view(B, nb, C, L)
    ref = torch.stack([O.residual_block_forward(xq[:, k], cond.double(), wc.to(dt).double(), bc.double(), wf.double(),
                                                bf.double(), wp_.to(dt).double(), bp.double(), d) for k in range(nb)], 1)
    ref = ref.reshape(N, C, L)
    film = F.conv1d(cond, wf, bf)
    film[:, :C] += 1.0
    film_cl = film.transpose(1, 2).contiguous().cuda()
    a_cl = a.transpose(1, 2).contiguous().cuda()
    wpk = torch.empty(lib.b200voc_resblock_packed_elems(C), dtype=dt, device="cuda")
    st = _l.current_stream()
    wcd, wpd, bcd, bpd = wc.cuda(), wp_.cuda(), bc.cuda(), bp.cuda()
    _l.check(lib.b200voc_pack_resblock_weights(_l.ptr(wcd), _l.ptr(wpd), C, fmt, _l.ptr(wpk), st))
    out = torch.full((N, L, C), float("nan"), dtype=dt, device="cuda")
    _l.check(lib.b200voc_resblock(_l.ptr(a_cl), _l.ptr(wpk), _l.ptr(bcd), _l.ptr(bpd), _l.ptr(film_cl), N, L, C, d, T,
                                  nb, fmt, 0, _l.ptr(out), st))
    torch.cuda.synchronize()
    got = out.float().cpu().transpose(1, 2).double()
    assert not bool(torch.isnan(got).any())
    ulp = 2.0 ** -11 if fmt == 0 else 2.0 ** -8
    assert float(((got - ref).abs() / (ref.abs() + 1.0)).max()) <= 8 * ulp


@pytest.mark.parametrize("C", [128, 256, 64])
def test_resblock_layer_odd_tile_count(C):
    """an odd number of 128-row tiles: the CTA-pair kernel's peer CTA gets one all-out-of-bounds tile
    (TMA zero-fills its loads and clips its stores); also num_bands = 1 and a single FiLM frame per sequence."""
    _l, lib = _lib()
    dt = torch.float16
    N, T, P, d = 3, 1, 300, 3
    L = T * P                                    # 3 tiles per sequence -> 9 tiles
    g = torch.Generator().manual_seed(C)
    x = torch.randn(N, C, L, generator=g) * 0.5
    if lib.b200voc_resblock_input_is_lrelu(C):
        a = F.leaky_relu(x, 0.1).to(dt)
        xr = torch.where(a.float() >= 0, a.float(), a.float() * 10.0)
    else:
        a = x.to(dt)
        xr = a.float()
    cond = torch.randn(N, 128, T, generator=g)
    wc = torch.randn(2 * C, C, 3, generator=g) / (3 * C) ** 0.5
    bc = torch.randn(2 * C, generator=g) * 0.1
    wf = torch.randn(2 * C, 128, 1, generator=g) / 128 ** 0.5
    bf = torch.randn(2 * C, generator=g) * 0.1
    wp_ = torch.randn(C, C, 1, generator=g) / C ** 0.5
    bp = torch.randn(C, generator=g) * 0.1
    ref = O.residual_block_forward(xr.double(), cond.double(), wc.to(dt).double(), bc.double(), wf.double(),
                                   bf.double(), wp_.to(dt).double(), bp.double(), d)
    film = F.conv1d(cond, wf, bf)
    film[:, :C] += 1.0
    film_cl = film.transpose(1, 2).contiguous().cuda()
    a_cl = a.transpose(1, 2).contiguous().cuda()
    wpk = torch.empty(lib.b200voc_resblock_packed_elems(C), dtype=dt, device="cuda")
    st = _l.current_stream()
    wcd, wpd, bcd, bpd = wc.cuda(), wp_.cuda(), bc.cuda(), bp.cuda()
    _l.check(lib.b200voc_pack_resblock_weights(_l.ptr(wcd), _l.ptr(wpd), C, 0, _l.ptr(wpk), st))
    guard = torch.full((N * L * C + 4096,), float("nan"), dtype=dt, device="cuda")     # detects writes past the end
    out = guard[:N * L * C].view(N, L, C)
    _l.check(lib.b200voc_resblock(_l.ptr(a_cl), _l.ptr(wpk), _l.ptr(bcd), _l.ptr(bpd), _l.ptr(film_cl), N, L, C, d, T,
                                  1, 0, 0, _l.ptr(out), st))
    torch.cuda.synchronize()
    got = out.float().cpu().transpose(1, 2).double()
    assert not bool(torch.isnan(got).any())
    assert bool(torch.isnan(guard[N * L * C:]).all())
    assert float(((got - ref).abs() / (ref.abs() + 1.0)).max()) <= 3 * 2.0 ** -11


def test_rowshifted_umma_descriptors():
    """a K-major SWIZZLE_128B descriptor may start at any 128-byte row of a TMA-written
    tile with base_offset = 0 (the swizzle is a function of the absolute smem address)."""
    import ctypes
    _l, _ = _lib()
    if not os.path.exists(_l.DEV_LIB_PATH):
        pytest.skip("development build (make dev) not present")
    lib = ctypes.CDLL(_l.DEV_LIB_PATH)          # the experiment export lives in the dev build only
    lib.b200voc_exp_rowshift.restype = ctypes.c_int
    lib.b200voc_exp_rowshift.argtypes = [ctypes.c_void_p] * 4
    g = torch.Generator().manual_seed(0)
    a = torch.randn(144, 64, generator=g).half().cuda()
    b = torch.randn(64, 64, generator=g).half().cuda()
    out = torch.zeros(2, 16, 128, 64, device="cuda")
    assert lib.b200voc_exp_rowshift(_l.ptr(a), _l.ptr(b), _l.ptr(out), _l.current_stream()) == 0
    torch.cuda.synchronize()
    for s in range(16):
        ref = a[s:s + 128].float() @ b.float().t()
        assert float((out[0, s] - ref).abs().max()) <= 1e-3


# ------------------------------------------------------------------ SelfAttention (K6)
def _attn_models(window=None):
    from b200voc import GANConfig, Generator
    ocfg = O.OracleConfig(use_attention=True, attn_window=window)
    ora = O.make_generator(ocfg, seed=1234)
    gen = Generator(GANConfig(use_attention=True, attn_window=window)).eval()
    gen.load_state_dict(ora.state_dict())
    return ocfg, ora, gen.cuda()


@pytest.mark.parametrize("name,window", [("gen_b1_t12_attn", None), ("gen_b1_t16_attnwin", 512)])
def test_generator_with_attention_matches_golden(golden_dir, name, window):
    """generator.py:43-44,91-92: SelfAttention after the residual blocks of stage 2 (global, and
    block-local with a 512-position window)."""
    _, _, gen = _attn_models(window)
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    t = lambda k: torch.from_numpy(gold[k]).cuda()
    with torch.no_grad():
        wav, tap = gen(t("mel"), t("prosody"), t("style"), t("emotion"), _tap="attn")
    ref = torch.from_numpy(gold["wav"])
    assert float((wav.cpu() - ref).abs().max()) <= 1e-3
    assert O.snr_db(ref, wav.cpu()) >= 40.0
    want = torch.from_numpy(gold["tap_attn"])                  # band 0, batch 0, 8 channels, 64 steps
    got = tap.view(-1, 64, wav.shape[-1] // 2).cpu()[0, :8, :64]
    assert float((got - want).abs().max()) <= 4e-3 * max(1.0, float(want.abs().max()))
    assert gen.launch_count() in (19, 25)    # fused narrow stages: 19; layer by layer (B200VOC_FUSED=0): 24 + band_split im2col


def test_generator_with_attention_matches_oracle_batch():
    ocfg, ora, gen = _attn_models(None)
    mel, pros, sty, emo = O.synthetic_inputs(2, 21, seed=31)
    with torch.no_grad():
        ref = O.generator_forward(ora.state_dict(), ocfg, mel, pros, sty, emo)
        wav = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda()).cpu()
    assert float((wav - ref).abs().max()) <= 1e-3
    assert O.snr_db(ref, wav) >= 40.0


@pytest.mark.parametrize("window", [None, 4096])
def test_generator_with_attention_long_sequence(window):
    """round-1 VERDICT: the attention kernel was only compared with the oracle up to L = 2688.  T = 200 -> L = 25600
    positions at stage 2 = 200 key tiles of online-softmax rescaling per query tile (global), or windows of 4096
    positions of which the last is ragged (25600 = 6 x 4096 + 1024).  L is always 128 * T at this layer, so a
    sequence length that is not a multiple of 128 cannot occur."""
    ocfg, ora, gen = _attn_models(window)
    mel, pros, sty, emo = O.synthetic_inputs(1, 200, seed=71)
    with torch.no_grad():
        ref = O.generator_forward(ora.state_dict(), ocfg, mel, pros, sty, emo)
        wav = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda()).cpu()
    assert float((wav - ref).abs().max()) <= 1e-3
    assert O.snr_db(ref, wav) >= 40.0


# ------------------------------------------------------------------ input scale / 16-bit range (round-1 VERDICT)
def _scaled_models(scale_keys=None, factor=1.0, plan="fp16"):
    from b200voc import GANConfig, Generator
    ocfg = O.OracleConfig(use_attention=False)
    ora = O.make_generator(ocfg, seed=1234)
    sd = {k: v.clone() for k, v in ora.state_dict().items()}
    if scale_keys:
        for k in sd:
            if any(s in k for s in scale_keys) and k.endswith("weight"):
                sd[k] *= factor
    gen = Generator(GANConfig(use_attention=False, precision=plan)).eval()
    gen.load_state_dict(sd)
    gen = gen.cuda()
    gen.set_overflow_check(True)
    return ocfg, sd, gen


def test_generator_log_mel_scale_inputs():
    """mels in the log-mel range (randn * 2 - 4) instead of unit-variance noise: same gates, and the
    overflow check (every stored 16-bit activation is inspected) stays silent"""
    ocfg, sd, gen = _scaled_models()
    mel, pros, sty, emo = O.synthetic_inputs(2, 64, seed=33)
    mel = mel * 2.0 - 4.0
    with torch.no_grad():
        ref = O.generator_forward(sd, ocfg, mel, pros, sty, emo)
        wav = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda()).cpu()
    assert float((wav - ref).abs().max()) <= 1e-3
    assert O.snr_db(ref, wav) >= 40.0


def test_generator_large_activations_stay_in_range():
    """band_split weights x4 -> every activation of the network ~4x larger (max |x| ~ 8): still far inside the fp16
    range, no overflow, SNR gate holds (the absolute error scales with the activations, so max-abs is checked
    relative to the 4x signal before the tanh)"""
    ocfg, sd, gen = _scaled_models(("band_split",), 4.0)
    mel, pros, sty, emo = O.synthetic_inputs(2, 64, seed=34)
    with torch.no_grad():
        ref = O.generator_forward(sd, ocfg, mel, pros, sty, emo)
        wav = gen(mel.cuda(), pros.cuda(), sty.cuda(), emo.cuda()).cpu()
    assert bool(torch.isfinite(wav).all())
    assert O.snr_db(ref, wav) >= 40.0
    assert float((wav - ref).abs().max()) <= 4e-3


def test_generator_fp16_overflow_is_detected_not_silent():
    """all upsampling / residual conv weights x6: activations grow by orders of magnitude per stage and leave the fp16
    range; with the overflow check enabled the forward fails loudly and names the layer, and the bf16 plan (8 exponent
    bits) runs the same weights without overflow"""
    from b200voc import _lib
    ocfg, sd, gen = _scaled_models(("upsample_blocks",), 6.0)
    mel, pros, sty, emo = [x.cuda() for x in O.synthetic_inputs(1, 32, seed=35)]
    with torch.no_grad():
        with pytest.raises(_lib.B200VocOverflowError) as ei:
            gen(mel, pros, sty, emo)
        assert "overflow" in str(ei.value)
        _, _, gen_bf = _scaled_models(("upsample_blocks",), 6.0, plan="bf16")
        wav = gen_bf(mel, pros, sty, emo)
    assert bool(torch.isfinite(wav).all())
