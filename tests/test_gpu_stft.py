"""GPU parity tests for the STFT family (fp32): CUDA kernels through the C ABI vs the CPU oracle
(torch.stft / torch.istft / torchaudio-equivalent mel) and the committed golden vectors.

Tolerances (north_star): log-mel mean-abs (L1) <= 1e-4; round trip istft(stft(x)) max-abs <= 1e-5.
Magnitudes are compared with an absolute tolerance scaled by sqrt(n_fft) (the unnormalised STFT of
a unit-range signal has O(sqrt(n)) magnitudes, fp32 round-off ~1e-6 relative)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vocoder7_oracle as O  # noqa: E402


def test_stft_family_matches_golden(golden_dir):
    import b200voc
    gold = np.load(os.path.join(golden_dir, "stft_b3_n4000.npz"))
    wav = torch.from_numpy(gold["wav"]).cuda()
    for n in (512, 1024, 2048):
        m = b200voc.LearnableSTFT(n, 256).cuda()
        m.load_state_dict({"window": torch.hann_window(n), "filterbank": torch.from_numpy(gold[f"gain_{n}"])})
        with torch.no_grad():      # the forward kernel has no backward and says so when a gradient is expected
            mag = m(wav).cpu()
        assert mag.shape == gold[f"mag_{n}"].shape
        assert float((mag - torch.from_numpy(gold[f"mag_{n}"])).abs().max()) <= 2e-5 * n ** 0.5 * 4
    lm = b200voc.log_mel(wav).cpu()
    assert float((lm - torch.from_numpy(gold["logmel"])).abs().mean()) <= 1e-4
    assert float((lm.double() - torch.from_numpy(gold["logmel_fp64"])).abs().mean()) <= 1e-4
    spec = torch.complex(torch.from_numpy(gold["spec_re"]), torch.from_numpy(gold["spec_im"])).cuda()
    rt = b200voc.istft(spec, 1024, 256, 4000).cpu()
    assert float((rt - torch.from_numpy(gold["istft"])).abs().max()) <= 1e-5
    loss_mod = b200voc.STFTLoss(b200voc.GANConfig()).cuda()
    for m, n in zip(loss_mod.stfts, (512, 1024, 2048)):
        m.filterbank.data.copy_(torch.from_numpy(gold[f"gain_{n}"]))
    with torch.no_grad():
        loss = float(loss_mod(wav, torch.from_numpy(gold["wav2"]).cuda()))
    assert abs(loss - float(gold["stft_loss"])) <= 1e-4 * abs(float(gold["stft_loss"]))


@pytest.mark.parametrize("n_fft", [512, 1024, 2048])
@pytest.mark.parametrize("B,N", [(1, 1025), (2, 4000), (3, 22050), (5, 2049)])
def test_stft_complex_and_roundtrip(n_fft, B, N):
    import b200voc
    if N <= n_fft // 2:
        pytest.skip("reflect padding needs N > n_fft/2")
    g = torch.Generator().manual_seed(n_fft + N)
    wav = torch.rand(B, N, generator=g) * 2 - 1
    ref = O.stft_complex(wav, n_fft, 256)
    got = b200voc.stft(wav.cuda(), n_fft, 256)
    assert got.shape == ref.shape
    tol = 2e-5 * n_fft ** 0.5 * 4
    assert float((got.cpu() - ref).abs().max()) <= tol
    back = b200voc.istft(got, n_fft, 256, N).cpu()
    assert float((back - wav).abs().max()) <= 1e-5
    # kernel iSTFT on the oracle's spectrum vs torch.istft
    ref_back = O.istft(ref, n_fft, 256, N)
    got_back = b200voc.istft(ref.cuda(), n_fft, 256, N).cpu()
    assert float((got_back - ref_back).abs().max()) <= 1e-5


def test_logmel_matches_oracle_tolerance():
    import b200voc
    g = torch.Generator().manual_seed(5)
    wav = torch.rand(4, 88200, generator=g) * 2 - 1
    ref = O.log_mel(wav)
    got = b200voc.log_mel(wav.cuda()).cpu()
    assert got.shape == ref.shape == (4, 80, 345)
    assert float((got - ref).abs().mean()) <= 1e-4             # north_star: mel L1 <= 1e-4 (log-mel)
    lin = b200voc.mel_spectrogram(wav.cuda()).cpu()
    refl = O.mel_spectrogram(wav)
    assert float(((lin - refl).abs() / (refl.abs() + 1.0)).max()) <= 1e-4


def test_stft_linearity_and_full_size_roundtrip():
    """BASELINE configs[2] size (1024 x 4 s): size-independent properties -- linearity of the complex
    STFT, and iSTFT(STFT(x)) = x -- instead of a full CPU oracle pass."""
    import b200voc
    g = torch.Generator(device="cuda").manual_seed(9)
    B, N = 1024, 88200
    x = torch.rand(B, N, generator=g, device="cuda") * 2 - 1
    y = torch.rand(B, N, generator=g, device="cuda") * 2 - 1
    sx, sy = b200voc.stft(x, 1024, 256), b200voc.stft(y, 1024, 256)
    sxy = b200voc.stft(0.5 * x - 2.0 * y, 1024, 256)
    assert sx.shape == (B, 513, 345)
    assert float((sxy - (0.5 * sx - 2.0 * sy)).abs().max()) <= 2e-3
    back = b200voc.istft(sx, 1024, 256, N)
    assert float((back - x).abs().max()) <= 1e-5
    lm = b200voc.log_mel(x)
    assert lm.shape == (B, 80, 345) and bool(torch.isfinite(lm).all())
    # spot-check 2 rows of the big batch against the oracle
    ref = O.log_mel(x[[0, 777]].cpu())
    assert float((lm[[0, 777]].cpu() - ref).abs().mean()) <= 1e-4


def test_stft_errors_are_loud():
    import b200voc
    with pytest.raises(ValueError):
        b200voc.stft(torch.zeros(1, 100).cuda(), 1024, 256)        # N <= n_fft/2: torch.stft refuses too
    with pytest.raises(ValueError):
        b200voc.stft(torch.zeros(1, 4000).cuda(), 768, 256)        # unsupported n_fft
    with pytest.raises(b200voc._lib.B200VocError):
        b200voc.stft(torch.zeros(1, 4000), 1024, 256)               # CPU tensor: no fallback


@pytest.mark.parametrize("B,N", [(2, 4000), (1, 8192), (3, 2817)])
def test_stft_loss_backward_matches_autograd(B, N):
    """STFTLoss backward: the CUDA adjoint chain against torch autograd through the
    fp64 oracle -- gradients w.r.t. the generated waveform and the learnable per-bin gains."""
    from b200voc import GANConfig, STFTLoss
    cfg = GANConfig()
    torch.manual_seed(5)
    loss_mod = STFTLoss(cfg).cuda()
    g = torch.Generator().manual_seed(N)
    fake = (torch.rand(B, 1, N, generator=g) * 2 - 1)
    real = (torch.rand(B, 1, N, generator=g) * 2 - 1)
    fake_d = fake.cuda().requires_grad_(True)
    loss = loss_mod(fake_d, real.cuda())
    loss.backward()
    # oracle in fp64 with autograd
    fbs = [m.filterbank.detach().cpu().double().requires_grad_(True) for m in loss_mod.stfts]
    fake64 = fake.double().requires_grad_(True)
    ref = O.stft_loss_forward(fake64, real.double(), fbs, list(cfg.stft_sizes), cfg.hop_length, cfg.lambda_stft)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-4 * max(1.0, abs(float(ref)))
    gw, gw_ref = fake_d.grad.cpu().double(), fake64.grad
    assert gw.shape == gw_ref.shape
    rel = float((gw - gw_ref).norm() / gw_ref.norm())
    assert rel <= 2e-3, rel                      # fp32 FFTs; isolated sign(|X_f|-|X_r|) flips near zero
    for m, fb in zip(loss_mod.stfts, fbs):
        gg, gg_ref = m.filterbank.grad.cpu().double(), fb.grad
        assert float((gg - gg_ref).abs().max()) <= 1e-4 * max(1.0, float(gg_ref.abs().max()))


@pytest.mark.parametrize("n_fft,B,N", [(512, 2, 4000), (1024, 1, 8192), (2048, 3, 2817), (1024, 2, 700)])
def test_learnable_stft_backward_matches_autograd(n_fft, B, N):
    """LearnableSTFT.forward is differentiable like the reference module (vocoder7/stft.py:22-34): an arbitrary upstream
    gradient through b200voc_stft_mag_backward against fp64 autograd of the oracle -- d/d waveform (incl. the reflect
    padding: N = 700 is shorter than two frames) and d/d filterbank."""
    import b200voc
    torch.manual_seed(n_fft + N)
    mod = b200voc.LearnableSTFT(n_fft, 256).cuda()
    g = torch.Generator().manual_seed(N)
    wav = torch.rand(B, 1, N, generator=g) * 2 - 1
    wd = wav.cuda().requires_grad_(True)
    out = mod(wd)
    assert out.requires_grad
    up = torch.randn(out.shape, generator=g)
    (out * up.cuda()).sum().backward()
    fb = mod.filterbank.detach().cpu().double().requires_grad_(True)
    w64 = wav.double().requires_grad_(True)
    ref = O.learnable_stft_forward(w64, fb, n_fft, 256)
    assert float((out.detach().cpu().double() - ref.detach()).abs().max()) <= 2e-5 * float(ref.detach().abs().max())
    (ref * up.double()).sum().backward()
    gw, gw_ref = wd.grad.cpu().double(), w64.grad
    assert gw.shape == gw_ref.shape
    assert float((gw - gw_ref).abs().max()) <= 1e-4 * float(gw_ref.abs().max())
    gg, gg_ref = mod.filterbank.grad.cpu().double(), fb.grad
    assert float((gg - gg_ref).abs().max()) <= 1e-4 * float(gg_ref.abs().max())
    # a frozen gain: only the waveform gradient is produced; no_grad still takes the plain kernel
    mod.filterbank.requires_grad_(False)
    wd2 = wav.cuda().requires_grad_(True)
    (mod(wd2) * up.cuda()).sum().backward()
    assert torch.equal(wd2.grad, wd.grad)
    with torch.no_grad():
        assert not mod(wd2).requires_grad


@pytest.mark.parametrize("hop", [110, 250, 128, 200])
def test_stft1024_hops_not_multiple_of_4(hop):
    """n_fft=1024 fast kernel with hops where 7*hop+1024 is not a multiple of 4 (hop % 4 == 2): the
    interior vector-load path must not leave the tail of the span unwritten (round-1 ADVICE)."""
    import b200voc
    g = torch.Generator().manual_seed(hop)
    wav = torch.rand(2, 8000, generator=g) * 2 - 1
    ref = O.stft_complex(wav, 1024, hop)
    got = b200voc.stft(wav.cuda(), 1024, hop).cpu()
    assert got.shape == ref.shape
    assert torch.isfinite(torch.view_as_real(got)).all()
    assert float((got - ref).abs().max()) <= 2e-5 * 32 * 4


def test_stft_loss_backward_applies_upstream_gradient_on_device():
    """the upstream gradient reaches the kernels' outputs as a device scalar (no host read of grad_out): scaling the loss
    scales every gradient"""
    import b200voc
    torch.manual_seed(5)
    mod = b200voc.STFTLoss(b200voc.GANConfig()).cuda()
    fake, real = torch.rand(2, 4096) * 2 - 1, torch.rand(2, 4096) * 2 - 1
    grads = []
    for s in (1.0, -2.5):
        f = fake.cuda().requires_grad_(True)
        for m in mod.stfts:
            m.filterbank.grad = None
        (mod(f, real.cuda()) * s).backward()
        grads.append((f.grad.clone(), [m.filterbank.grad.clone() for m in mod.stfts]))
    (g1, fb1), (g2, fb2) = grads
    assert torch.allclose(g2, -2.5 * g1, rtol=1e-6, atol=1e-12) and float(g1.abs().max()) > 0
    for a, b in zip(fb1, fb2):
        assert torch.allclose(b, -2.5 * a, rtol=1e-6, atol=1e-12)
