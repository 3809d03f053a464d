"""STFT family: drop-ins for ``vocoder7.stft.LearnableSTFT`` / ``STFTLoss`` (vocoder7/stft.py:9-54)
and the functional transforms north_star names (stft, istft, mel_spectrogram, log_mel), all running
the hand-written sm_100a kernels of csrc/stft.cu through the C ABI.  fp32 throughout; no CPU
fallback.  The transforms are inference-only; ``STFTLoss`` is differentiable w.r.t. ``wav_fake`` and
the learnable per-bin gains (its backward is a CUDA kernel chain too, b200voc_stft_l1_backward).

Semantics (identical to what the reference calls, see oracle/vocoder7_oracle.py):
  STFT  = torch.stft(center=True, pad_mode="reflect", win_length=n_fft, periodic Hann, onesided,
          unnormalised) -- i.e. torchaudio ``Spectrogram(power=None)``;
  mel   = HTK filterbank, f_min 0, f_max sr/2, norm None, power 2 (torchaudio ``MelSpectrogram``
          defaults, reference_encoder/utils.py:31-36);  log_mel = log(clamp(mel, 1e-5));
  iSTFT = torch.istft(center=True, same window, length=N).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .config import GANConfig


def _prep(wav: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(wav)
    if wav.dim() == 3 and wav.shape[1] == 1:
        wav = wav.squeeze(1)
    if wav.dim() != 2:
        raise ValueError(f"waveform must be [B,N] or [B,1,N], got {tuple(wav.shape)}")
    return wav.detach().to(torch.float32).contiguous()


def stft_prepare(n_fft: int, n_mels: int = 0, sample_rate: int = 22050, device=None) -> None:
    """Build the window / twiddle (and, for ``n_mels > 0``, mel filterbank) tables of this configuration on `device`
    now, so that the transforms never allocate or copy synchronously afterwards -- required before capturing them in a
    CUDA graph (``b200voc_stft_prepare``)."""
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        _lib.check(_lib.load().b200voc_stft_prepare(int(n_fft), int(n_mels), int(sample_rate)), "stft_prepare")


def stft_magnitude(wav: torch.Tensor, n_fft: int, hop_length: int, gain: Optional[torch.Tensor] = None) -> torch.Tensor:
    """|STFT(wav)| * gain[:, None] -> [B, n_fft/2+1, 1+N//hop]."""
    w = _prep(wav)
    B, N = w.shape
    out = torch.empty(B, n_fft // 2 + 1, 1 + N // hop_length, device=w.device, dtype=torch.float32)
    g = None if gain is None else gain.detach().to(torch.float32).contiguous()
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().b200voc_stft_mag(_lib.ptr(w), B, N, n_fft, hop_length, _lib.ptr(g), _lib.ptr(out),
                                                _lib.current_stream()), "stft_mag")
    return out


def stft(wav: torch.Tensor, n_fft: int, hop_length: int) -> torch.Tensor:
    """complex64 [B, n_fft/2+1, 1+N//hop]."""
    w = _prep(wav)
    B, N = w.shape
    out = torch.empty(B, n_fft // 2 + 1, 1 + N // hop_length, 2, device=w.device, dtype=torch.float32)
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().b200voc_stft_complex(_lib.ptr(w), B, N, n_fft, hop_length, _lib.ptr(out),
                                                    _lib.current_stream()), "stft_complex")
    return torch.view_as_complex(out)


def mel_spectrogram(wav: torch.Tensor, n_fft: int = 1024, hop_length: int = 256, n_mels: int = 80,
                    sample_rate: int = 22050, log: bool = False) -> torch.Tensor:
    """power mel [B, n_mels, frames] (log=True: log(clamp(mel, 1e-5))), fused STFT -> |X|^2 -> mel."""
    w = _prep(wav)
    B, N = w.shape
    out = torch.empty(B, n_mels, 1 + N // hop_length, device=w.device, dtype=torch.float32)
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().b200voc_stft_logmel(_lib.ptr(w), B, N, n_fft, hop_length, n_mels, sample_rate, int(log),
                                                   _lib.ptr(out), _lib.current_stream()), "stft_logmel")
    return out


def log_mel(wav: torch.Tensor, n_fft: int = 1024, hop_length: int = 256, n_mels: int = 80,
            sample_rate: int = 22050) -> torch.Tensor:
    return mel_spectrogram(wav, n_fft, hop_length, n_mels, sample_rate, log=True)


def istft(spec: torch.Tensor, n_fft: int, hop_length: int, length: int) -> torch.Tensor:
    """complex [B, n_fft/2+1, frames] -> wav [B, length] (torch.istft semantics, Hann, center)."""
    _lib.require_cuda(spec)
    if not spec.is_complex() or spec.dim() != 3 or spec.shape[1] != n_fft // 2 + 1:
        raise ValueError(f"spec must be complex [B,{n_fft // 2 + 1},frames], got {tuple(spec.shape)} {spec.dtype}")
    ri = torch.view_as_real(spec.detach().to(torch.complex64).contiguous())
    B, _, frames = spec.shape
    out = torch.empty(B, length, device=spec.device, dtype=torch.float32)
    with torch.cuda.device(spec.device):
        _lib.check(_lib.load().b200voc_istft(_lib.ptr(ri), B, frames, n_fft, hop_length, length, _lib.ptr(out),
                                             _lib.current_stream()), "istft")
    return out


class LearnableSTFT(nn.Module):
    """``LearnableSTFT(n_fft, hop_length)(wav[B,1,T]) -> [B, n_fft/2+1, frames]`` (stft.py:9-34):
    Hann STFT magnitude times a learnable per-bin gain (``filterbank`` ~ randn, same names/shapes
    so a reference state_dict loads unchanged)."""

    def __init__(self, n_fft: int, hop_length: int):
        super().__init__()
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.register_buffer("window", torch.hann_window(n_fft))
        self.filterbank = nn.Parameter(torch.randn(n_fft // 2 + 1))

    def forward(self, wav: torch.Tensor) -> torch.Tensor:
        # differentiable like the reference module (stft.py:22-34) w.r.t. the waveform and the per-bin gains
        if torch.is_grad_enabled() and (wav.requires_grad or self.filterbank.requires_grad):
            return _LearnableSTFTFn.apply(wav, self.filterbank, self.n_fft, self.hop_length)
        return stft_magnitude(wav, self.n_fft, self.hop_length, self.filterbank)


class _LearnableSTFTFn(torch.autograd.Function):
    """out = |STFT(wav)| * gain[:, None]; backward = b200voc_stft_mag_backward (complex STFT -> G * gain * X / |X| ->
    adjoint overlap-add -> reflect fold; gain gradient = sum G * |X|)."""

    @staticmethod
    def forward(ctx, wav, gain, n_fft, hop):
        w = _prep(wav)
        ctx.save_for_backward(w, gain.detach())
        ctx.meta = (int(n_fft), int(hop), tuple(wav.shape))
        return stft_magnitude(w, n_fft, hop, gain.detach())

    @staticmethod
    def backward(ctx, grad_out):
        w, gain = ctx.saved_tensors
        n_fft, hop, shape = ctx.meta
        B, N = w.shape
        lib = _lib.load()
        go = grad_out.detach().to(torch.float32).contiguous()
        gs = gain.to(torch.float32).contiguous()
        grad_wav = torch.zeros_like(w)
        grad_gain = torch.zeros(n_fft // 2 + 1, device=w.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(w.device):
            ws = torch.empty(int(lib.b200voc_stft_mag_backward_workspace_bytes(B, N, n_fft, hop)), dtype=torch.uint8,
                             device=w.device)
            _lib.check(lib.b200voc_stft_mag_backward(_lib.ptr(w), B, N, n_fft, hop, _lib.ptr(gs), _lib.ptr(go),
                                                     _lib.ptr(grad_wav), _lib.ptr(grad_gain), _lib.ptr(ws), ws.numel(),
                                                     _lib.current_stream()), "stft_mag_backward")
        return grad_wav.view(shape), grad_gain, None, None


def _stft_loss_values(f: torch.Tensor, r: torch.Tensor, gains, n_ffts, hop: int) -> torch.Tensor:
    """sum over resolutions of mean(|g| * ||X_f| - |X_r||) (fused STFT + L1 kernels)."""
    B, N = f.shape
    lib = _lib.load()
    sums = torch.zeros(len(n_ffts), device=f.device, dtype=torch.float64)
    loss = torch.zeros((), device=f.device, dtype=torch.float32)
    with torch.cuda.device(f.device):
        for i, (g, n_fft) in enumerate(zip(gains, n_ffts)):
            # L1(|X_f| g, |X_r| g) = mean(|g| * | |X_f| - |X_r| |): the kernel is given |g|
            ga = g.detach().abs().to(torch.float32).contiguous()
            _lib.check(lib.b200voc_stft_l1(_lib.ptr(f), _lib.ptr(r), B, N, n_fft, hop, _lib.ptr(ga),
                                           sums[i:i + 1].data_ptr(), _lib.current_stream()), "stft_l1")
            numel = B * (n_fft // 2 + 1) * (1 + N // hop)
            loss = loss + (sums[i] / numel).to(torch.float32)
    return loss


class _STFTLossFn(torch.autograd.Function):
    """loss = lambda * sum_res mean | |STFT(fake)| g - |STFT(real)| g |; gradients w.r.t. ``wav_fake`` and
    every gain vector come from b200voc_stft_l1_backward (STFT -> spectral gradient -> adjoint STFT as
    an un-normalised overlap-add -> reflect fold), accumulated over the resolutions in place."""

    @staticmethod
    def forward(ctx, wav_fake, wav_real, lam, hop, n_ffts, *gains):
        f, r = _prep(wav_fake), _prep(wav_real)
        ctx.save_for_backward(f, r, *[g.detach() for g in gains])
        ctx.meta = (float(lam), int(hop), tuple(int(n) for n in n_ffts), tuple(wav_fake.shape))
        return _stft_loss_values(f, r, gains, n_ffts, hop) * lam

    @staticmethod
    def backward(ctx, grad_out):
        f, r, *gains = ctx.saved_tensors
        lam, hop, n_ffts, shape = ctx.meta
        B, N = f.shape
        lib = _lib.load()
        grad_wav = torch.zeros_like(f)
        grad_gains = [torch.zeros(n // 2 + 1, device=f.device, dtype=torch.float32) for n in n_ffts]
        # the kernels are linear in `scale`: run them with the constant lambda and apply the upstream gradient as a
        # DEVICE scalar afterwards -- float(grad_out) would be a host synchronisation (and not graph-capturable)
        scale = lam
        with torch.cuda.device(f.device):
            need = max(int(lib.b200voc_stft_l1_backward_workspace_bytes(B, N, n, hop)) for n in n_ffts)
            ws = torch.empty(need, dtype=torch.uint8, device=f.device)
            for g, gg, n in zip(gains, grad_gains, n_ffts):
                gs = g.to(torch.float32).contiguous()
                _lib.check(lib.b200voc_stft_l1_backward(_lib.ptr(f), _lib.ptr(r), B, N, n, hop, _lib.ptr(gs), scale,
                                                        _lib.ptr(grad_wav), _lib.ptr(gg), _lib.ptr(ws), ws.numel(),
                                                        _lib.current_stream()), "stft_l1_backward")
        go = grad_out.detach().to(device=f.device, dtype=torch.float32)
        grad_wav.mul_(go)
        for gg in grad_gains:
            gg.mul_(go)
        return (grad_wav.view(shape), None, None, None, None, *grad_gains)


class STFTLoss(nn.Module):
    """Multi-resolution STFT L1 loss (stft.py:36-54).  Differentiable w.r.t. ``wav_fake`` and the
    per-resolution ``filterbank`` gains (``wav_real`` is data)."""

    def __init__(self, cfg: GANConfig):
        super().__init__()
        self.stfts = nn.ModuleList([LearnableSTFT(n_fft, cfg.hop_length) for n_fft in cfg.stft_sizes])
        self.lambda_stft = cfg.lambda_stft

    def forward(self, wav_fake: torch.Tensor, wav_real: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(wav_fake, wav_real)
        if wav_fake.shape != wav_real.shape:
            raise ValueError(f"shape mismatch {tuple(wav_fake.shape)} vs {tuple(wav_real.shape)}")
        _prep(wav_fake)                                  # shape validation
        n_ffts = [m.n_fft for m in self.stfts]
        hops = {m.hop_length for m in self.stfts}
        gains = [m.filterbank for m in self.stfts]
        needs_grad = torch.is_grad_enabled() and (wav_fake.requires_grad or any(g.requires_grad for g in gains))
        if needs_grad and len(hops) == 1:
            return _STFTLossFn.apply(wav_fake, wav_real, self.lambda_stft, hops.pop(), n_ffts, *gains)
        f, r = _prep(wav_fake), _prep(wav_real)
        loss = torch.zeros((), device=f.device, dtype=torch.float32)
        for m in self.stfts:
            loss = loss + _stft_loss_values(f, r, [m.filterbank], [m.n_fft], m.hop_length)
        return loss * self.lambda_stft
