"""GlobalStyleTokens: drop-in for ``vocoder7.gst.GlobalStyleTokens`` (vocoder7/gst.py:8-35), the step
right before the Generator in its only caller (vocoder7/trainer.py:73).  Same constructor, same
parameter names (``tokens``, ``attn_conv.0.*``, ``attn_conv.2.*``) so a reference checkpoint loads
unchanged; ``forward`` runs the CUDA kernels through the C ABI (b200voc_gst_forward).  No fallback."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .config import GANConfig


class GlobalStyleTokens(nn.Module):
    def __init__(self, cfg: GANConfig):
        super().__init__()
        self.cfg = cfg
        # construction order == vocoder7/gst.py:15-22 so a seeded default init reproduces the reference's
        self.tokens = nn.Parameter(torch.randn(cfg.num_style_tokens, cfg.style_dim))
        self.attn_conv = nn.Sequential(
            nn.Conv1d(cfg.channels, cfg.style_dim, kernel_size=3, padding=1),
            nn.ReLU(),
            nn.Conv1d(cfg.style_dim, cfg.num_style_tokens, kernel_size=1),
        )

    def forward(self, mel_ref: torch.Tensor, *, mel_layout: str = "BCT") -> torch.Tensor:
        """mel_ref[B,channels,T] (or [B,T,channels] with ``mel_layout="BTC"``) -> style[B,style_dim]
        (vocoder7/gst.py:24-35)."""
        if mel_layout not in ("BCT", "BTC"):
            raise ValueError(f"mel_layout must be 'BCT' or 'BTC', got {mel_layout!r}")
        _lib.require_cuda(mel_ref)
        ch_axis = 1 if mel_layout == "BCT" else 2
        if mel_ref.dim() != 3 or mel_ref.shape[ch_axis] != self.cfg.channels:
            raise ValueError(f"mel_ref has shape {tuple(mel_ref.shape)}; expected {self.cfg.channels} channels "
                             f"on axis {ch_axis}")
        B, T = mel_ref.shape[0], mel_ref.shape[3 - ch_axis]
        lib = _lib.load()
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        ws = [f32(self.attn_conv[0].weight), f32(self.attn_conv[0].bias), f32(self.attn_conv[2].weight),
              f32(self.attn_conv[2].bias), f32(self.tokens)]
        for w in ws:
            if not w.is_cuda:
                raise _lib.B200VocError("GlobalStyleTokens parameters must be on a CUDA device (call .to('cuda'))")
        with torch.cuda.device(mel_ref.device):
            _lib.check(lib.b200voc_device_supported(mel_ref.device.index or 0), "device check")
            mel = f32(mel_ref)
            nt, sd = self.cfg.num_style_tokens, self.cfg.style_dim
            scratch = torch.empty(max(1, int(lib.b200voc_gst_scratch_bytes(B, T, nt))), dtype=torch.uint8,
                                  device=mel.device)
            style = torch.empty(B, sd, device=mel.device, dtype=torch.float32)
            _lib.check(lib.b200voc_gst_forward(_lib.ptr(mel), int(mel_layout == "BTC"), B, T, self.cfg.channels, sd, nt,
                                               *[_lib.ptr(w) for w in ws], _lib.ptr(scratch), scratch.numel(),
                                               _lib.ptr(style), _lib.current_stream()), "gst_forward")
        return style
