"""Generator: drop-in for ``vocoder7.generator.Generator`` (vocoder7/generator.py:9-98).

Same constructor (``Generator(cfg)``), same ``forward`` signature, same sub-module names and
``state_dict`` key/shape layout -- so ``load_state_dict`` of a reference checkpoint works
unchanged -- but ``forward`` runs the hand-written sm_100a kernels through the C ABI
(include/b200voc.h).  PyTorch only owns the device memory and the stream.  There is no CPU or
eager-PyTorch fallback: CPU tensors or a missing ``libb200voc.so`` raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .config import GANConfig


class ResidualBlock(nn.Module):
    """Parameter container for the block generator.py:41 constructs as
    ``ResidualBlock(channels, dilation, cond_dim)`` (the reference never ships the class; layout
    per DESIGN.md D2): conv Conv1d(C,2C,3,dilation=d), film Conv1d(cond,2C,1), proj Conv1d(C,C,1)."""

    def __init__(self, channels: int, dilation: int, cond_dim: int):
        super().__init__()
        self.channels, self.dilation = channels, dilation
        self.conv = nn.Conv1d(channels, 2 * channels, kernel_size=3, dilation=dilation, padding=dilation)
        self.film = nn.Conv1d(cond_dim, 2 * channels, kernel_size=1)
        self.proj = nn.Conv1d(channels, channels, kernel_size=1)

    def forward(self, x, cond):  # pragma: no cover - the fused CUDA path never calls this
        raise _lib.B200VocError("ResidualBlock is evaluated inside Generator.forward (fused CUDA kernel)")


class SelfAttention(nn.Module):
    """Parameter container for ``SelfAttention(channels)`` (generator.py:44)."""

    def __init__(self, channels: int):
        super().__init__()
        self.channels = channels
        self.q = nn.Conv1d(channels, channels, 1)
        self.k = nn.Conv1d(channels, channels, 1)
        self.v = nn.Conv1d(channels, channels, 1)
        self.out = nn.Conv1d(channels, channels, 1)

    def forward(self, x):  # pragma: no cover
        raise _lib.B200VocError("SelfAttention is evaluated inside Generator.forward (fused CUDA kernel)")


class Generator(nn.Module):
    """BigVGAN-style multi-band waveform synthesiser, B200-native inference."""

    def __init__(self, cfg: GANConfig):
        super().__init__()
        self.cfg = cfg
        hidden = getattr(cfg, "hidden_dim", 512)
        band_size = cfg.channels // cfg.num_bands
        # construction order == generator.py:17-48 so a seeded default init reproduces the reference's
        self.band_split = nn.ModuleList(
            [nn.Conv1d(band_size, hidden, kernel_size=7, padding=3) for _ in range(cfg.num_bands)])
        self.cond_prosody = nn.Sequential(
            nn.Linear(18, cfg.cond_dim // 2), nn.SiLU(), nn.Linear(cfg.cond_dim // 2, cfg.cond_dim))
        self.style_proj = nn.Linear(cfg.style_dim, cfg.cond_dim)
        self.emotion_proj = nn.Linear(6, cfg.cond_dim)
        self.upsample_blocks = nn.ModuleList()
        ch = hidden
        for i, factor in enumerate(cfg.upsample_factors):
            block = nn.ModuleList()
            block.append(nn.ConvTranspose1d(ch, ch // 2, kernel_size=factor * 2, stride=factor, padding=factor // 2))
            for dilation in cfg.res_dilations:
                block.append(ResidualBlock(ch // 2, dilation, cfg.cond_dim))
            if i == len(cfg.upsample_factors) // 2:
                block.append(SelfAttention(ch // 2))
            self.upsample_blocks.append(block)
            ch //= 2
        self.band_merge = nn.Conv1d(ch * cfg.num_bands, 1, kernel_size=7, padding=3)
        self.hop = 1
        for f in cfg.upsample_factors:
            self.hop *= f
        self._handle: Optional[int] = None
        self._packed_key = None
        self._workspace: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ handle management
    def _c_config(self) -> _lib.GenConfig:
        cfg = self.cfg
        c = _lib.GenConfig()
        c.channels, c.cond_dim, c.style_dim, c.num_bands = cfg.channels, cfg.cond_dim, cfg.style_dim, cfg.num_bands
        c.n_stages = len(cfg.upsample_factors)
        for i, f in enumerate(cfg.upsample_factors):
            c.upsample_factors[i] = f
        c.n_dilations = len(cfg.res_dilations)
        for i, d in enumerate(cfg.res_dilations):
            c.res_dilations[i] = d
        c.hidden_dim = getattr(cfg, "hidden_dim", 512)
        c.use_attention = int(bool(getattr(cfg, "use_attention", True)))
        c.attn_window = int(getattr(cfg, "attn_window", None) or 0)
        plan = getattr(cfg, "precision", "fp16")
        if plan not in _lib.PLANS:
            raise ValueError(f"unknown precision plan {plan!r} (expected one of {sorted(_lib.PLANS)})")
        c.precision_plan = _lib.PLANS[plan]
        return c

    def _weights_key(self):
        return tuple((k, v.data_ptr(), v._version, v.device.index) for k, v in self.state_dict().items())

    def _ensure_packed(self, device: torch.device) -> int:
        """(re)pack the state_dict into kernel layouts when it changed (load_state_dict, .to())."""
        lib = _lib.load()
        key = self._weights_key()
        if self._handle is not None and key == self._packed_key:
            return self._handle
        if self._handle is None:
            _lib.check(lib.b200voc_device_supported(device.index or 0), "device check")
            h = C.c_void_p()
            cc = self._c_config()
            _lib.check(lib.b200voc_gen_create(C.byref(cc), C.byref(h)), "gen_create")
            self._handle = h.value
            if getattr(self, "_overflow_check", False):
                _lib.check(lib.b200voc_gen_set_overflow_check(self._handle, 1))
        sd = self.state_dict()
        n = lib.b200voc_gen_num_weights(self._handle)
        expected = {lib.b200voc_gen_weight_name(self._handle, i).decode(): lib.b200voc_gen_weight_numel(self._handle, i)
                    for i in range(n)}
        missing = sorted(set(expected) - set(sd))
        if missing:
            raise KeyError(f"state_dict is missing keys the kernel plan needs: {missing[:4]}...")
        stream = _lib.current_stream()
        keep = []
        for name, numel in expected.items():
            w = sd[name].detach()
            if not w.is_cuda:
                raise _lib.B200VocError("Generator parameters must be on a CUDA device (call .to('cuda'))")
            w = w.to(torch.float32).contiguous()
            keep.append(w)
            _lib.check(lib.b200voc_gen_set_weight(self._handle, name.encode(), _lib.ptr(w), w.numel(), stream),
                       f"set_weight({name})")
        _lib.check(lib.b200voc_gen_finalize(self._handle), "gen_finalize")
        torch.cuda.current_stream().synchronize()   # staging copies in `keep` may now be freed
        self._packed_key = key
        return self._handle

    def invalidate(self) -> None:
        """Force a re-pack of the kernel-layout weights on the next forward.  The pack is keyed on every parameter's
        (storage pointer, version counter); in-place updates made through ``.data`` (the reference trainer's EMA
        update, vocoder7/trainer.py:53-55, some checkpoint loaders) do NOT bump the version counter, so call this
        after such an update -- ``load_state_dict`` and ``.to()`` are detected without it."""
        self._packed_key = None

    repack = invalidate

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return out

    def __getstate__(self):
        # the native handle (a raw pointer), its pack key and the workspace belong to THIS object: a deepcopy / pickle
        # of the module (the usual EMA / eval copy idiom) gets none of them and packs its own on first use
        state = dict(self.__dict__)
        state["_handle"], state["_packed_key"], state["_workspace"] = None, None, None
        return state

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def set_overflow_check(self, enable: bool = True) -> None:
        """Debug aid: every forward then counts Inf / NaN values in each layer's stored 16-bit activations and raises
        ``B200VocOverflowError`` naming the first layer that left the storage format's range (fp16: 65504)."""
        self._overflow_check = bool(enable)
        if self._handle is not None:
            _lib.check(_lib.load().b200voc_gen_set_overflow_check(self._handle, int(self._overflow_check)))

    def workspace_bytes(self, B: int, T: int) -> int:
        return int(_lib.load().b200voc_gen_workspace_bytes(self._handle, B, T))

    def launch_count(self) -> int:
        return int(_lib.load().b200voc_gen_launch_count(self._handle)) if self._handle else 0

    def __del__(self):
        try:
            if self._handle is not None and _lib._lib is not None:
                _lib._lib.b200voc_gen_destroy(self._handle)
            self._handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ forward
    def forward(self, mel: torch.Tensor, prosody: torch.Tensor, style: torch.Tensor, emotion: torch.Tensor,
                style_drop: bool = False, emo_drop: bool = False, w_style: float = 1.0, w_emo: float = 1.0,
                *, out: Optional[torch.Tensor] = None, _tap: Optional[str] = None, mel_layout: str = "BCT",
                out_dtype: torch.dtype = torch.float32, frame_lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """mel[B,channels,T], prosody[B,T,18], style[B,style_dim], emotion[B,6] -> wav[B,1,hop*T]
        (generator.py:50-98).

        Keyword-only extras are the wire formats either side of the path (none changes the math):
        ``mel_layout="BTC"`` takes the mel as the refiner / acoustic model emit it, [B,T,channels]
        (sde_refiner5/model.py:304-306; vocoder7/trainer.py:77 transposes on the host -- here the
        transpose is folded into the first kernel's load); ``out_dtype=torch.int16`` writes 16-bit
        PCM ``round(clamp(wav,-1,1)*32767)``; ``frame_lengths[B]`` (valid mel frames of a padded
        batch, the collator's frame_length, batching2/colate.py:140-146) zeroes every sample at or
        past ``hop*frame_lengths[b]``."""
        _lib.require_cuda(mel, prosody, style, emotion)
        if mel_layout not in ("BCT", "BTC"):
            raise ValueError(f"mel_layout must be 'BCT' or 'BTC', got {mel_layout!r}")
        if out_dtype not in (torch.float32, torch.int16):
            raise ValueError(f"out_dtype must be torch.float32 or torch.int16, got {out_dtype}")
        ch_axis = 1 if mel_layout == "BCT" else 2
        if mel.dim() != 3 or mel.shape[ch_axis] != self.cfg.channels:
            want = f"[B,{self.cfg.channels},T]" if mel_layout == "BCT" else f"[B,T,{self.cfg.channels}]"
            raise ValueError(f"mel must be {want}, got {tuple(mel.shape)}")
        B, T = mel.shape[0], mel.shape[3 - ch_axis]
        if tuple(prosody.shape) != (B, T, 18):
            raise ValueError(f"prosody must be [B,T,18]=({B},{T},18), got {tuple(prosody.shape)}")
        if tuple(style.shape) != (B, self.cfg.style_dim):
            raise ValueError(f"style must be [B,{self.cfg.style_dim}], got {tuple(style.shape)}")
        if tuple(emotion.shape) != (B, 6):
            raise ValueError(f"emotion must be [B,6], got {tuple(emotion.shape)}")
        lib = _lib.load()
        with torch.cuda.device(mel.device):
            h = self._ensure_packed(mel.device)
            f32 = lambda t: t.detach().to(torch.float32).contiguous()
            mel_, pros_, sty_, emo_ = f32(mel), f32(prosody), f32(style), f32(emotion)
            if out is None:
                out = torch.empty(B, 1, self.hop * T, device=mel.device, dtype=out_dtype)
            elif out.dtype != out_dtype or out.numel() != B * self.hop * T or not out.is_contiguous():
                raise ValueError("out must be a contiguous tensor of B*hop*T elements of dtype out_dtype")
            io = _lib.GenIO()
            io.mel_time_major = int(mel_layout == "BTC")
            io.out_format = _lib.OUT_PCM16 if out_dtype == torch.int16 else _lib.OUT_F32
            valid = None
            if frame_lengths is not None:
                if tuple(frame_lengths.shape) != (B,):
                    raise ValueError(f"frame_lengths must be [B]=({B},), got {tuple(frame_lengths.shape)}")
                valid = (frame_lengths.to(device=mel.device, dtype=torch.int64) * self.hop).clamp_(0, self.hop * T)
                valid = valid.to(torch.int32).contiguous()
            io.valid_samples = _lib.ptr(valid)
            need = self.workspace_bytes(B, T)
            ws = self._workspace
            if ws is None or ws.numel() < need or ws.device != mel.device:
                ws = torch.empty(need, dtype=torch.uint8, device=mel.device)
                self._workspace = ws
            tap_out = None
            if _tap is not None:
                tap_out = torch.empty(self._tap_numel(_tap, B, T), device=mel.device, dtype=torch.float32)
            _lib.check(lib.b200voc_gen_forward_ex(
                h, _lib.ptr(mel_), _lib.ptr(pros_), _lib.ptr(sty_), _lib.ptr(emo_), B, T, int(style_drop),
                int(emo_drop), float(w_style), float(w_emo), C.byref(io), _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                _tap.encode() if _tap else None, _lib.ptr(tap_out), _lib.current_stream()), "gen_forward")
        if _tap is not None:
            return out, tap_out
        return out

    def _tap_numel(self, name: str, B: int, T: int) -> int:
        nb, H = self.cfg.num_bands, getattr(self.cfg, "hidden_dim", 512)
        if name == "cond":
            return B * T * self.cfg.cond_dim
        if name == "split":
            return B * nb * T * H
        ch, L = H, T
        for i, f in enumerate(self.cfg.upsample_factors):
            ch //= 2
            L *= f
            if name in (f"up{i}", "attn") or name.startswith(f"res{i}."):
                if name == "attn" and i != len(self.cfg.upsample_factors) // 2:
                    continue
                return B * nb * L * ch
        raise ValueError(f"unknown tap {name!r}")
