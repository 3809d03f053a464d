"""Numerics emulation of the CUDA kernels' rounding points, on the CPU oracle.  TEST INFRASTRUCTURE.

Predicts (before a kernel exists, and afterwards as an independent cross-check) the waveform
error of a given operand/storage plan.  It follows the *kernel* data flow documented in
DESIGN.md section "Numerics plan":

  * every inter-layer activation is stored once in HBM as a 16-bit tensor (fp16 or bf16, per
    stage); in the wide stages (C >= 128), when the consumer is a residual block the stored value
    is leaky_relu(x) and the raw x is recovered exactly-invertibly (x = a if a >= 0 else 10*a); the
    narrow stages (C <= 64) store raw x, apply leaky_relu in 16-bit arithmetic on chip and add the
    residual x on the tensor core;
  * tensor-core operands (activations and weights) are rounded to the stage's 16-bit format,
    accumulation, bias, GLU, FiLM and the residual add are fp32;
  * conditioning MLPs, FiLM projections, band_split and band_merge + tanh are fp32 CUDA-core
    kernels (band_split/band_merge read/write the 16-bit activations).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from . import vocoder7_oracle as O

_DT = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}


def _q(x: torch.Tensor, fmt: str) -> torch.Tensor:
    return x if fmt == "fp32" else x.to(_DT[fmt]).to(torch.float32)


def _store_lrelu(x, fmt):
    """what the producing epilogue writes: leaky_relu(x) rounded to fmt; returns (a_stored, x_rec)."""
    a = _q(F.leaky_relu(x, O.LRELU_SLOPE), fmt)
    return a, torch.where(a >= 0, a, a * 10.0)


def _store_raw(x, fmt):
    """narrow stages: raw x rounded to fmt; the MMA operand is leaky_relu evaluated in the 16-bit
    format (slope constant and product rounded to fmt)."""
    xs = _q(x, fmt)
    if fmt == "fp32":
        return F.leaky_relu(xs, O.LRELU_SLOPE), xs
    slope = _q(torch.tensor(O.LRELU_SLOPE), fmt)
    return torch.maximum(xs, _q(xs * slope, fmt)), xs


def emulated_forward(sd: Dict[str, torch.Tensor], cfg: O.OracleConfig, mel, prosody, style, emotion,
                     stage_fmt: List[str], split_fmt: str = None, **kw) -> torch.Tensor:
    """stage_fmt[i] = 16-bit format of stage i's operands AND of the activations it writes."""
    cond = O.conditioning_forward(sd, prosody, style, emotion, kw.get("style_drop", False),
                                  kw.get("emo_drop", False), kw.get("w_style", 1.0), kw.get("w_emo", 1.0))
    B, C, T = mel.shape
    nb = cfg.num_bands
    band = C // nb
    outs = []
    split_fmt = split_fmt or stage_fmt[0]
    for b in range(nb):
        x = F.conv1d(mel[:, b * band:(b + 1) * band], sd[f"band_split.{b}.weight"], sd[f"band_split.{b}.bias"], padding=3)
        x = _q(x, split_fmt)                       # raw, consumer is ConvT
        for i, f in enumerate(cfg.upsample_factors):
            fmt = stage_fmt[i]
            p = f"upsample_blocks.{i}"
            x = F.conv_transpose1d(_q(x, fmt), _q(sd[f"{p}.0.weight"], fmt), sd[f"{p}.0.bias"], stride=f, padding=f // 2)
            narrow = x.shape[1] <= 64
            a, x = _store_raw(x, fmt) if narrow else _store_lrelu(x, fmt)
            nres = len(cfg.res_dilations)
            for j, d in enumerate(cfg.res_dilations):
                q = f"{p}.{j + 1}"
                h = F.conv1d(a, _q(sd[f"{q}.conv.weight"], fmt), sd[f"{q}.conv.bias"], dilation=d, padding=d)
                h = F.glu(h, dim=1)
                scale, shift = F.conv1d(cond, sd[f"{q}.film.weight"], sd[f"{q}.film.bias"]).chunk(2, dim=1)
                rep = x.shape[-1] // T
                h = h * (1.0 + scale.repeat_interleave(rep, -1)) + shift.repeat_interleave(rep, -1)
                y = x + F.conv1d(_q(h, fmt), _q(sd[f"{q}.proj.weight"], fmt), sd[f"{q}.proj.bias"])
                last = j == nres - 1
                if last:
                    nxt = stage_fmt[i + 1] if i + 1 < len(stage_fmt) else fmt
                    x = _q(y, nxt)                 # raw store for the next ConvT / band_merge
                else:
                    a, x = _store_raw(y, fmt) if narrow else _store_lrelu(y, fmt)
        outs.append(x)
    wav = F.conv1d(torch.cat(outs, 1), sd["band_merge.weight"], sd["band_merge.bias"], padding=3)
    return torch.tanh(wav)


PLANS = {
    "fp16": ["fp16"] * 4,
    "bf16": ["bf16"] * 4,
    "mixed": ["bf16", "fp16", "fp16", "fp16"],
    "fp32": ["fp32"] * 4,
}
