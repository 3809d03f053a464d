"""Discriminator forwards: drop-ins for ``vocoder7.discriminators`` (vocoder7/discriminators.py:8-157), the
critics the trainer runs on every generated and real waveform (vocoder7/trainer.py:86-92).  SURVEY.md
section 8(f) rank 4, forward half.

Same constructors, same ``state_dict`` keys (``discriminators.<i>.<j>.{weight_orig,weight_u,weight_v,bias}``
from ``torch.nn.utils.spectral_norm``) and the same return value ``(outputs, features)`` -- one score map per
sub-discriminator and the list of every intermediate conv / LeakyReLU map -- so a reference checkpoint loads
unchanged and ``compute_gan_loss`` (vocoder7/losses.py:8-52) consumes the result as is.  ``forward`` runs the
CUDA kernels of ``csrc/disc.cu`` / ``csrc/disc_gemm.cu`` through the C ABI; the torch modules below only hold
parameters.  Spectral normalisation follows the module's mode like ``torch.nn.utils.spectral_norm``: in ``.eval()``
sigma comes from the stored ``u``/``v`` (applied on the GPU once per parameter version), in ``.train()`` every forward
first runs one power iteration and updates the ``weight_u`` / ``weight_v`` buffers in place
(``b200voc_spectral_norm_train``).  No CPU / PyTorch fallback, no autograd (the backward kernels are the other half
of rank 4)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from .config import GANConfig

LRELU_SLOPE = 0.2   # nn.LeakyReLU(0.2), discriminators.py:26,83,132

# one conv layer: (Cin, Cout, K, stride, pad, followed_by_leaky_relu)
LayerSpec = Tuple[int, int, int, int, int, bool]


def _stack(n_hidden: int, k: int, strides: Sequence[int], two_d: bool) -> Tuple[nn.Sequential, List[LayerSpec]]:
    """n_hidden spectral-normalised convs (channels x4 each, LeakyReLU after each) and a final k3 conv to one
    channel -- the shape shared by all three critics.  Module order inside the Sequential (conv, act, conv, act,
    ..., conv) and RNG consumption order (conv init, then u, v of its spectral norm) equal the reference's."""
    mods: List[nn.Module] = []
    specs: List[LayerSpec] = []
    ch = 1
    plan = [(ch * 4 ** i, ch * 4 ** (i + 1), k, strides[i], k // 2, True) for i in range(n_hidden)]
    plan.append((ch * 4 ** n_hidden, 1, 3, 1, 1, False))
    for cin, cout, ks, st, pad, act in plan:
        if two_d:
            conv = nn.Conv2d(cin, cout, kernel_size=(ks, 1), stride=(st, 1), padding=(pad, 0))
        else:
            conv = nn.Conv1d(cin, cout, kernel_size=ks, stride=st, padding=pad)
        mods.append(nn.utils.spectral_norm(conv))
        if act:
            mods.append(nn.LeakyReLU(LRELU_SLOPE))
        specs.append((cin, cout, ks, st, pad, act))
    return nn.Sequential(*mods), specs


def _tc_enabled() -> bool:
    """B200VOC_DISC_TC=0 keeps every layer on the fp32 CUDA-core kernel (A/B runs)."""
    import os
    return os.environ.get("B200VOC_DISC_TC", "1") != "0"


class _CriticBase(nn.Module):
    """Parameter container + the layer walker shared by the three critics."""

    def __init__(self, cfg: GANConfig):
        super().__init__()
        self.cfg = cfg
        self.discriminators = nn.ModuleList()
        self._specs: List[List[LayerSpec]] = []
        self._wcache = {}

    def _add(self, n_hidden: int, k: int, strides: Sequence[int], two_d: bool) -> None:
        seq, specs = _stack(n_hidden, k, strides, two_d)
        self.discriminators.append(seq)
        self._specs.append(specs)

    def invalidate(self) -> None:
        """Drop the cached spectral-normalised weights (needed after in-place updates made through ``.data``, which
        do not bump the version counters the cache is keyed on)."""
        self._wcache = {}

    # ---- spectral-normalised weights, cached per parameter version ------------------------------------
    def _weights(self, d: int):
        convs = [m for m in self.discriminators[d] if not isinstance(m, nn.LeakyReLU)]
        key = tuple((c.weight_orig.data_ptr(), c.weight_orig._version, c.weight_u._version, c.weight_v._version,
                     c.bias.data_ptr(), c.bias._version) for c in convs)
        train = self.training            # .train(): one power iteration per forward, u / v updated in place, nothing cached
        hit = self._wcache.get(d)
        if not train and hit is not None and hit[0] == key:
            return hit[1]
        lib = _lib.load()
        out = []
        for c in convs:
            w0, u, v, b = c.weight_orig.detach(), c.weight_u.detach(), c.weight_v.detach(), c.bias.detach()
            _lib.require_cuda(w0, u, v, b)           # parameters must have been moved with .to('cuda')
            w0, b = w0.to(torch.float32).contiguous(), b.to(torch.float32).contiguous()
            rows, cols = w0.shape[0], w0.numel() // w0.shape[0]
            w = torch.empty_like(w0)
            sigma = torch.empty(1, device=w0.device, dtype=torch.float32)
            if train:
                # torch.nn.utils.spectral_norm, training mode (the reference trainer's every critic forward,
                # vocoder7/trainer.py:86-115): the module's weight_u / weight_v buffers ARE the kernel's in/out vectors
                if u.dtype != torch.float32 or v.dtype != torch.float32 or not (u.is_contiguous() and v.is_contiguous()):
                    raise _lib.B200VocError("training-mode spectral norm needs contiguous fp32 weight_u / weight_v buffers")
                scratch = torch.empty(rows + cols, device=w0.device, dtype=torch.float32)
                _lib.check(lib.b200voc_spectral_norm_train(_lib.ptr(w0), _lib.ptr(u), _lib.ptr(v), rows, cols, 1e-12,
                                                           _lib.ptr(w), _lib.ptr(sigma), _lib.ptr(scratch),
                                                           _lib.current_stream()), "spectral_norm_train")
            else:
                u, v = u.to(torch.float32).contiguous(), v.to(torch.float32).contiguous()
                _lib.check(lib.b200voc_spectral_norm_weight(_lib.ptr(w0), _lib.ptr(u), _lib.ptr(v), rows, cols,
                                                            _lib.ptr(w), _lib.ptr(sigma), _lib.current_stream()),
                           "spectral_norm_weight")
            # GEMM-shaped layers (stride 1, wide) run on the tensor cores with split-bf16 operands: pack [hi | lo] once
            wsplit = None
            cout, cin, k = int(w0.shape[0]), int(w0.shape[1]), int(w0.shape[2])
            st = c.stride[0]
            P1 = w0.dim() == 3
            if (_tc_enabled() and P1 and lib.b200voc_disc_conv_tc_supported(cin, cout, k, int(st), 1)):
                wsplit = torch.empty(int(lib.b200voc_disc_split_weight_elems(cout, cin, k)), device=w0.device,
                                     dtype=torch.bfloat16)
                _lib.check(lib.b200voc_disc_pack_weight_split(_lib.ptr(w), cout, cin, k, _lib.ptr(wsplit),
                                                              _lib.current_stream()), "disc_pack_weight_split")
            out.append((w, b, wsplit))
        if train:
            self._wcache.pop(d, None)        # u / v changed: an eval-mode forward must recompute
        else:
            self._wcache[d] = (key, out)
        return out

    # ---- one critic: walk its conv stack --------------------------------------------------------------
    def _run(self, d: int, x_ptr: int, device, B: int, Lin: int, P: int, in_batch_stride: int, in_valid: int,
             two_d: bool):
        """Returns (score map, [every conv / activation map except the score])."""
        lib = _lib.load()
        maps: List[torch.Tensor] = []
        cur_ptr, cur_L, stride_b, valid = x_ptr, Lin, in_batch_stride, in_valid
        for (cin, cout, k, st, pad, act), (w, b, wsplit) in zip(self._specs[d], self._weights(d)):
            Lout = int(lib.b200voc_disc_conv_out_len(cur_L, k, st, pad))
            if Lout <= 0:
                raise ValueError(f"discriminator input of {cur_L} samples is shorter than the kernel ({k})")
            shape = (B, cout, Lout, P) if two_d else (B, cout, Lout)
            y_pre = torch.empty(shape, device=device, dtype=torch.float32)
            y_act = torch.empty(shape, device=device, dtype=torch.float32) if act else None
            if wsplit is not None and P == 1 and stride_b == 0:
                ws = torch.empty(int(lib.b200voc_disc_conv_tc_workspace_bytes(B, cin, cur_L)), device=device,
                                 dtype=torch.uint8)
                _lib.check(lib.b200voc_disc_conv_tc(cur_ptr, _lib.ptr(wsplit), _lib.ptr(b), B, cin, cout, cur_L, k, pad,
                                                    LRELU_SLOPE, _lib.ptr(y_pre), _lib.ptr(y_act), _lib.ptr(ws),
                                                    ws.numel(), _lib.current_stream()), "disc_conv_tc")
            else:
                _lib.check(lib.b200voc_disc_conv(cur_ptr, _lib.ptr(w), _lib.ptr(b), B, cin, cout, cur_L, P, k, st, pad,
                                                 stride_b, valid, LRELU_SLOPE, _lib.ptr(y_pre), _lib.ptr(y_act),
                                                 _lib.current_stream()), "disc_conv")
            maps.append(y_pre)
            if act:
                maps.append(y_act)
                cur_ptr = _lib.ptr(y_act)
            cur_L, stride_b, valid = Lout, 0, 0     # later layers read contiguous maps
        return maps[-1], maps[:-1]

    def _inputs(self, T: int) -> List[Tuple[int, int]]:
        """(rows, columns) of the map each sub-discriminator reads for a T-sample waveform."""
        raise NotImplementedError

    def forward_flops(self, B: int, T: int) -> float:
        """Multiply-add FLOPs (2 per MAC) of one forward on [B, 1, T] (bench.py's accounting)."""
        total = 0.0
        for specs, (L, P) in zip(self._specs, self._inputs(T)):
            for cin, cout, k, st, pad, _ in specs:
                L = (L + 2 * pad - k) // st + 1
                total += 2.0 * B * cout * cin * k * L * P
        return total

    @staticmethod
    def _prep(x: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(x)
        if x.dim() != 3 or x.shape[1] != 1:
            raise ValueError(f"waveform must be [B, 1, T], got {tuple(x.shape)}")
        if x.shape[0] == 0 or x.shape[2] == 0:
            raise ValueError("empty waveform batch")
        _lib.check(_lib.load().b200voc_device_supported(x.device.index or 0), "device check")
        return x.detach().to(torch.float32).contiguous()


class MultiPeriodDiscriminator(_CriticBase):
    """vocoder7/discriminators.py:8-60: for every period p the waveform is zero-padded to a multiple of p, viewed
    as [B, 1, T/p, p] and run through 4x Conv2d((5,1), stride (3,1)) + LeakyReLU and a final Conv2d((3,1))."""

    def __init__(self, cfg: GANConfig):
        super().__init__(cfg)
        for _ in cfg.disc_periods:
            self._add(4, 5, [3, 3, 3, 3], two_d=True)

    def _inputs(self, T):
        return [((T + p - 1) // p, p) for p in self.cfg.disc_periods]

    def forward(self, x: torch.Tensor):
        x = self._prep(x)
        B, _, T = x.shape
        outputs, features = [], []
        with torch.cuda.device(x.device):
            for d, p in enumerate(self.cfg.disc_periods):
                rows = (T + p - 1) // p        # F.pad to a multiple of p == reads past T return zero
                out, feats = self._run(d, _lib.ptr(x), x.device, B, rows, p, T, T, two_d=True)
                outputs.append(out)
                features.append(feats)
        return outputs, features


class MultiScaleDiscriminator(_CriticBase):
    """vocoder7/discriminators.py:63-108: Conv1d stacks (k = 15, 41, 41; stride 2,2,2,1,1) on x and on
    avg_pool1d(x, 4, 2, 1).  As in the reference (discriminators.py:99) BOTH pooled scales are pooled from x."""

    def __init__(self, cfg: GANConfig):
        super().__init__(cfg)
        for ks in cfg.disc_kernel_sizes:
            self._add(5, ks, [2, 2, 2, 1, 1], two_d=False)

    def _inputs(self, T):
        return [(T if d == 0 else (T - 2) // 2 + 1, 1) for d in range(len(self.discriminators))]

    def forward(self, x: torch.Tensor):
        x = self._prep(x)
        B, _, T = x.shape
        lib = _lib.load()
        outputs, features = [], []
        with torch.cuda.device(x.device):
            pooled = None
            for d in range(len(self.discriminators)):
                if d == 0:
                    src, L = x, T
                else:
                    if pooled is None:
                        if T < 2:
                            raise ValueError("waveform too short for avg_pool1d(4, 2, 1)")
                        pooled = torch.empty(B, 1, (T - 2) // 2 + 1, device=x.device, dtype=torch.float32)
                        _lib.check(lib.b200voc_avg_pool1d_k4s2p1(_lib.ptr(x), B, T, _lib.ptr(pooled),
                                                                 _lib.current_stream()), "avg_pool1d")
                    src, L = pooled, pooled.shape[2]
                out, feats = self._run(d, _lib.ptr(src), x.device, B, L, 1, L, L, two_d=False)
                outputs.append(out)
                features.append(feats)
        return outputs, features


class MultiBandDiscriminator(_CriticBase):
    """vocoder7/discriminators.py:111-157: one Conv1d stack (k15, stride 2, x4) per ``torch.chunk`` of the TIME
    axis (discriminators.py:147 -- the "bands" are consecutive quarters of the waveform)."""

    def __init__(self, cfg: GANConfig):
        super().__init__(cfg)
        for _ in range(cfg.num_bands):
            self._add(4, 15, [2, 2, 2, 2], two_d=False)

    def _inputs(self, T):
        size = (T + self.cfg.num_bands - 1) // self.cfg.num_bands
        return [(min(size, T - d * size), 1) for d in range(self.cfg.num_bands) if d * size < T]

    def forward(self, x: torch.Tensor):
        x = self._prep(x)
        B, _, T = x.shape
        nb = self.cfg.num_bands
        size = (T + nb - 1) // nb                      # torch.chunk: ceil(T / chunks) per chunk, last one shorter
        outputs, features = [], []
        with torch.cuda.device(x.device):
            for d in range(nb):
                start = d * size
                if start >= T:                         # torch.chunk returned fewer chunks; zip() stops there
                    break
                L = min(size, T - start)
                out, feats = self._run(d, _lib.ptr(x) + 4 * start, x.device, B, L, 1, T, L, two_d=False)
                outputs.append(out)
                features.append(feats)
        return outputs, features
