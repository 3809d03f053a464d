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
(``b200voc_spectral_norm_train``).  No CPU / PyTorch fallback.

Backward (the critic half of the training step, vocoder7/trainer.py:86-115): when autograd is recording and the waveform
or any parameter requires grad, every sub-discriminator stack runs as ONE ``torch.autograd.Function`` whose backward
chains the kernels of ``csrc/disc_bwd.cu`` layer by layer -- LeakyReLU / feature-gradient merge, bias gradient, wgrad,
spectral-norm backward (``u``, ``v`` constants of the graph as in torch) and dgrad down to the waveform -- with the
GEMM-shaped layers on the tensor cores (dgrad = the forward implicit GEMM with flipped weights, wgrad = one split-bf16
GEMM over positions).  ``B200VOC_DISC_BWD_TC=0`` keeps the backward on the fp32 CUDA-core kernels (A/B runs)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

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


def _bwd_tc_enabled() -> bool:
    """B200VOC_DISC_BWD_TC=0 keeps dgrad / wgrad on the fp32 CUDA-core kernels (A/B runs)."""
    import os
    return _tc_enabled() and os.environ.get("B200VOC_DISC_BWD_TC", "1") != "0"


def _avg_pool(x: torch.Tensor) -> torch.Tensor:
    """F.avg_pool1d(x, 4, 2, 1) of discriminators.py:99 on [B, 1, T]."""
    B, _, T = x.shape
    pooled = torch.empty(B, 1, (T - 2) // 2 + 1, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().b200voc_avg_pool1d_k4s2p1(_lib.ptr(x), B, T, _lib.ptr(pooled), _lib.current_stream()),
               "avg_pool1d")
    return pooled


class _AvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return _avg_pool(x)

    @staticmethod
    def backward(ctx, gy):
        B, _, T = ctx.shape
        gy = gy.to(torch.float32).contiguous()
        dx = torch.empty(ctx.shape, device=gy.device, dtype=torch.float32)
        with torch.cuda.device(gy.device):
            _lib.check(_lib.load().b200voc_avg_pool1d_k4s2p1_bwd(_lib.ptr(gy), B, T, _lib.ptr(dx), _lib.current_stream()),
                       "avg_pool1d_bwd")
        return dx


class _CriticStackFn(torch.autograd.Function):
    """One sub-discriminator (conv stack) as a single autograd node: inputs = the waveform (or pooled / chunked view of
    it) and (weight_orig, bias) of every layer, outputs = every returned map.  ``weight_u`` / ``weight_v`` are constants
    of the graph (torch.nn.utils.spectral_norm runs the power iteration under no_grad)."""

    @staticmethod
    def forward(ctx, owner, d, geom, x, *params):
        x_offset, B, Lin, P, in_batch_stride, in_valid, two_d = geom
        weights = owner._weights(d)
        maps = owner._walk(d, weights, _lib.ptr(x) + 4 * x_offset, x.device, B, Lin, P, in_batch_stride, in_valid, two_d)
        ctx.owner_specs = owner._specs[d]
        ctx.geom = geom
        ctx.x_shape = tuple(x.shape)
        # a later forward (D(real) after D(fake)) updates weight_u / weight_v in place in .train() mode: keep copies
        ctx.layers = [(w, sigma, u.clone() if owner.training else u, v.clone() if owner.training else v)
                      for (w, _, _, sigma, u, v) in weights]
        ctx.set_materialize_grads(False)          # maps the loss does not touch arrive as None, not as zero tensors
        ctx.save_for_backward(x, *maps)
        return tuple(maps)

    @staticmethod
    def backward(ctx, *gmaps):
        lib = _lib.load()
        x, *maps = ctx.saved_tensors
        x_offset, B, Lin, P, in_batch_stride, in_valid, two_d = ctx.geom
        specs = ctx.owner_specs
        dev = x.device
        stream = None
        n_layers = len(specs)
        grads: List[Optional[torch.Tensor]] = [None] * (2 * n_layers)
        dx_out = None
        if all(g is None for g in gmaps):
            return (None, None, None, None) + tuple(grads)
        gmaps = [None if g is None else g.to(torch.float32).contiguous() for g in gmaps]
        # index of every layer's maps in the output tuple, and its input length
        pos, lens, idx, L = [], [], 0, Lin
        for (cin, cout, k, st, pad, act) in specs:
            pos.append(idx)
            lens.append(L)
            idx += 2 if act else 1
            L = (L + 2 * pad - k) // st + 1
        use_tc = _bwd_tc_enabled()
        with torch.cuda.device(dev):
            stream = _lib.current_stream()
            sn_scratch = torch.empty(int(lib.b200voc_spectral_norm_bwd_scratch_bytes()) // 8, device=dev, dtype=torch.float64)
            g_next = None
            for l in range(n_layers - 1, -1, -1):
                cin, cout, k, st, pad, act = specs[l]
                w, sigma, u, v = ctx.layers[l]
                y_pre = maps[pos[l]]
                gy_pre = gmaps[pos[l]]
                gy_act = gmaps[pos[l] + 1] if act else None
                if not act:
                    g_next = None                     # the score layer has no activation
                if gy_pre is None and gy_act is None and g_next is None:
                    g_next = None
                    continue                          # nothing flows into this layer (nor, then, into the ones below)
                if gy_act is None and g_next is None:
                    g = gy_pre
                else:
                    g = torch.empty_like(y_pre)
                    _lib.check(lib.b200voc_disc_lrelu_bwd(_lib.ptr(y_pre), _lib.ptr(gy_pre), _lib.ptr(gy_act),
                                                          _lib.ptr(g_next), LRELU_SLOPE, y_pre.numel(), _lib.ptr(g), stream),
                               "disc_lrelu_bwd")
                Lin_l, Lout = lens[l], y_pre.shape[2]
                if l == 0:
                    xin_ptr, sb, valid = _lib.ptr(x) + 4 * x_offset, in_batch_stride, in_valid
                else:
                    xin_ptr, sb, valid = _lib.ptr(maps[pos[l - 1] + 1]), 0, 0
                # ---- bias gradient
                if ctx.needs_input_grad[4 + 2 * l + 1]:
                    db = torch.empty(cout, device=dev, dtype=torch.float32)
                    _lib.check(lib.b200voc_disc_bias_grad(_lib.ptr(g), B, cout, Lout * P, _lib.ptr(db), stream), "disc_bias_grad")
                    grads[2 * l + 1] = db
                # ---- weight gradient, then through the spectral norm
                if ctx.needs_input_grad[4 + 2 * l]:
                    dw = torch.empty_like(w)
                    if use_tc and l > 0 and lib.b200voc_disc_conv_wgrad_tc_supported(B, cin, cout, Lin_l, k, st, P, pad):
                        nb = int(lib.b200voc_disc_conv_wgrad_tc_workspace_bytes(B, cin, cout, Lin_l, k, pad))
                        ws = torch.empty(nb + 1024, device=dev, dtype=torch.uint8)
                        wp = (_lib.ptr(ws) + 1023) & ~1023
                        _lib.check(lib.b200voc_disc_conv_wgrad_tc(xin_ptr, _lib.ptr(g), B, cin, cout, Lin_l, k, pad, _lib.ptr(dw),
                                                                  wp, nb, stream), "disc_conv_wgrad_tc")
                    else:
                        nb = int(lib.b200voc_disc_conv_wgrad_scratch_bytes(B, cin, cout, Lin_l, P, k, st, pad))
                        scratch = torch.empty(max(nb // 4, 1), device=dev, dtype=torch.float32)
                        _lib.check(lib.b200voc_disc_conv_wgrad(xin_ptr, _lib.ptr(g), B, cin, cout, Lin_l, P, k, st, pad, sb, valid,
                                                               _lib.ptr(dw), _lib.ptr(scratch) if nb else 0, stream), "disc_conv_wgrad")
                    dw_orig = torch.empty_like(w)
                    rows, cols = int(w.shape[0]), w.numel() // int(w.shape[0])
                    _lib.check(lib.b200voc_spectral_norm_bwd(_lib.ptr(dw), _lib.ptr(w), _lib.ptr(u), _lib.ptr(v), _lib.ptr(sigma),
                                                             rows, cols, _lib.ptr(dw_orig), _lib.ptr(sn_scratch), stream),
                               "spectral_norm_bwd")
                    grads[2 * l] = dw_orig
                # ---- data gradient: the next (lower) layer's g_next, or the waveform's gradient
                if l == 0:
                    if ctx.needs_input_grad[3]:
                        dx_out = torch.zeros(ctx.x_shape, device=dev, dtype=torch.float32)
                        _lib.check(lib.b200voc_disc_conv_dgrad(_lib.ptr(g), _lib.ptr(w), B, cin, cout, Lin_l, P, k, st, pad, sb, valid,
                                                               0, _lib.ptr(dx_out) + 4 * x_offset, stream), "disc_conv_dgrad")
                    break
                g_next = torch.empty_like(maps[pos[l - 1] + 1])
                # tensor-core dgrad = the forward implicit GEMM on g with flipped weights; it writes 128 channels per tile,
                # so a 64-channel input (MSD's 64 -> 256 layer) runs with its flipped weight zero-padded to 128 rows
                cin_p = (cin + 127) // 128 * 128
                if use_tc and cin >= 64 and lib.b200voc_disc_conv_dgrad_tc_supported(cin_p, cout, k, st, P, pad):
                    wt = (torch.empty if cin_p == cin else torch.zeros)(cin_p, cout, k, device=dev, dtype=torch.float32)
                    _lib.check(lib.b200voc_disc_flip_weight(_lib.ptr(w), cout, cin, k, _lib.ptr(wt), stream), "disc_flip_weight")
                    wts = torch.empty(int(lib.b200voc_disc_split_weight_elems(cin_p, cout, k)), device=dev, dtype=torch.bfloat16)
                    _lib.check(lib.b200voc_disc_pack_weight_split(_lib.ptr(wt), cin_p, cout, k, _lib.ptr(wts), stream),
                               "disc_pack_weight_split")
                    zb = torch.zeros(cin_p, device=dev, dtype=torch.float32)
                    ws = torch.empty(int(lib.b200voc_disc_conv_tc_workspace_bytes(B, cout, Lout)), device=dev, dtype=torch.uint8)
                    dst = g_next if cin_p == cin else torch.empty(B, cin_p, Lin_l, device=dev, dtype=torch.float32)
                    _lib.check(lib.b200voc_disc_conv_tc(_lib.ptr(g), _lib.ptr(wts), _lib.ptr(zb), B, cout, cin_p, Lout, k, k - 1 - pad,
                                                        LRELU_SLOPE, _lib.ptr(dst), 0, _lib.ptr(ws), ws.numel(), stream),
                               "disc_conv_tc (dgrad)")
                    if cin_p != cin:
                        g_next.copy_(dst[:, :cin])
                else:
                    _lib.check(lib.b200voc_disc_conv_dgrad(_lib.ptr(g), _lib.ptr(w), B, cin, cout, Lin_l, P, k, st, pad, 0, 0, 0,
                                                           _lib.ptr(g_next), stream), "disc_conv_dgrad")
        return (None, None, None, dx_out) + tuple(grads)


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
            out.append((w, b, wsplit, sigma, u, v))
        if train:
            self._wcache.pop(d, None)        # u / v changed: an eval-mode forward must recompute
        else:
            self._wcache[d] = (key, out)
        return out

    # ---- one critic: walk its conv stack --------------------------------------------------------------
    def _walk(self, d: int, weights, x_ptr: int, device, B: int, Lin: int, P: int, in_batch_stride: int, in_valid: int,
              two_d: bool) -> List[torch.Tensor]:
        """The conv stack of sub-discriminator d: every conv map and every activation map, in order."""
        lib = _lib.load()
        maps: List[torch.Tensor] = []
        cur_ptr, cur_L, stride_b, valid = x_ptr, Lin, in_batch_stride, in_valid
        for (cin, cout, k, st, pad, act), (w, b, wsplit, _, _, _) in zip(self._specs[d], weights):
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
        return maps

    def _convs(self, d: int):
        return [m for m in self.discriminators[d] if not isinstance(m, nn.LeakyReLU)]

    def _needs_grad(self, x: torch.Tensor) -> bool:
        return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))

    def _run(self, d: int, x: torch.Tensor, x_offset: int, B: int, Lin: int, P: int, in_batch_stride: int, in_valid: int,
             two_d: bool):
        """Returns (score map, [every conv / activation map except the score]).  ``x`` is the fp32 tensor the first
        layer reads in place, starting ``x_offset`` elements into it."""
        geom = (x_offset, B, Lin, P, in_batch_stride, in_valid, two_d)
        if self._needs_grad(x):
            params = []
            for c in self._convs(d):
                params += [c.weight_orig, c.bias]
            maps = list(_CriticStackFn.apply(self, d, geom, x, *params))
        else:
            maps = self._walk(d, self._weights(d), _lib.ptr(x) + 4 * x_offset, x.device, B, Lin, P, in_batch_stride,
                              in_valid, two_d)
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
        return x.to(torch.float32).contiguous()          # differentiable: the critics' dgrad reaches the waveform


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
                out, feats = self._run(d, x, 0, B, rows, p, T, T, two_d=True)
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
                        pooled = _AvgPoolFn.apply(x) if self._needs_grad(x) and x.requires_grad else _avg_pool(x)
                    src, L = pooled, pooled.shape[2]
                out, feats = self._run(d, src, 0, B, L, 1, L, L, two_d=False)
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
                out, feats = self._run(d, x, start, B, L, 1, T, L, two_d=False)
                outputs.append(out)
                features.append(feats)
        return outputs, features
