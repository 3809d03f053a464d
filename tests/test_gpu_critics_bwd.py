"""GPU parity of the critic BACKWARD -- SURVEY.md section 8(f) rank 4, the discriminator half of the training step
(vocoder7/trainer.py:86-115: d_loss.backward() / g_loss.backward() run autograd through vocoder7/discriminators.py) --
through the C ABI (csrc/disc_bwd.cu) against torch autograd: per layer in fp64, per critic against the CPU oracle.

Tolerance: fp32 sums of up to B*L = 1e4..1e5 products in a different order than the reference (and, on the tensor-core
paths, operands with 16 significant bits): every gradient tensor must agree to 1e-3 of its own max-abs -- three orders
of magnitude below what an indexing or sign error produces."""
import pytest
import torch
import torch.nn.functional as F

from oracle import vocoder7_oracle as O

pytestmark = pytest.mark.gpu
KINDS = ("mpd", "msd", "mbd")
TOL = 1e-3
G = 2048          # sentinel elements on both sides of every output buffer


def _lib():
    from b200voc import _lib as L
    return L, L.load()


def _guarded(n, dtype=torch.float32):
    big = torch.full((n + 2 * G,), -777.0, device="cuda", dtype=dtype)
    return big, big[G:G + n]


def _guards_intact(big, n):
    return bool((big[:G] == -777.0).all()) and bool((big[G + n:] == -777.0).all())


def _close(got, ref, what, tol=TOL):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    assert tuple(got.shape) == tuple(ref.shape), what
    err, scale = float((got - ref).abs().max()), float(ref.abs().max())
    assert err <= tol * scale + 1e-9, f"{what}: max-abs error {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("B,Cin,Cout,L,P,K,st,pad", [
    (2, 1, 4, 101, 1, 15, 2, 7),       # first MSD / MBD layer
    (3, 4, 16, 57, 3, 5, 3, 2),        # MPD: Conv2d (5,1) stride (3,1), period columns
    (2, 16, 64, 64, 1, 41, 2, 20),     # MSD k41 stride 2
    (1, 64, 256, 130, 1, 15, 1, 7),    # stride 1 (CUDA-core form of a tensor-core layer)
    (2, 256, 1, 33, 1, 3, 1, 1),       # score layer
    (1, 4, 16, 2, 7, 5, 3, 2),         # single output row
    (2, 64, 4, 50, 3, 5, 3, 2),        # few output channels, many input channels (small-output wgrad kernel, columns)
    (3, 128, 1, 7, 1, 1, 1, 0),        # pointwise score layer
])
def test_dgrad_wgrad_bias_layer_matches_fp64_autograd(B, Cin, Cout, L, P, K, st, pad):
    L_, lib = _lib()
    g_ = torch.Generator().manual_seed(B * 1000 + Cin + L + K)
    two_d = P > 1
    x = torch.randn(B, Cin, L, P, generator=g_, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(Cout, Cin, K, generator=g_, dtype=torch.float64) * 3.0).requires_grad_(True)
    b = torch.randn(Cout, generator=g_, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w.unsqueeze(-1), b, stride=(st, 1), padding=(pad, 0))
    gy = torch.randn(y.shape, generator=g_, dtype=torch.float64)
    y.backward(gy)
    Lout = y.shape[2]
    assert Lout == lib.b200voc_disc_conv_out_len(L, K, st, pad)
    xd, wd, gd = x.detach().float().cuda().contiguous(), w.detach().float().cuda(), gy.float().cuda().contiguous()
    s = L_.current_stream()
    n_x, n_w = B * Cin * L * P, Cout * Cin * K
    dx_big, dx = _guarded(n_x)
    dw_big, dw = _guarded(n_w)
    db_big, db = _guarded(Cout)
    dx.fill_(float("nan")); dw.fill_(float("nan")); db.fill_(float("nan"))
    L_.check(lib.b200voc_disc_conv_dgrad(L_.ptr(gd), L_.ptr(wd), B, Cin, Cout, L, P, K, st, pad, 0, 0, 0, L_.ptr(dx), s))
    nb = int(lib.b200voc_disc_conv_wgrad_scratch_bytes(B, Cin, Cout, L, P, K, st, pad))
    sc_big, sc = _guarded(max(nb // 4, 1))
    L_.check(lib.b200voc_disc_conv_wgrad(L_.ptr(xd), L_.ptr(gd), B, Cin, Cout, L, P, K, st, pad, 0, 0, L_.ptr(dw),
                                         L_.ptr(sc) if nb else 0, s))
    L_.check(lib.b200voc_disc_bias_grad(L_.ptr(gd), B, Cout, Lout * P, L_.ptr(db), s))
    torch.cuda.synchronize()
    for big, n in ((dx_big, n_x), (dw_big, n_w), (db_big, Cout), (sc_big, max(nb // 4, 1))):
        assert _guards_intact(big, n), "write outside the output buffer"
    _close(dx.view(B, Cin, L, P), x.grad, "dgrad")
    _close(dw.view(Cout, Cin, K), w.grad, "wgrad")
    _close(db, b.grad, "bias gradient")
    # accumulate: dx += dgrad
    L_.check(lib.b200voc_disc_conv_dgrad(L_.ptr(gd), L_.ptr(wd), B, Cin, Cout, L, P, K, st, pad, 0, 0, 1, L_.ptr(dx), s))
    torch.cuda.synchronize()
    _close(dx.view(B, Cin, L, P), 2 * x.grad, "dgrad (accumulate)")
    assert two_d or P == 1


def test_dgrad_wgrad_padded_period_view_and_time_chunks():
    """The first layer's input is read in place (MPD: [B, 1, rows, p] over a waveform whose tail is F.pad's zeros,
    in_valid = T; MBD: a time chunk, batch stride = T): wgrad must read, and dgrad write, exactly those elements."""
    L_, lib = _lib()
    s = L_.current_stream()
    g_ = torch.Generator().manual_seed(3)
    B, T, p, K, st, pad, Cout = 2, 103, 7, 5, 3, 2, 4
    rows = -(-T // p)
    xw = torch.randn(B, 1, T, generator=g_, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, 1, K, generator=g_, dtype=torch.float64, requires_grad=True)
    xp = F.pad(xw, (0, rows * p - T)).view(B, 1, rows, p)
    y = F.conv2d(xp, w.unsqueeze(-1), None, stride=(st, 1), padding=(pad, 0))
    gy = torch.randn(y.shape, generator=g_, dtype=torch.float64)
    y.backward(gy)
    xd, wd, gd = xw.detach().float().cuda(), w.detach().float().cuda(), gy.float().cuda().contiguous()
    dx_big, dx = _guarded(B * T)
    dx.fill_(0.0)
    dw = torch.empty(Cout, 1, K, device="cuda")
    L_.check(lib.b200voc_disc_conv_dgrad(L_.ptr(gd), L_.ptr(wd), B, 1, Cout, rows, p, K, st, pad, T, T, 0, L_.ptr(dx), s))
    nb = int(lib.b200voc_disc_conv_wgrad_scratch_bytes(B, 1, Cout, rows, p, K, st, pad))
    sc = torch.empty(max(nb // 4, 1), device="cuda")
    L_.check(lib.b200voc_disc_conv_wgrad(L_.ptr(xd), L_.ptr(gd), B, 1, Cout, rows, p, K, st, pad, T, T, L_.ptr(dw),
                                         L_.ptr(sc) if nb else 0, s))
    torch.cuda.synchronize()
    assert _guards_intact(dx_big, B * T)
    _close(dx.view(B, 1, T), xw.grad, "MPD dgrad into the unpadded waveform")
    _close(dw, w.grad, "MPD wgrad")
    # MBD: second of four chunks
    xw.grad = None; w.grad = None
    K, st, pad = 15, 2, 7
    w = torch.randn(Cout, 1, K, generator=g_, dtype=torch.float64, requires_grad=True)
    size = -(-T // 4)
    chunk = xw[:, :, size:2 * size]
    y = F.conv1d(chunk, w, None, stride=st, padding=pad)
    gy = torch.randn(y.shape, generator=g_, dtype=torch.float64)
    y.backward(gy)
    gd, wd = gy.float().cuda().contiguous(), w.detach().float().cuda()
    dx.fill_(0.0)
    L_.check(lib.b200voc_disc_conv_dgrad(L_.ptr(gd), L_.ptr(wd), B, 1, Cout, size, 1, K, st, pad, T, size, 0,
                                         L_.ptr(dx) + 4 * size, s))
    dw = torch.empty(Cout, 1, K, device="cuda")
    nb = int(lib.b200voc_disc_conv_wgrad_scratch_bytes(B, 1, Cout, size, 1, K, st, pad))
    sc = torch.empty(max(nb // 4, 1), device="cuda")
    L_.check(lib.b200voc_disc_conv_wgrad(L_.ptr(xd) + 4 * size, L_.ptr(gd), B, 1, Cout, size, 1, K, st, pad, T, size,
                                         L_.ptr(dw), L_.ptr(sc) if nb else 0, s))
    torch.cuda.synchronize()
    assert _guards_intact(dx_big, B * T)
    _close(dx.view(B, 1, T), xw.grad, "MBD dgrad into its chunk of the waveform")
    _close(dw, w.grad, "MBD wgrad")


@pytest.mark.parametrize("B,Cin,Cout,L,K", [(2, 64, 256, 300, 15), (1, 256, 128, 130, 41), (3, 64, 128, 65, 15),
                                            (2, 256, 1024, 345, 41), (1, 128, 64, 64, 5)])
def test_wgrad_tensor_core_matches_fp64(B, Cin, Cout, L, K):
    """b200voc_disc_conv_wgrad_tc (one split-bf16 tcgen05 GEMM over positions) against fp64 autograd and against the fp32
    CUDA-core kernel: ragged lengths, both kernel sizes, N = Cin*K that is not a multiple of the 128-column tile
    (64*41, 128*5), M = Cout below / above one tile, gradients at 1e-4 scale (bf16 range, not fp16)."""
    L_, lib = _lib()
    s = L_.current_stream()
    pad = K // 2
    assert lib.b200voc_disc_conv_wgrad_tc_supported(B, Cin, Cout, L, K, 1, 1, pad) == 1
    assert lib.b200voc_disc_conv_wgrad_tc_supported(B, Cin, Cout, L, K, 2, 1, pad) == 0
    assert lib.b200voc_disc_conv_wgrad_tc_supported(B, 16, Cout, L, K, 1, 1, pad) == 0
    g_ = torch.Generator().manual_seed(Cin + Cout + L)
    x = torch.randn(B, Cin, L, generator=g_, dtype=torch.float64) * 2.0
    gy = torch.randn(B, Cout, L, generator=g_, dtype=torch.float64) * 1e-4
    w = torch.zeros(Cout, Cin, K, dtype=torch.float64, requires_grad=True)
    F.conv1d(x, w, None, padding=pad).backward(gy)
    xd, gd = x.float().cuda(), gy.float().cuda()
    nb = int(lib.b200voc_disc_conv_wgrad_tc_workspace_bytes(B, Cin, Cout, L, K, pad))
    ws_big = torch.full((nb + 2048 + 1024,), 0xA5, device="cuda", dtype=torch.uint8)
    base = (ws_big.data_ptr() + 1024 + 1023) & ~1023
    off = base - ws_big.data_ptr()
    n_w = Cout * Cin * K
    dw_big, dw = _guarded(n_w)
    dw.fill_(float("nan"))
    L_.check(lib.b200voc_disc_conv_wgrad_tc(L_.ptr(xd), L_.ptr(gd), B, Cin, Cout, L, K, pad, L_.ptr(dw), base, nb, s))
    torch.cuda.synchronize()
    assert _guards_intact(dw_big, n_w), "write outside dw"
    assert bool((ws_big[:off] == 0xA5).all()) and bool((ws_big[off + nb:] == 0xA5).all()), "write outside the workspace"
    _close(dw.view(Cout, Cin, K), w.grad, "tensor-core wgrad", tol=1e-4)
    dw2 = torch.empty(n_w, device="cuda")
    nb2 = int(lib.b200voc_disc_conv_wgrad_scratch_bytes(B, Cin, Cout, L, 1, K, 1, pad))
    sc = torch.empty(max(nb2 // 4, 1), device="cuda")
    L_.check(lib.b200voc_disc_conv_wgrad(L_.ptr(xd), L_.ptr(gd), B, Cin, Cout, L, 1, K, 1, pad, 0, 0, L_.ptr(dw2),
                                         L_.ptr(sc) if nb2 else 0, s))
    torch.cuda.synchronize()
    _close(dw2.view(Cout, Cin, K), w.grad, "CUDA-core wgrad", tol=1e-4)
    assert lib.b200voc_disc_conv_wgrad_tc(L_.ptr(xd), L_.ptr(gd), B, Cin, Cout, L, K, pad, L_.ptr(dw), base, 16, s) == L_.ERR_BAD_ARG


def test_wgrad_tensor_core_batch_chunks(monkeypatch):
    """A batch whose packed operands exceed the cap is processed in chunks of batch items whose partial dW are added: the
    cap lowered to 2 MB makes B = 5 run as five chunks; the result must equal the single-pass one to fp32 rounding."""
    L_, lib = _lib()
    s = L_.current_stream()
    B, Cin, Cout, L, K, pad = 5, 64, 128, 200, 15, 7
    g_ = torch.Generator().manual_seed(1)
    xd = (torch.randn(B, Cin, L, generator=g_) * 2.0).cuda()
    gd = (torch.randn(B, Cout, L, generator=g_) * 1e-2).cuda()

    def run():
        nb = int(lib.b200voc_disc_conv_wgrad_tc_workspace_bytes(B, Cin, Cout, L, K, pad))
        ws = torch.empty(nb + 1024, device="cuda", dtype=torch.uint8)
        base = (ws.data_ptr() + 1023) & ~1023
        dw = torch.full((Cout, Cin, K), float("nan"), device="cuda")
        L_.check(lib.b200voc_disc_conv_wgrad_tc(L_.ptr(xd), L_.ptr(gd), B, Cin, Cout, L, K, pad, L_.ptr(dw), base, nb, s))
        torch.cuda.synchronize()
        return dw, nb
    one, nb_one = run()
    monkeypatch.setenv("B200VOC_WGRAD_TC_CAP_MB", "2")      # one item packs to 1.67 MB
    many, nb_many = run()
    assert nb_many < nb_one
    w = torch.zeros(Cout, Cin, K, dtype=torch.float64, requires_grad=True)
    F.conv1d(xd.cpu().double(), w, None, padding=pad).backward(gd.cpu().double())
    _close(one, w.grad, "single pass", tol=1e-4)
    _close(many, w.grad, "five chunks", tol=1e-4)


@pytest.mark.parametrize("B,Cin,Cout,L,K", [(2, 128, 256, 300, 15), (1, 256, 1024, 130, 41), (2, 256, 64, 129, 15)])
def test_dgrad_tensor_core_matches_fp64(B, Cin, Cout, L, K):
    """dgrad of a stride-1 layer = the forward tcgen05 implicit GEMM on g with flipped weights (b200voc_disc_flip_weight ->
    b200voc_disc_pack_weight_split -> b200voc_disc_conv_tc), against fp64 autograd."""
    L_, lib = _lib()
    s = L_.current_stream()
    pad = K // 2
    assert lib.b200voc_disc_conv_dgrad_tc_supported(Cin, Cout, K, 1, 1, pad) == 1
    assert lib.b200voc_disc_conv_dgrad_tc_supported(64, Cout, K, 1, 1, pad) == 0      # dgrad writes Cin channels: 128 per tile
    g_ = torch.Generator().manual_seed(Cin + Cout + L)
    x = torch.zeros(B, Cin, L, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin, K, generator=g_, dtype=torch.float64) * 30.0 / (Cin * K) ** 0.5
    gy = torch.randn(B, Cout, L, generator=g_, dtype=torch.float64) * 1e-3
    F.conv1d(x, w, None, padding=pad).backward(gy)
    wd, gd = w.float().cuda(), gy.float().cuda()
    wt = torch.empty(Cin, Cout, K, device="cuda")
    L_.check(lib.b200voc_disc_flip_weight(L_.ptr(wd), Cout, Cin, K, L_.ptr(wt), s))
    torch.cuda.synchronize()
    assert torch.equal(wt.cpu(), w.float().flip(2).transpose(0, 1).contiguous())
    wts = torch.empty(int(lib.b200voc_disc_split_weight_elems(Cin, Cout, K)), device="cuda", dtype=torch.bfloat16)
    L_.check(lib.b200voc_disc_pack_weight_split(L_.ptr(wt), Cin, Cout, K, L_.ptr(wts), s))
    zb = torch.zeros(Cin, device="cuda")
    ws = torch.empty(int(lib.b200voc_disc_conv_tc_workspace_bytes(B, Cout, L)), device="cuda", dtype=torch.uint8)
    dx_big, dx = _guarded(B * Cin * L)
    dx.fill_(float("nan"))
    L_.check(lib.b200voc_disc_conv_tc(L_.ptr(gd), L_.ptr(wts), L_.ptr(zb), B, Cout, Cin, L, K, K - 1 - pad, 0.2, L_.ptr(dx), 0,
                                      L_.ptr(ws), ws.numel(), s))
    torch.cuda.synchronize()
    assert _guards_intact(dx_big, B * Cin * L)
    # K = Cout * taps reaches 42 k products x 3 split terms here: the tensor core's fp32 accumulator drifts by ~1e-4 of the
    # result over that many sequential additions (measured 1.35e-4 at 1024 x 41; 3e-5 at the forward's 256 x 41)
    _close(dx.view(B, Cin, L), x.grad, "tensor-core dgrad", tol=4e-4)


def test_small_backward_entry_points():
    """LeakyReLU / feature-gradient merge (every combination of absent sources), avg_pool1d(4, 2, 1) backward on odd and
    even lengths, spectral-norm backward against autograd of W / (u . W v)."""
    L_, lib = _lib()
    s = L_.current_stream()
    g_ = torch.Generator().manual_seed(0)
    n = 5003
    y = torch.randn(n, generator=g_)
    gy, ga, gn = (torch.randn(n, generator=g_) for _ in range(3))
    yd, gyd, gad, gnd = y.cuda(), gy.cuda(), ga.cuda(), gn.cuda()
    mask = torch.where(y > 0, torch.tensor(1.0), torch.tensor(0.2))
    for use in ((1, 1, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1)):
        out = torch.full((n,), float("nan"), device="cuda")
        L_.check(lib.b200voc_disc_lrelu_bwd(L_.ptr(yd), L_.ptr(gyd) if use[0] else 0, L_.ptr(gad) if use[1] else 0,
                                            L_.ptr(gnd) if use[2] else 0, 0.2, n, L_.ptr(out), s))
        ref = use[0] * gy + (use[1] * ga + use[2] * gn) * mask
        assert float((out.cpu() - ref).abs().max()) <= 1e-6
    for Lw in (2, 3, 7, 8, 2403):
        x = torch.randn(3, 1, Lw, generator=g_, requires_grad=True)
        yp = F.avg_pool1d(x, 4, 2, 1)
        gyp = torch.randn(yp.shape, generator=g_)
        yp.backward(gyp)
        dx = torch.full((3, 1, Lw), float("nan"), device="cuda")
        gypd = gyp.cuda()
        L_.check(lib.b200voc_avg_pool1d_k4s2p1_bwd(L_.ptr(gypd), 3, Lw, L_.ptr(dx), s))
        assert float((dx.cpu() - x.grad).abs().max()) <= 1e-6, Lw
    for rows, cols in ((4, 15), (64, 16 * 41), (1024, 256 * 9)):
        w0 = torch.randn(rows, cols, generator=g_, dtype=torch.float64, requires_grad=True)
        u = F.normalize(torch.randn(rows, generator=g_, dtype=torch.float64), dim=0)
        v = F.normalize(torch.mv(w0.detach().t(), u), dim=0)
        u = F.normalize(torch.mv(w0.detach(), v), dim=0)
        sigma = torch.dot(u, torch.mv(w0, v))
        wn = w0 / sigma
        dw = torch.randn(rows, cols, generator=g_, dtype=torch.float64)
        wn.backward(dw)
        out = torch.full((rows, cols), float("nan"), device="cuda")
        scr = torch.empty(int(lib.b200voc_spectral_norm_bwd_scratch_bytes()) // 8, device="cuda", dtype=torch.float64)
        dwd, wnd, ud, vd = dw.float().cuda(), wn.detach().float().cuda(), u.float().cuda(), v.float().cuda()
        sd_ = sigma.detach().float().reshape(1).cuda()
        L_.check(lib.b200voc_spectral_norm_bwd(L_.ptr(dwd), L_.ptr(wnd), L_.ptr(ud), L_.ptr(vd), L_.ptr(sd_), rows, cols,
                                               L_.ptr(out), L_.ptr(scr), s))
        _close(out, w0.grad, f"spectral-norm backward {rows}x{cols}", tol=1e-4)
    assert lib.b200voc_disc_lrelu_bwd(0, 0, L_.ptr(gad), 0, 0.2, n, L_.ptr(gad), s) == L_.ERR_BAD_ARG


def _loss(outs, feats):
    """Touches every returned map like compute_gan_loss (vocoder7/losses.py:8-52): least-squares score terms and
    mean-abs feature terms."""
    loss = 0.0
    for o in outs:
        loss = loss + ((o - 1.0) ** 2).mean()
    for fs in feats:
        for j, f in enumerate(fs):
            loss = loss + (0.5 + 0.1 * j) * f.abs().mean()
    return loss


def _host(kind, cfg, seed):
    import b200voc
    cls = {"mpd": b200voc.MultiPeriodDiscriminator, "msd": b200voc.MultiScaleDiscriminator,
           "mbd": b200voc.MultiBandDiscriminator}[kind]
    torch.manual_seed(seed)
    return cls(cfg).cuda()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("training", [False, True])
def test_critic_backward_matches_oracle_autograd(kind, training):
    """loss.backward() through the module (default config: MSD's 256 -> 1024 k41 layers take the tensor-core dgrad /
    wgrad) against torch autograd over the CPU oracle: d loss / d weight_orig, d bias of every layer and d loss / d
    waveform; in .train() the power iteration runs first and u / v are constants of the graph, as in torch."""
    import b200voc
    cfg, ocfg = b200voc.GANConfig(), O.OracleConfig()
    mod = _host(kind, cfg, seed=1234)
    # Let the power iteration converge first (sigma -> the spectral norm, what a critic looks like a few steps into
    # training).  With the FRESH u, v of the initialiser sigma = u . (W v) is a nearly cancelling sum of order 1e-3 and the
    # weights W / sigma are huge: the score bias gradient of MSD (a cancelling mean) is then off by 2e-2 whichever kernels run.
    mod.train()
    with torch.no_grad():
        for _ in range(6):
            mod(torch.rand(1, 1, 600, device="cuda"))
    mod.train(training)
    sd = {k: v.detach().cpu().double() for k, v in mod.state_dict().items()}      # fp64 reference: no noise of its own
    x = (torch.rand(2, 1, 2403, generator=torch.Generator().manual_seed(8)) * 2 - 1)
    xg = x.cuda().requires_grad_(True)
    outs, feats = mod(xg)
    if training:
        mod(torch.rand(2, 1, 2403, device="cuda"))       # D(real) after D(fake): updates u / v before the backward runs
    _loss(outs, feats).backward()
    torch.cuda.synchronize()
    for k in sd:
        if k.endswith("weight_orig") or k.endswith("bias"):
            sd[k].requires_grad_(True)
    xr = x.double().requires_grad_(True)
    r_outs, r_feats = O.critic_forward(kind, sd, ocfg, xr, training=training)
    _loss(r_outs, r_feats).backward()
    # MPD / MBD run fp32 end to end and agree with fp64 autograd to 6e-7 (measured): bound 1e-5.  MSD's wide layers run on
    # the tensor cores with split-bf16 operands: its feature maps carry 4e-5 of round-off (bound 2e-4, test_gpu_critics.py)
    # and the deep stack's gradients -- cancelling sums over 1e4..1e5 positions -- amplify that ~60x whichever kernels
    # compute the backward (measured worst 2.8e-3 with tensor-core forwards, 8e-4 with all-fp32 forwards:
    # tests/diag_critic_bwd.py); the MSD backward kernels themselves are held to 1e-4 per layer against fp64 above and to
    # 2e-4 against the fp32 path on the same forward below.
    tol = 5e-3 if kind == "msd" else 1e-5
    _close(xg.grad, xr.grad, f"{kind}: d loss / d waveform", tol=tol)
    for name, p in mod.named_parameters():
        assert p.grad is not None, name
        _close(p.grad, sd[name].grad, f"{kind}: {name}", tol=tol)


@pytest.mark.parametrize("kind,mode,tol", [("mpd", "eval", 1e-4), ("mpd", "train", 1e-4), ("mbd", "eval", 1e-4),
                                           ("mbd", "train", 1e-4), ("msd", "train", 2e-2)])
def test_critic_backward_matches_reference_golden(golden_dir, kind, mode, tol):
    """Gradients through the CUDA backward against the golden gradients autograd produced through the REFERENCE classes
    (tests/golden/critics_grad_b2_t2403.npz, oracle/make_golden.py): fresh default-init weights under seed 1234, the same
    loss over every score and feature map.  MPD / MBD: 1e-4 of each tensor's scale (measured 6e-7).  MSD only in .train()
    mode and loosely: with the initialiser's fresh u, v its sigma is a cancelling sum, the stack is ill conditioned and the
    split-bf16 forward's round-off is amplified (2e-3 measured; eval mode: 2e-2 on the score bias, not asserted)."""
    import os
    import numpy as np
    import b200voc
    gold = np.load(os.path.join(golden_dir, "critics_grad_b2_t2403.npz"))
    mod = _host(kind, b200voc.GANConfig(), seed=1234)
    mod.train(mode == "train")
    x = torch.from_numpy(gold["x"]).cuda().requires_grad_(True)
    outs, feats = mod(x)
    _loss(outs, feats).backward()
    torch.cuda.synchronize()
    for name, g in [("x", x.grad)] + [(n, p.grad) for n, p in mod.named_parameters()]:
        flat = g.reshape(-1).cpu()
        _, a_ref, m_ref = gold[f"{kind}.{mode}.{name}.sum"]
        got = flat[torch.from_numpy(gold[f"{kind}.{mode}.{name}.idx"])].numpy()
        err = float(np.abs(got - gold[f"{kind}.{mode}.{name}.val"]).max())
        assert err <= tol * m_ref, f"{kind} {mode} {name}: {err:.3e} vs scale {m_ref:.3e}"


def test_critic_backward_cuda_core_path_and_partial_losses(monkeypatch):
    """B200VOC_DISC_BWD_TC=0 (everything on the fp32 kernels) gives the same gradients as the tensor-core path; a loss on
    the scores alone leaves the feature gradients undefined (None) and still reaches every weight; a waveform that does not
    require grad skips the last dgrad."""
    import b200voc
    cfg = b200voc.GANConfig(disc_kernel_sizes=[15, 41, 41])
    mod = _host("msd", cfg, seed=5).eval()
    x = torch.rand(2, 1, 1500, device="cuda") * 2 - 1
    outs, _ = mod(x)
    sum((o ** 2).mean() for o in outs).backward()
    tc = {n: p.grad.clone() for n, p in mod.named_parameters()}
    assert all(g is not None for g in tc.values())
    mod.zero_grad()
    monkeypatch.setenv("B200VOC_DISC_BWD_TC", "0")
    outs, _ = mod(x)
    sum((o ** 2).mean() for o in outs).backward()
    for n, p in mod.named_parameters():
        _close(p.grad, tc[n], n, tol=2e-4)


def test_critic_forward_backward_is_cuda_graph_capturable():
    """Forward + loss + backward of a critic in ONE CUDA graph (PyTorch's whole-network capture recipe): nothing in the
    autograd node synchronises or allocates outside the capture pool, every launch goes to the capturing stream.  The
    replayed gradients on NEW input values equal the eager ones (the kernels are deterministic: bit for bit)."""
    import b200voc
    mod = _host("mbd", b200voc.GANConfig(), seed=3).eval()
    static_x = (torch.rand(2, 1, 4096, device="cuda") * 2 - 1).requires_grad_(True)

    def step():
        outs, feats = mod(static_x)
        _loss(outs, feats).backward()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            mod.zero_grad(set_to_none=True)
            static_x.grad = None
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    mod.zero_grad(set_to_none=True)
    static_x.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    new_x = torch.rand(2, 1, 4096, device="cuda") * 2 - 1
    with torch.no_grad():
        static_x.copy_(new_x)
    graph.replay()
    torch.cuda.synchronize()
    got = {n: p.grad.clone() for n, p in mod.named_parameters()}
    got_x = static_x.grad.clone()
    mod.zero_grad(set_to_none=True)
    xe = new_x.clone().requires_grad_(True)
    outs, feats = mod(xe)
    _loss(outs, feats).backward()
    torch.cuda.synchronize()
    assert torch.equal(got_x, xe.grad)
    for n, p in mod.named_parameters():
        assert torch.equal(got[n], p.grad), n
