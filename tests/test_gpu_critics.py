"""GPU parity of the critic (discriminator) forwards -- SURVEY.md section 8(f) rank 4, forward half -- through
the C ABI (b200voc_disc_conv / b200voc_spectral_norm_weight / b200voc_avg_pool1d_k4s2p1) against the golden
vectors made from the reference classes and against the CPU oracle.

Tolerance: the kernels accumulate in fp32 like the reference but in a different order (up to Cin*K = 10 496
terms per output), and the reference's own fp32 sigma = u.(W v) of a freshly initialised spectral norm is a
nearly cancelling sum known only to ~1e-5 relative (every later map inherits that factor), so every score /
feature map must agree to 2e-4 * max(1, max|ref|) -- round-off, 3-4 orders of magnitude below what an indexing
error produces."""
import os

import numpy as np
import pytest
import torch

from oracle import vocoder7_oracle as O

pytestmark = pytest.mark.gpu
KINDS = ("mpd", "msd", "mbd")
TOL = 2e-4


def _host(kind, cfg=None, seed=1234):
    import b200voc
    cls = {"mpd": b200voc.MultiPeriodDiscriminator, "msd": b200voc.MultiScaleDiscriminator,
           "mbd": b200voc.MultiBandDiscriminator}[kind]
    torch.manual_seed(seed)
    return cls(cfg or b200voc.GANConfig()).eval().cuda()


def _close(got: torch.Tensor, ref: torch.Tensor, what: str):
    assert tuple(got.shape) == tuple(ref.shape), what
    err = float((got.cpu() - ref).abs().max())
    assert err <= TOL * max(1.0, float(ref.abs().max())), f"{what}: max-abs error {err:.3e}"


@pytest.mark.parametrize("kind", KINDS)
def test_critic_matches_reference_golden_and_oracle(golden_dir, kind):
    gold = np.load(os.path.join(golden_dir, "critics_b2_t2403.npz"))
    x = torch.from_numpy(gold["x"])
    mod = _host(kind)
    outs, feats = mod(x.cuda())
    torch.cuda.synchronize()
    # (1) the reference's own outputs (scores in full, 512 samples + checksums of every feature map)
    for i, o in enumerate(outs):
        _close(o, torch.from_numpy(gold[f"{kind}.out{i}"]), f"{kind} score {i}")
        for j, fm in enumerate(feats[i]):
            assert tuple(fm.shape) == tuple(gold[f"{kind}.f{i}.{j}.shape"])
            flat = fm.reshape(-1).cpu()
            ref = torch.from_numpy(gold[f"{kind}.f{i}.{j}.val"])
            _close(flat[torch.from_numpy(gold[f"{kind}.f{i}.{j}.idx"])], ref, f"{kind} feature {i}.{j} samples")
            s_ref, a_ref = gold[f"{kind}.f{i}.{j}.sum"]
            assert abs(float(flat.double().sum()) - s_ref) <= 1e-4 * max(1.0, a_ref)
            assert abs(float(flat.double().abs().sum()) - a_ref) <= 1e-4 * max(1.0, a_ref)
    # (2) every element of every map against the CPU oracle on the same seed-regenerated weights
    sd = O.make_critic_state(kind, O.OracleConfig(), seed=1234)
    with torch.no_grad():
        r_outs, r_feats = O.critic_forward(kind, sd, O.OracleConfig(), x)
    assert len(outs) == len(r_outs)
    for i in range(len(outs)):
        _close(outs[i], r_outs[i], f"{kind} score {i} vs oracle")
        assert len(feats[i]) == len(r_feats[i])
        for j in range(len(feats[i])):
            _close(feats[i][j], r_feats[i][j], f"{kind} feature {i}.{j} vs oracle")


@pytest.mark.parametrize("kind,B,T", [("mpd", 3, 4099), ("mbd", 3, 4099), ("mpd", 1, 23), ("mbd", 1, 61),
                                      ("msd", 1, 50)])
def test_critic_ragged_and_tiny_inputs(kind, B, T):
    """Lengths that are not multiples of any period / of the chunk count, and inputs barely longer than the
    receptive field (single-row maps)."""
    import b200voc
    cfg = b200voc.GANConfig(disc_kernel_sizes=[15, 9, 9]) if kind == "msd" else b200voc.GANConfig()
    ocfg = O.OracleConfig(disc_kernel_sizes=list(cfg.disc_kernel_sizes))
    mod = _host(kind, cfg, seed=99)
    x = torch.randn(B, 1, T, generator=torch.Generator().manual_seed(T)) * 0.5
    outs, feats = mod(x.cuda())
    sd = {k: v.detach().cpu() for k, v in mod.state_dict().items()}
    with torch.no_grad():
        r_outs, r_feats = O.critic_forward(kind, sd, ocfg, x)
    assert len(outs) == len(r_outs)
    for i in range(len(outs)):
        _close(outs[i], r_outs[i], f"{kind} score {i}")
        for j in range(len(feats[i])):
            _close(feats[i][j], r_feats[i][j], f"{kind} feature {i}.{j}")


def test_spectral_norm_and_pool_entry_points():
    from b200voc import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    w0 = torch.randn(64, 16, 41, generator=g)
    u = torch.nn.functional.normalize(torch.randn(64, generator=g), dim=0)
    v = torch.nn.functional.normalize(torch.mv(w0.flatten(1).t(), u), dim=0)      # one power iteration, as a
    u = torch.nn.functional.normalize(torch.mv(w0.flatten(1), v), dim=0)          # trained critic's u, v are
    ref = O.spectral_norm_weight(w0, u, v)
    w0d, ud, vd = w0.cuda(), u.cuda(), v.cuda()
    out, sigma = torch.empty_like(w0d), torch.empty(1, device="cuda")
    _lib.check(lib.b200voc_spectral_norm_weight(w0d.data_ptr(), ud.data_ptr(), vd.data_ptr(), 64, 16 * 41,
                                                out.data_ptr(), sigma.data_ptr(), _lib.current_stream()))
    s_ref = float(torch.dot(u, torch.mv(w0.flatten(1), v)))
    assert abs(float(sigma) - s_ref) <= 1e-5 * abs(s_ref) + 1e-7
    assert float((out.cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    for L in (2, 3, 7, 2403):
        x = torch.randn(5, 1, L, generator=g)
        want = torch.nn.functional.avg_pool1d(x, 4, 2, 1)
        y = torch.empty(5, 1, want.shape[2], device="cuda")
        _lib.check(lib.b200voc_avg_pool1d_k4s2p1(x.cuda().data_ptr(), 5, L, y.data_ptr(), _lib.current_stream()))
        assert float((y.cpu() - want).abs().max()) <= 2e-6


def test_critic_errors_and_weight_cache():
    import b200voc
    mod = _host("msd", b200voc.GANConfig(disc_kernel_sizes=[5, 5, 5]))
    with pytest.raises(ValueError):
        mod(torch.zeros(2, 2, 100, device="cuda"))          # not [B, 1, T]
    with pytest.raises(ValueError):
        mod(torch.zeros(1, 1, 1, device="cuda"))            # too short for avg_pool1d(4, 2, 1)
    x = torch.randn(2, 1, 300, device="cuda")
    a, _ = mod(x)
    with torch.no_grad():                                    # an optimiser step bumps the parameter version:
        mod.discriminators[0][0].weight_orig.add_(0.05)      # the cached normalised weights must be rebuilt
    b, _ = mod(x)
    sd = {k: v.detach().cpu() for k, v in mod.state_dict().items()}
    with torch.no_grad():
        r, _ = O.critic_forward("msd", sd, O.OracleConfig(disc_kernel_sizes=[5, 5, 5]), x.cpu())
    _close(b[0], r[0], "after in-place weight update")
    assert a[0].shape == b[0].shape


@pytest.mark.parametrize("B,Cin,Cout,L,K", [(2, 64, 256, 300, 15), (1, 256, 128, 130, 41), (3, 64, 128, 1, 15),
                                            (1, 128, 256, 129, 41), (2, 256, 1024, 345, 41), (1, 64, 128, 127, 1)])
def test_disc_conv_tensor_core_layer_matches_fp64(B, Cin, Cout, L, K):
    """b200voc_disc_conv_tc (tcgen05 implicit GEMM, split-bf16 operands) through the C ABI against fp64 conv1d and against
    the fp32 CUDA-core kernel it replaces: ragged lengths (L = 1, 127, 129, 130: tiles that end inside / just past a
    128-row block), both kernel sizes of the reference (15, 41), the largest layer shape (256 -> 1024), large-magnitude
    weights (a fresh spectral norm divides by sigma ~ 1e-3).  Bound: 6e-5 of the map's scale -- operands carry 16
    significant bits (bf16 hi + lo, the lo.lo product dropped), a third of the critics' 2e-4 feature tolerance."""
    import ctypes as C
    from b200voc import _lib
    lib = _lib.load()
    assert lib.b200voc_disc_conv_tc_supported(Cin, Cout, K, 1, 1) == 1
    assert lib.b200voc_disc_conv_tc_supported(Cin, Cout, K, 2, 1) == 0 and lib.b200voc_disc_conv_tc_supported(16, Cout, K, 1, 1) == 0
    g = torch.Generator().manual_seed(Cin + L + K)
    x = (torch.randn(B, Cin, L, generator=g) * 3.0)
    w = torch.randn(Cout, Cin, K, generator=g) * 40.0 / (Cin * K) ** 0.5
    b = torch.randn(Cout, generator=g)
    pad = K // 2
    ref = torch.nn.functional.conv1d(x.double(), w.double(), b.double(), padding=pad)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    wsplit = torch.empty(int(lib.b200voc_disc_split_weight_elems(Cout, Cin, K)), device="cuda", dtype=torch.bfloat16)
    st = _lib.current_stream()
    _lib.check(lib.b200voc_disc_pack_weight_split(_lib.ptr(wd), Cout, Cin, K, _lib.ptr(wsplit), st))
    # outputs and workspace sit inside larger buffers with sentinel guards on both sides: the kernels must not write a
    # single element outside what the ABI says they own (ragged last tiles, the all-out-of-range rows of a 128-row block)
    G = 4096
    ws_bytes = int(lib.b200voc_disc_conv_tc_workspace_bytes(B, Cin, L))
    ws_big = torch.full((ws_bytes + 2 * G,), 0xA5, device="cuda", dtype=torch.uint8)
    ws = ws_big[G:G + ws_bytes]
    n_out = B * Cout * L
    pre_big = torch.full((n_out + 2 * G,), -777.0, device="cuda")
    act_big = torch.full((n_out + 2 * G,), -777.0, device="cuda")
    y_pre, y_act = pre_big[G:G + n_out].view(B, Cout, L), act_big[G:G + n_out].view(B, Cout, L)
    y_pre.fill_(float("nan")); y_act.fill_(float("nan"))
    _lib.check(lib.b200voc_disc_conv_tc(_lib.ptr(xd), _lib.ptr(wsplit), _lib.ptr(bd), B, Cin, Cout, L, K, pad, 0.2,
                                        _lib.ptr(y_pre), _lib.ptr(y_act), _lib.ptr(ws), ws.numel(), st))
    torch.cuda.synchronize()
    for big in (pre_big, act_big):
        assert bool((big[:G] == -777.0).all()) and bool((big[G + n_out:] == -777.0).all()), "write outside the output map"
    assert bool((ws_big[:G] == 0xA5).all()) and bool((ws_big[G + ws_bytes:] == 0xA5).all()), "write outside the workspace"
    scale = float(ref.abs().max())
    err = float((y_pre.cpu().double() - ref).abs().max())
    assert err <= 6e-5 * scale, f"conv map: {err:.3e} vs scale {scale:.3e}"
    err_a = float((y_act.cpu().double() - torch.nn.functional.leaky_relu(ref, 0.2)).abs().max())
    assert err_a <= 6e-5 * scale
    # the fp32 CUDA-core kernel on the same inputs (the path B200VOC_DISC_TC=0 keeps)
    y2 = torch.empty_like(y_pre)
    _lib.check(lib.b200voc_disc_conv(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), B, Cin, Cout, L, 1, K, 1, pad, 0, 0, 0.2,
                                     _lib.ptr(y2), None, st))
    torch.cuda.synchronize()
    assert float((y2 - y_pre).abs().max()) <= 6e-5 * scale
    # argument validation: loud, not silent
    assert lib.b200voc_disc_conv_tc(_lib.ptr(xd), _lib.ptr(wsplit), _lib.ptr(bd), B, Cin, Cout, L, K, pad, 0.2,
                                    _lib.ptr(y_pre), _lib.ptr(y_act), _lib.ptr(ws), 16, st) == _lib.ERR_BAD_ARG


@pytest.mark.parametrize("kind", KINDS)
def test_critic_training_mode_spectral_norm(kind):
    """.train(): every forward first runs one power iteration per layer (b200voc_spectral_norm_train) and updates the
    weight_u / weight_v buffers in place, as torch.nn.utils.spectral_norm does for the reference trainer
    (vocoder7/trainer.py:86-115).  Two consecutive forwards against the oracle's training-mode forward: maps, u and v;
    then .eval() uses the updated vectors."""
    cfg = O.OracleConfig()
    sd = O.make_critic_state(kind, cfg, seed=1234)
    mod = _host(kind).train()
    x = torch.rand(2, 1, 2403) * 2 - 1
    sd_t = {k: t.clone() for k, t in sd.items()}
    for step in range(2):
        outs, feats = mod(x.cuda())
        torch.cuda.synchronize()
        with torch.no_grad():
            r_outs, r_feats = O.critic_forward(kind, sd_t, cfg, x, training=True)
        for i in range(len(outs)):
            _close(outs[i], r_outs[i], f"{kind} score {i} (training step {step})")
            for j in range(len(feats[i])):
                _close(feats[i][j], r_feats[i][j], f"{kind} feature {i}.{j} (training step {step})")
        got = mod.state_dict()
        for k in sd_t:
            if k.endswith("weight_u") or k.endswith("weight_v"):
                assert float((got[k].cpu() - sd_t[k]).abs().max()) <= 2e-5, (k, step)
    assert not torch.equal(mod.state_dict()["discriminators.0.0.weight_u"].cpu(), sd["discriminators.0.0.weight_u"])
    mod.eval()
    outs, _ = mod(x.cuda())
    with torch.no_grad():
        r_outs, _ = O.critic_forward(kind, sd_t, cfg, x)
    for i in range(len(outs)):
        _close(outs[i], r_outs[i], f"{kind} score {i} (eval after training steps)")
