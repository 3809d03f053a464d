"""CPU: the critic (discriminator) oracle against the golden vectors made from the reference classes, the
host-side mirror's parameter layout / seeded init, and the host layer walker driven through a numpy stand-in
for the four C-ABI entry points (tests/host/fake_critic_lib.py).  SURVEY.md section 8(f) rank 4, forward half."""
import contextlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import vocoder7_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host"))

KINDS = ("mpd", "msd", "mbd")


def _host_cls(kind):
    import b200voc
    return {"mpd": b200voc.MultiPeriodDiscriminator, "msd": b200voc.MultiScaleDiscriminator,
            "mbd": b200voc.MultiBandDiscriminator}[kind]


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_reference_golden(golden_dir, kind):
    """tests/golden/critics_b2_t2403.npz holds outputs of the reference classes themselves
    (oracle/make_golden.py); the restatement must reproduce them from seed-regenerated weights."""
    gold = np.load(os.path.join(golden_dir, "critics_b2_t2403.npz"))
    cfg = O.OracleConfig()
    sd = O.make_critic_state(kind, cfg, seed=1234)
    with torch.no_grad():
        outs, feats = O.critic_forward(kind, sd, cfg, torch.from_numpy(gold["x"]))
    n = len([k for k in gold.files if k.startswith(kind + ".out")])
    assert len(outs) == n == {"mpd": 5, "msd": 3, "mbd": 4}[kind]
    for i, o in enumerate(outs):
        ref = gold[f"{kind}.out{i}"]
        assert tuple(o.shape) == ref.shape
        assert float(np.abs(o.numpy() - ref).max()) <= 1e-6
        assert len(feats[i]) == {"mpd": 8, "msd": 10, "mbd": 8}[kind]
        for j, fm in enumerate(feats[i]):
            assert tuple(fm.shape) == tuple(gold[f"{kind}.f{i}.{j}.shape"])
            got = fm.reshape(-1)[torch.from_numpy(gold[f"{kind}.f{i}.{j}.idx"])].numpy()
            assert float(np.abs(got - gold[f"{kind}.f{i}.{j}.val"]).max()) <= 1e-6


@pytest.mark.parametrize("kind", KINDS)
def test_host_module_layout_and_seeded_init(kind):
    """Same state_dict keys / shapes as the reference (spectral_norm's weight_orig, weight_u, weight_v + bias) and,
    constructed under the same seed, the same values: a reference checkpoint loads unchanged."""
    from b200voc import GANConfig
    torch.manual_seed(1234)
    mod = _host_cls(kind)(GANConfig())
    sd, ref = mod.state_dict(), O.make_critic_state(kind, O.OracleConfig(), seed=1234)
    assert sorted(sd.keys()) == sorted(ref.keys())
    for k in ref:
        assert torch.equal(sd[k], ref[k]), k
    missing, unexpected = mod.load_state_dict(ref, strict=True)
    assert not missing and not unexpected


def test_no_cpu_path():
    from b200voc import GANConfig, MultiBandDiscriminator, _lib
    with pytest.raises(_lib.B200VocError):
        MultiBandDiscriminator(GANConfig())(torch.zeros(1, 1, 400))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("B,T", [(2, 487), (1, 64)])
def test_host_layer_walker_against_oracle(monkeypatch, kind, B, T):
    """The host logic (period view with implicit zero padding, pooled scales, time chunks read in place through
    a pointer offset + batch stride) drives a numpy stand-in that uses the kernel's index arithmetic; the result
    must equal the oracle.  Small kernel sizes keep the 1024-channel layers cheap."""
    from fake_critic_lib import FakeCriticLib
    from b200voc import GANConfig, _lib
    cfg = GANConfig(disc_kernel_sizes=[5, 9, 9])
    ocfg = O.OracleConfig(disc_kernel_sizes=[5, 9, 9])
    torch.manual_seed(7)
    mod = _host_cls(kind)(cfg).eval()
    fake = FakeCriticLib()
    monkeypatch.setattr(_lib, "load", lambda: fake)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(_lib, "current_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    x = torch.randn(B, 1, T, generator=torch.Generator().manual_seed(3))
    outs, feats = mod(x)
    sd = {k: v.detach() for k, v in mod.state_dict().items()}
    with torch.no_grad():
        r_outs, r_feats = O.critic_forward(kind, sd, ocfg, x)
    assert len(outs) == len(r_outs) and [len(f) for f in feats] == [len(f) for f in r_feats]
    for a, b in zip(outs, r_outs):
        assert a.shape == b.shape
        assert float((a.detach() - b).abs().max()) <= 1e-4 * max(1.0, float(b.abs().max()))
    for fa, fb in zip(feats, r_feats):
        for a, b in zip(fa, fb):
            assert a.shape == b.shape
            assert float((a.detach() - b).abs().max()) <= 1e-4 * max(1.0, float(b.abs().max()))
    # the spectral-normalised weights are cached per parameter version: a second call launches no new sigma pass
    n_calls = len(fake.conv_calls)
    mod(x)
    assert len(fake.conv_calls) == 2 * n_calls
    if kind == "mbd":      # chunks are read in place: batch stride = T, valid = chunk length
        first = [c for c in fake.conv_calls[:n_calls] if c[1] == 1]
        assert all(c[8] == T for c in first) and {c[9] for c in first} <= {-(-T // 4), T - 3 * (-(-T // 4))}


def _gan_like_loss(outs, feats):
    """A scalar that touches every returned map, the way compute_gan_loss (vocoder7/losses.py:8-52) does: squared
    scores (least-squares GAN terms) and mean-abs features (feature matching)."""
    loss = 0.0
    for o in outs:
        loss = loss + ((o - 1.0) ** 2).mean()
    for fs in feats:
        for j, f in enumerate(fs):
            loss = loss + (0.5 + 0.1 * j) * f.abs().mean()
    return loss


@pytest.mark.parametrize("kind", KINDS)
def test_host_autograd_node_against_oracle(monkeypatch, kind):
    """The backward chain of b200voc/discriminators.py (_CriticStackFn: feature-gradient merge, bias / weight gradients,
    spectral-norm backward, dgrad down to the waveform through the period view / pooled scales / time chunks) driven
    through the numpy stand-in, against torch autograd over the oracle: d loss / d weight_orig, d bias, d waveform."""
    from fake_critic_lib import FakeCriticLib
    from b200voc import GANConfig, _lib
    cfg = GANConfig(disc_kernel_sizes=[5, 9, 9])
    ocfg = O.OracleConfig(disc_kernel_sizes=[5, 9, 9])
    torch.manual_seed(11)
    mod = _host_cls(kind)(cfg).eval()
    fake = FakeCriticLib()
    monkeypatch.setattr(_lib, "load", lambda: fake)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(_lib, "current_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    B, T = 2, 203
    x = torch.randn(B, 1, T, generator=torch.Generator().manual_seed(5)).requires_grad_(True)
    outs, feats = mod(x)
    assert outs[0].requires_grad
    _gan_like_loss(outs, feats).backward()
    # oracle: the same scalar under torch autograd
    sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    for k in sd:
        if k.endswith("weight_orig") or k.endswith("bias"):
            sd[k].requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    r_outs, r_feats = O.critic_forward(kind, sd, ocfg, xr)
    _gan_like_loss(r_outs, r_feats).backward()

    def close(a, b, what):
        assert a is not None, what
        assert a.shape == b.shape, what
        err, scale = float((a - b).abs().max()), float(b.abs().max())
        assert err <= 2e-4 * scale + 1e-7, f"{what}: {err:.3e} vs scale {scale:.3e}"
    close(x.grad, xr.grad, "d loss / d waveform")
    for name, p in mod.named_parameters():
        close(p.grad, sd[name].grad, name)
    # a loss that touches only the scores (the critics' own adversarial terms): feature gradients arrive as None
    mod.zero_grad()
    outs, _ = mod(x.detach())
    sum((o ** 2).mean() for o in outs).backward()
    for k in sd:
        sd[k].grad = None
    r_outs, _ = O.critic_forward(kind, sd, ocfg, x.detach())
    sum((o ** 2).mean() for o in r_outs).backward()
    for name, p in mod.named_parameters():
        close(p.grad, sd[name].grad, name + " (scores only)")


def test_host_autograd_node_training_mode_keeps_its_own_u_v(monkeypatch):
    """.train(): the forward runs the power iteration and updates weight_u / weight_v in place; the reference trainer runs
    D(fake) AND D(real) before any backward (vocoder7/trainer.py:86-115), so the node must keep the u, v of ITS forward --
    the gradients of the first forward, taken after a second forward has moved the buffers, equal torch autograd over the
    oracle's training-mode forward of the first call."""
    from fake_critic_lib import FakeCriticLib
    from b200voc import GANConfig, _lib
    cfg = GANConfig(disc_kernel_sizes=[5, 9, 9])
    ocfg = O.OracleConfig(disc_kernel_sizes=[5, 9, 9])
    torch.manual_seed(13)
    mod = _host_cls("mbd")(cfg).train()
    fake = FakeCriticLib()
    monkeypatch.setattr(_lib, "load", lambda: fake)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(_lib, "current_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    x = torch.randn(2, 1, 211, generator=torch.Generator().manual_seed(6))
    outs, feats = mod(x)                                     # D(fake)
    u_after_first = mod.state_dict()["discriminators.0.0.weight_u"].clone()
    mod(torch.randn(2, 1, 211))                              # D(real): moves u / v again
    assert not torch.equal(mod.state_dict()["discriminators.0.0.weight_u"], u_after_first)
    _gan_like_loss(outs, feats).backward()
    for k in sd:
        if k.endswith("weight_orig") or k.endswith("bias"):
            sd[k].requires_grad_(True)
    r_outs, r_feats = O.critic_forward("mbd", sd, ocfg, x, training=True)
    assert torch.allclose(sd["discriminators.0.0.weight_u"], u_after_first, atol=1e-6)
    _gan_like_loss(r_outs, r_feats).backward()
    for name, p in mod.named_parameters():
        g, r = p.grad, sd[name].grad
        assert float((g - r).abs().max()) <= 5e-4 * float(r.abs().max()) + 1e-7, name


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_oracle_backward_matches_reference_golden(golden_dir, kind, mode):
    """tests/golden/critics_grad_b2_t2403.npz holds gradients that autograd produced through the REFERENCE classes
    (oracle/make_golden.py: .eval() and .train(), a loss over every score and feature map); autograd through the
    restatement on seed-regenerated weights must reproduce them -- this is what pins the reference the critic backward
    kernels are tested against (tests/test_gpu_critics_bwd.py) to the reference's own code."""
    gold = np.load(os.path.join(golden_dir, "critics_grad_b2_t2403.npz"))
    cfg = O.OracleConfig()
    sd = O.make_critic_state(kind, cfg, seed=1234)
    for k in sd:
        if k.endswith("weight_orig") or k.endswith("bias"):
            sd[k].requires_grad_(True)
    x = torch.from_numpy(gold["x"]).clone().requires_grad_(True)
    outs, feats = O.critic_forward(kind, sd, cfg, x, training=(mode == "train"))
    _gan_like_loss(outs, feats).backward()
    names = [k for k in sd if sd[k].requires_grad]
    assert len(names) == {"mpd": 50, "msd": 36, "mbd": 40}[kind]
    for name, g in [("x", x.grad)] + [(n, sd[n].grad) for n in names]:
        flat = g.reshape(-1)
        s_ref, a_ref, m_ref = gold[f"{kind}.{mode}.{name}.sum"]
        got = flat[torch.from_numpy(gold[f"{kind}.{mode}.{name}.idx"])].numpy()
        assert float(np.abs(got - gold[f"{kind}.{mode}.{name}.val"]).max()) <= 1e-5 * m_ref, name
        assert abs(float(flat.double().abs().sum()) - a_ref) <= 1e-5 * a_ref, name


def test_short_waveforms_raise():
    from fake_critic_lib import FakeCriticLib  # noqa: F401  (only the out_len rule is needed)
    from b200voc import _lib
    lib = _lib.load()
    assert lib.b200voc_disc_conv_out_len(2403, 41, 2, 20) == 1202
    assert lib.b200voc_disc_conv_out_len(1202, 5, 3, 2) == 401
    assert lib.b200voc_disc_conv_out_len(1, 3, 1, 1) == 1
    assert lib.b200voc_disc_conv_out_len(0, 3, 1, 1) == 0
    assert lib.b200voc_disc_conv_out_len(2, 15, 2, 3) == 0


def test_entry_points_validate_arguments_before_touching_the_device():
    """Bad arguments are rejected with ERR_BAD_ARG and a message, without a CUDA call (works on a CPU-only host)."""
    from b200voc import _lib
    lib = _lib.load()
    assert lib.b200voc_disc_conv(0, 0, 0, 1, 1, 4, 100, 1, 5, 3, 2, 0, 0, 0.2, 0, 0, 0) == _lib.ERR_BAD_ARG
    assert b"null" in lib.b200voc_last_error_string()
    one = 16   # any non-null value: the shape checks come before the pointers are used
    assert lib.b200voc_disc_conv(one, one, one, 1, 1, 4, 2, 1, 15, 2, 3, 0, 0, 0.2, one, 0, 0) == _lib.ERR_BAD_ARG
    assert b"shorter than the kernel" in lib.b200voc_last_error_string()
    assert lib.b200voc_disc_conv(one, one, one, 0, 1, 4, 100, 1, 5, 3, 2, 0, 0, 0.2, one, 0, 0) == _lib.ERR_BAD_ARG
    assert lib.b200voc_spectral_norm_weight(one, one, one, 0, 5, one, one, 0) == _lib.ERR_BAD_ARG
    assert lib.b200voc_avg_pool1d_k4s2p1(one, 1, 1, one, 0) == _lib.ERR_BAD_ARG


def test_oracle_training_mode_spectral_norm_is_pinned_to_torch():
    """The oracle's training-mode spectral norm (one power iteration per forward, u / v updated in place) against the
    implementation the reference calls, ``torch.nn.utils.spectral_norm`` on a module in .train() (discriminators.py:76-89
    builds exactly this; the trainer never switches the critics to eval, vocoder7/trainer.py:86-115): weight, u and v
    after one and after two forwards, then a whole MSD stack in training mode against the restated forward."""
    import torch.nn as nn
    torch.manual_seed(7)
    conv = nn.utils.spectral_norm(nn.Conv1d(16, 64, kernel_size=41, stride=1, padding=20)).train()
    x = torch.randn(2, 16, 50)
    w0, u, v = conv.weight_orig.detach().clone(), conv.weight_u.detach().clone(), conv.weight_v.detach().clone()
    for _ in range(2):
        with torch.no_grad():
            y = conv(x)
        w, u, v = O.spectral_norm_power_iteration(w0, u, v)
        assert torch.allclose(conv.weight_u, u, rtol=0, atol=1e-6) and torch.allclose(conv.weight_v, v, rtol=0, atol=1e-6)
        ref = torch.nn.functional.conv1d(x, w, conv.bias.detach(), padding=20)
        assert float((y - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
    # whole stack: a reference-shaped module in .train() vs critic_forward(training=True)
    cfg = O.OracleConfig()
    sd = O.make_critic_state("msd", cfg, seed=1234)
    plans, _ = O.critic_plans("msd", cfg)
    torch.manual_seed(1234)
    discs = []
    for plan in plans:                                        # same construction order as make_critic_state
        mods = []
        for cin, cout, k, st, pad, act in plan:
            mods.append(nn.utils.spectral_norm(nn.Conv1d(cin, cout, k, stride=st, padding=pad)))
            if act:
                mods.append(nn.LeakyReLU(0.2))
        discs.append(nn.Sequential(*mods).train())
    for d, seq in enumerate(discs):
        for j, m in enumerate(seq):
            if not isinstance(m, nn.LeakyReLU):
                for name in ("weight_orig", "weight_u", "weight_v", "bias"):
                    assert torch.equal(getattr(m, name), sd[f"discriminators.{d}.{j}.{name}"]), (d, j, name)
    xw = torch.rand(1, 1, 900) * 2 - 1
    sd_t = {k: t.clone() for k, t in sd.items()}
    outs, feats = O.critic_forward("msd", sd_t, cfg, xw, training=True)
    with torch.no_grad():
        got = discs[0](xw)
    assert float((got - outs[0]).abs().max()) <= 2e-4 * max(1.0, float(outs[0].abs().max()))
    assert torch.allclose(discs[0][0].weight_u, sd_t["discriminators.0.0.weight_u"], atol=1e-6)
    assert not torch.equal(sd_t["discriminators.0.0.weight_u"], sd["discriminators.0.0.weight_u"])


def test_oracle_training_mode_gradient_is_pinned_to_torch():
    """Backward of the oracle's training-mode spectral norm == autograd through ``torch.nn.utils.spectral_norm`` on a
    module in .train() (u, v from the power iteration are constants of the graph): the reference the critic backward
    kernels are tested against (tests/test_gpu_critics_bwd.py)."""
    import torch.nn as nn
    torch.manual_seed(3)
    conv = nn.utils.spectral_norm(nn.Conv1d(8, 16, 5, padding=2)).train()
    x = torch.randn(2, 8, 30)
    w0 = conv.weight_orig.detach().clone().requires_grad_(True)
    u, v = conv.weight_u.clone(), conv.weight_v.clone()
    (conv(x) ** 2).mean().backward()
    w, _, _ = O.spectral_norm_power_iteration(w0, u, v)
    (torch.nn.functional.conv1d(x, w, conv.bias.detach(), padding=2) ** 2).mean().backward()
    assert float((w0.grad - conv.weight_orig.grad).abs().max()) <= 1e-6 * float(w0.grad.abs().max())


def test_backward_entry_points_validate_arguments_before_touching_the_device():
    """The backward half of the critic ABI rejects bad arguments with ERR_BAD_ARG and a message, without a CUDA call, and
    its capability / size queries are pure host functions (works on a CPU-only host)."""
    from b200voc import _lib
    lib = _lib.load()
    one = 16
    assert lib.b200voc_disc_conv_dgrad(0, one, 1, 4, 16, 100, 1, 5, 3, 2, 0, 0, 0, one, 0) == _lib.ERR_BAD_ARG
    assert b"null" in lib.b200voc_last_error_string()
    assert lib.b200voc_disc_conv_dgrad(one, one, 1, 4, 16, 2, 1, 15, 2, 3, 0, 0, 0, one, 0) == _lib.ERR_BAD_ARG
    assert b"shorter than the kernel" in lib.b200voc_last_error_string()
    assert lib.b200voc_disc_conv_wgrad(one, one, 0, 4, 16, 100, 1, 5, 3, 2, 0, 0, one, 0, 0) == _lib.ERR_BAD_ARG
    assert lib.b200voc_disc_conv_wgrad(one, 0, 1, 4, 16, 100, 1, 5, 3, 2, 0, 0, one, 0, 0) == _lib.ERR_BAD_ARG
    assert lib.b200voc_disc_bias_grad(one, 1, 0, 10, one, 0) == _lib.ERR_BAD_ARG
    assert lib.b200voc_disc_lrelu_bwd(0, 0, one, 0, 0.2, 10, one, 0) == _lib.ERR_BAD_ARG          # ga without y
    assert lib.b200voc_avg_pool1d_k4s2p1_bwd(one, 1, 1, one, 0) == _lib.ERR_BAD_ARG
    assert lib.b200voc_spectral_norm_bwd(one, one, one, one, one, 0, 5, one, one, 0) == _lib.ERR_BAD_ARG
    assert lib.b200voc_spectral_norm_bwd(one, one, one, one, one, 4, 5, one, 4, 0) == _lib.ERR_BAD_ARG   # unaligned scratch
    assert lib.b200voc_spectral_norm_bwd_scratch_bytes() >= 8
    # tensor-core forms: shapes that qualify / do not
    assert lib.b200voc_disc_conv_wgrad_tc_supported(4, 256, 1024, 1379, 41, 1, 1, 20) == 1
    assert lib.b200voc_disc_conv_wgrad_tc_supported(4, 64, 256, 2757, 15, 1, 1, 7) == 1
    assert lib.b200voc_disc_conv_wgrad_tc_supported(4, 16, 64, 2757, 15, 2, 1, 7) == 0            # strided
    assert lib.b200voc_disc_conv_wgrad_tc_supported(4, 256, 1, 1379, 3, 1, 1, 1) == 0             # score layer
    assert lib.b200voc_disc_conv_wgrad_tc_supported(4, 64, 256, 100, 5, 1, 3, 2) == 0             # period columns
    assert lib.b200voc_disc_conv_dgrad_tc_supported(256, 1024, 41, 1, 1, 20) == 1
    assert lib.b200voc_disc_conv_dgrad_tc_supported(64, 256, 15, 1, 1, 7) == 0                    # needs the padded weight
    assert lib.b200voc_disc_conv_dgrad_tc_supported(128, 256, 15, 1, 1, 7) == 1
    nb = lib.b200voc_disc_conv_wgrad_tc_workspace_bytes(4, 256, 1024, 1379, 41, 20)
    kdim = 3 * 4 * 1408                                                                           # positions padded to 64 per item
    assert nb >= (256 * 41 + 1024) * kdim * 2 and nb % 1024 == 0
    assert lib.b200voc_disc_conv_wgrad_tc(one, one, 4, 16, 64, 100, 15, 7, one, 1024, 1 << 30, 0) == _lib.ERR_BAD_ARG
    # a batch whose packed operands exceed the 1 GiB cap is processed in chunks: the workspace stays bounded
    big = lib.b200voc_disc_conv_wgrad_tc_workspace_bytes(64, 256, 1024, 1379, 41, 20)
    assert 0 < big < (5 << 28)
    assert lib.b200voc_disc_conv_wgrad_scratch_bytes(4, 1, 4, 22050, 1, 15, 2, 7) > 0             # sliced first layer
    assert lib.b200voc_disc_conv_wgrad_scratch_bytes(4, 256, 1024, 1379, 1, 41, 1, 20) == 0
