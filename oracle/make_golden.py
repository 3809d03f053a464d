"""Pin the oracle against the reference itself and write tests/golden/*.npz.  TEST INFRASTRUCTURE.

Run in the authoring container only (needs /root/reference, which does not travel to the GPU
box):   python oracle/make_golden.py

What it does
  1. imports ``/root/reference/vocoder7/{config,generator,stft}.py`` byte-for-byte; the two
     modules generator.py needs but the reference does not ship (``vocoder7.residual``,
     ``vocoder7.attention``) are injected through ``sys.modules`` from oracle/vocoder7_oracle.py
     (repairs R2/R3); ``hidden_dim`` is added by subclassing the reference GANConfig (R1);
  2. checks that the functional restatement ``generator_forward`` equals the reference
     ``Generator.forward`` (same seed -> same default-init weights -> same output to fp32
     round-off), that the state_dict key/shape layout is identical, and that every GANConfig
     default equals the OracleConfig default;
  3. documents that ``LearnableSTFT.forward`` raises as shipped (F2) and checks the repaired
     restatement against torchaudio's own Spectrogram / MelSpectrogram;
  4. writes small golden input/output vectors used by the CPU and GPU parity tests.
"""
from __future__ import annotations

import dataclasses
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import vocoder7_oracle as O  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.path.insert(0, REF)
    res = types.ModuleType("vocoder7.residual")
    res.ResidualBlock = O.ResidualBlock
    att = types.ModuleType("vocoder7.attention")
    att.SelfAttention = O.SelfAttention
    import vocoder7  # noqa: F401  (empty __init__)
    sys.modules["vocoder7.residual"] = res
    sys.modules["vocoder7.attention"] = att
    from vocoder7 import config as rcfg, generator as rgen, stft as rstft
    return rcfg, rgen, rstft


def weight_fingerprint(sd):
    """Order-independent checksum of a state_dict so tests can verify that seed-regenerated
    weights on another box are the ones the goldens were made with."""
    out = {}
    for k, v in sd.items():
        v = v.double()
        out[k] = np.array([float(v.sum()), float(v.abs().sum()), float((v * v).sum())])
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    rcfg, rgen, rstft = import_reference()

    # ---- config defaults field by field ------------------------------------------------------
    ref_defaults = rcfg.GANConfig()
    ora_defaults = O.OracleConfig()
    for f in dataclasses.fields(rcfg.GANConfig):
        assert getattr(ref_defaults, f.name) == getattr(ora_defaults, f.name), f.name
    assert not hasattr(ref_defaults, "hidden_dim")          # F1
    print("config: all", len(dataclasses.fields(rcfg.GANConfig)), "GANConfig defaults match")

    @dataclasses.dataclass
    class RepairedCfg(rcfg.GANConfig):                       # R1
        hidden_dim: int = 512

    # ---- Generator: reference class (with injected R2/R3) vs functional restatement ----------
    cases = {
        "gen_b1_t12_attn": dict(B=1, T=12, attn=True, window=None, kw={}),
        "gen_b2_t9_noattn": dict(B=2, T=9, attn=False, window=None, kw={}),
        "gen_b2_t9_noattn_drop": dict(B=2, T=9, attn=False, window=None,
                                      kw=dict(style_drop=True, emo_drop=False, w_style=0.7, w_emo=1.3)),
        "gen_b1_t16_attnwin": dict(B=1, T=16, attn=True, window=512, kw={}),
    }
    fp_written = False
    for name, c in cases.items():
        torch.manual_seed(1234)
        ref_model = rgen.Generator(RepairedCfg()).eval()
        for m in ref_model.modules():
            if isinstance(m, O.SelfAttention):
                m.enabled, m.window = c["attn"], c["window"]
        ocfg = O.OracleConfig(use_attention=c["attn"], attn_window=c["window"])
        ora_model = O.make_generator(ocfg, seed=1234)
        sd_ref, sd_ora = ref_model.state_dict(), ora_model.state_dict()
        assert list(sd_ref.keys()) == list(sd_ora.keys())
        for k in sd_ref:
            assert sd_ref[k].shape == sd_ora[k].shape and torch.equal(sd_ref[k], sd_ora[k]), k
        mel, pros, sty, emo = O.synthetic_inputs(c["B"], c["T"], seed=4321)
        with torch.no_grad():
            y_ref = ref_model(mel, pros, sty, emo, **c["kw"])
            taps = {}
            y_ora = O.generator_forward(sd_ora, ocfg, mel, pros, sty, emo, taps=taps, **c["kw"])
            y64 = O.generator_forward({k: v.double() for k, v in sd_ora.items()}, ocfg, mel.double(),
                                      pros.double(), sty.double(), emo.double(), **c["kw"])
        d = float((y_ref - y_ora).abs().max())
        d64 = float((y_ora.double() - y64).abs().max())
        print(f"{name}: |reference - restatement|max = {d:.3e}   fp32-vs-fp64 floor = {d64:.3e}  "
              f"out {tuple(y_ref.shape)} range [{float(y_ref.min()):.3f},{float(y_ref.max()):.3f}]")
        assert d <= 1e-6, "restatement diverges from the reference Generator"
        save = dict(mel=mel.numpy(), prosody=pros.numpy(), style=sty.numpy(), emotion=emo.numpy(),
                    wav=y_ref.numpy(), wav_fp64=y64.numpy(),
                    cond=taps["cond"].numpy())
        # a thin slice of every intermediate (band 0, batch 0, first 8 channels, first 64 steps)
        for k, v in taps.items():
            if k == "cond":
                continue
            save["tap_" + k] = v[0][0, :8, :64].numpy().copy()
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **save)
        if not fp_written:
            fp = weight_fingerprint(sd_ora)
            np.savez_compressed(os.path.join(GOLD, "weights_seed1234_fingerprint.npz"), **fp)
            fp_written = True

    # ---- STFT: the shipped forward raises (F2); repaired restatement vs torchaudio ------------
    import torchaudio
    g = torch.Generator().manual_seed(99)
    wav = (torch.rand(3, 1, 4000, generator=g) * 2 - 1)
    try:
        rstft.LearnableSTFT(1024, 256)(wav)
        shipped_ok = True
    except TypeError as e:
        shipped_ok = False
        print("stft: LearnableSTFT.forward as shipped raises TypeError (F2):", str(e)[:70])
    assert not shipped_ok
    save = dict(wav=wav.numpy())
    for n_fft in (512, 1024, 2048):
        torch.manual_seed(7 + n_fft)
        ref_mod = rstft.LearnableSTFT(n_fft, 256)            # window buffer + filterbank ~ randn
        assert torch.equal(ref_mod.window, O.hann_window(n_fft))
        ta = torchaudio.transforms.Spectrogram(n_fft=n_fft, hop_length=256, power=None)(wav.squeeze(1))
        mine = O.stft_complex(wav.squeeze(1), n_fft, 256)
        assert float((ta - mine).abs().max()) <= 1e-5, n_fft
        mag = O.learnable_stft_forward(wav, ref_mod.filterbank.detach(), n_fft, 256)
        save[f"gain_{n_fft}"] = ref_mod.filterbank.detach().numpy()
        save[f"mag_{n_fft}"] = mag.numpy()
        print(f"stft n_fft={n_fft}: restatement == torchaudio Spectrogram; out {tuple(mag.shape)}")
    ta_mel = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, hop_length=256,
                                                  n_mels=80)(wav.squeeze(1))
    my_mel = O.mel_spectrogram(wav.squeeze(1))
    rel = float(((ta_mel - my_mel).abs() / (ta_mel.abs() + 1e-3)).max())
    print(f"mel: restatement vs torchaudio MelSpectrogram max rel diff {rel:.2e}")
    assert rel < 1e-5
    fb_ta = torchaudio.functional.melscale_fbanks(513, 0.0, 11025.0, 80, 22050, norm=None, mel_scale="htk")
    assert float((fb_ta - O.mel_filterbank()).abs().max()) < 1e-7
    save["logmel"] = O.log_mel(wav.squeeze(1)).numpy()
    save["logmel_fp64"] = O.log_mel(wav.squeeze(1).double()).numpy()
    spec = O.stft_complex(wav.squeeze(1), 1024, 256)
    rt = O.istft(spec, 1024, 256, wav.shape[-1])
    print(f"istft(stft(x)) max-abs {float((rt - wav.squeeze(1)).abs().max()):.2e}")
    save["spec_re"], save["spec_im"] = spec.real.numpy(), spec.imag.numpy()
    save["istft"] = rt.numpy()
    # STFTLoss (stft.py:48-54) with the three gains above
    wav2 = (torch.rand(3, 1, 4000, generator=g) * 2 - 1)
    loss = O.stft_loss_forward(wav, wav2, [torch.from_numpy(save[f"gain_{n}"]) for n in (512, 1024, 2048)],
                               [512, 1024, 2048], 256, 2.0)
    save["wav2"], save["stft_loss"] = wav2.numpy(), np.array(float(loss))
    np.savez_compressed(os.path.join(GOLD, "stft_b3_n4000.npz"), **save)
    print("golden vectors written to", GOLD)
    for f in sorted(os.listdir(GOLD)):
        print(f"  {f}: {os.path.getsize(os.path.join(GOLD, f)) / 1024:.1f} KiB")

    # ---- GlobalStyleTokens: reference class vs restatement, golden vectors -------
    from vocoder7 import gst as rgst
    torch.manual_seed(1234)
    ref_gst = rgst.GlobalStyleTokens(rcfg.GANConfig()).eval()
    sd_ref = {k: v.detach() for k, v in ref_gst.state_dict().items()}
    sd_ora = O.make_gst_state(seed=1234)
    assert list(sd_ref.keys()) == list(sd_ora.keys())
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_ora[k]), k
    g = torch.Generator().manual_seed(77)
    mel_ref = torch.randn(3, 80, 150, generator=g) * 2.0 - 4.0     # log-mel-like range
    with torch.no_grad():
        y_ref = ref_gst(mel_ref)
        y_ora = O.gst_forward(sd_ora, mel_ref)
    d = float((y_ref - y_ora).abs().max())
    d0 = float((y_ref - sd_ref["tokens"].sum(0)).abs().max())
    print(f"gst: |reference - restatement|max = {d:.3e};  |reference - sum_n tokens|max = {d0:.3e} "
          "(softmax and einsum run over the same axis: the style is input-independent up to fp32 rounding)")
    assert d <= 1e-6
    np.savez_compressed(os.path.join(GOLD, "gst_b3_t150.npz"), mel=mel_ref.numpy(), style=y_ref.numpy(),
                        **{"sd." + k: v.numpy() for k, v in sd_ref.items()})

    # ---- critics (forward half): reference classes vs restatement, golden vectors ----------
    from vocoder7 import discriminators as rdisc
    cfg_ref = rcfg.GANConfig()
    x = torch.randn(2, 1, 2403, generator=torch.Generator().manual_seed(5)) * 0.3   # 2403: ragged for p=2,5,7,11 and chunk(4)
    save = {"x": x.numpy()}
    for kind, cls in (("mpd", rdisc.MultiPeriodDiscriminator), ("msd", rdisc.MultiScaleDiscriminator),
                      ("mbd", rdisc.MultiBandDiscriminator)):
        torch.manual_seed(1234)
        ref = cls(cfg_ref).eval()
        sd_ref = {k: v.detach() for k, v in ref.state_dict().items()}
        sd_ora = O.make_critic_state(kind, cfg_ref, seed=1234)
        assert sorted(sd_ref) == sorted(sd_ora), kind
        assert all(torch.equal(sd_ref[k], sd_ora[k]) for k in sd_ref), kind
        with torch.no_grad():
            o_ref, f_ref = ref(x)
            o_ora, f_ora = O.critic_forward(kind, sd_ora, cfg_ref, x)
        assert len(o_ref) == len(o_ora) and [len(f) for f in f_ref] == [len(f) for f in f_ora]
        d = max([float((a - b).abs().max()) for a, b in zip(o_ref, o_ora)] +
                [float((a - b).abs().max()) for fa, fb in zip(f_ref, f_ora) for a, b in zip(fa, fb)])
        print(f"{kind}: |reference - restatement|max over {len(o_ref)} scores and {sum(len(f) for f in f_ref)} "
              f"feature maps = {d:.3e}")
        assert d <= 1e-6
        for i, o in enumerate(o_ref):
            save[f"{kind}.out{i}"] = o.numpy()
            for j, fm in enumerate(f_ref[i]):
                flat = fm.reshape(-1)
                pick = torch.linspace(0, flat.numel() - 1, min(512, flat.numel())).long()
                save[f"{kind}.f{i}.{j}.shape"] = np.array(fm.shape)
                save[f"{kind}.f{i}.{j}.idx"] = pick.numpy()
                save[f"{kind}.f{i}.{j}.val"] = flat[pick].numpy()
                save[f"{kind}.f{i}.{j}.sum"] = np.array([float(flat.double().sum()), float(flat.double().abs().sum())])
    np.savez_compressed(os.path.join(GOLD, "critics_b2_t2403.npz"), **save)
    print(f"  critics_b2_t2403.npz: {os.path.getsize(os.path.join(GOLD, 'critics_b2_t2403.npz')) / 1024:.1f} KiB")

    # ---- critics, backward (the critic half of the training step, vocoder7/trainer.py:86-115): autograd
    # through the REFERENCE classes vs autograd through the restatement, in .eval() (stored u, v) and in .train() (power
    # iteration first, u / v constants of the graph); golden gradients for tests/test_critics_cpu.py ---------------------
    def gan_like_loss(outs, feats):
        loss = 0.0
        for o in outs:
            loss = loss + ((o - 1.0) ** 2).mean()
        for fs in feats:
            for j, f in enumerate(fs):
                loss = loss + (0.5 + 0.1 * j) * f.abs().mean()
        return loss

    gsave = {"x": x.numpy()}
    for kind, cls in (("mpd", rdisc.MultiPeriodDiscriminator), ("msd", rdisc.MultiScaleDiscriminator),
                      ("mbd", rdisc.MultiBandDiscriminator)):
        for mode in ("eval", "train"):
            torch.manual_seed(1234)
            ref = cls(cfg_ref).train(mode == "train")
            xr = x.clone().requires_grad_(True)
            gan_like_loss(*ref(xr)).backward()
            sd_ora = O.make_critic_state(kind, cfg_ref, seed=1234)
            for k in sd_ora:
                if k.endswith("weight_orig") or k.endswith("bias"):
                    sd_ora[k].requires_grad_(True)
            xo = x.clone().requires_grad_(True)
            gan_like_loss(*O.critic_forward(kind, sd_ora, cfg_ref, xo, training=(mode == "train"))).backward()
            worst = float((xr.grad - xo.grad).abs().max() / xr.grad.abs().max())
            grads = {"x": xr.grad}
            for name, prm in ref.named_parameters():
                worst = max(worst, float((prm.grad - sd_ora[name].grad).abs().max() / prm.grad.abs().max()))
                grads[name] = prm.grad
            print(f"{kind} backward ({mode}): |reference autograd - restatement autograd|max / scale = {worst:.3e}")
            assert worst <= 1e-5
            for name, gr in grads.items():
                flat = gr.detach().reshape(-1)
                pick = torch.linspace(0, flat.numel() - 1, min(64, flat.numel())).long()
                gsave[f"{kind}.{mode}.{name}.idx"] = pick.numpy()
                gsave[f"{kind}.{mode}.{name}.val"] = flat[pick].numpy()
                gsave[f"{kind}.{mode}.{name}.sum"] = np.array([float(flat.double().sum()), float(flat.double().abs().sum()),
                                                              float(flat.abs().max())])
    np.savez_compressed(os.path.join(GOLD, "critics_grad_b2_t2403.npz"), **gsave)
    print(f"  critics_grad_b2_t2403.npz: {os.path.getsize(os.path.join(GOLD, 'critics_grad_b2_t2403.npz')) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
