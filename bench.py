"""bench.py -- vocoder7 Generator synthesis throughput (audio-seconds per second) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one Generator.forward over one batch of synthetic mels (BASELINE.json configs[1]:
batch 16 x 10 s, T = 861 frames, default GANConfig, hidden_dim 512, random-init weights).
For N > 1 the driver launches one rank per GPU with torch.distributed.run; every rank synthesises
its own 16 x 10 s shard (independent utterances, no data-path collective -> "weak" scaling), time
is the max over ranks, value is the whole-job audio-s/s.

`--impl reference` times the reference path (the CPU oracle port of vocoder7/generator.py, see
oracle/vocoder7_oracle.py) on the host cores with all threads, on a bounded sample of the same
workload.  It is the only place besides cpu_baseline where oracle/ is executed.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tts-core-remastered-1_b200"))

METRIC = "vocoder7_audio_seconds_per_second"
UNIT = "audio-s/s"
B_PER_GPU, T_FRAMES, SR, HOP = 16, 861, 22050, 256


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt, self.proc = index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append(line.strip())
                if self._stop_evt.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = max(smax, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_throughput(sample_B: int, sample_T: int, reps: int, threads: int):
    """Reference path on the host cores: the oracle port of Generator.forward, fp32, all threads."""
    import torch
    from oracle import vocoder7_oracle as O
    torch.set_num_threads(threads)
    cfg = O.OracleConfig(use_attention=False)
    gen = O.make_generator(cfg, seed=1234)
    sd = gen.state_dict()
    mel, pros, sty, emo = O.synthetic_inputs(sample_B, sample_T, seed=4321)
    with torch.no_grad():
        O.generator_forward(sd, cfg, mel, pros, sty, emo)          # warm-up
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            O.generator_forward(sd, cfg, mel, pros, sty, emo)
            times.append(time.perf_counter() - t0)
    audio_s = sample_B * HOP * sample_T / SR
    return audio_s / min(times), times


def library_gpu_throughput(dev, B: int, T: int, reps: int = 3):
    """SURVEY.md 8(d) asks for the 'library path' bar next to the CPU baseline: the same oracle module run by
    torch eager on the GPU (cuDNN / cuBLAS kernels, bf16 autocast as vocoder7/trainer.py:75 runs the reference).
    Part of the cpu_baseline leg: a reported baseline, never the product path.  Returns (audio-s/s, ms)."""
    import torch
    from oracle import vocoder7_oracle as O
    gen = O.make_generator(O.OracleConfig(use_attention=False), seed=1234).to(dev)
    ins = [t.to(dev) for t in O.synthetic_inputs(B, T, seed=4321)]
    cuda = torch.device(dev).type == "cuda"
    with torch.no_grad(), torch.autocast(torch.device(dev).type, dtype=torch.bfloat16):
        gen(*ins)                                                   # warm-up (cuDNN heuristics, allocator)
        if cuda:
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        t0 = time.perf_counter()
        for _ in range(reps):
            gen(*ins)
        if cuda:
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
        else:
            ms = 1e3 * (time.perf_counter() - t0) / reps
    del gen, ins
    if cuda:
        torch.cuda.empty_cache()
    return B * HOP * T / SR / (ms / 1e3), ms


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (the oracle port of vocoder7/generator.py --
    the reference as shipped cannot be imported) on the box's host cores with all threads.  Exactly
    --warmup untimed and --steps timed steps; a step is a bounded sample of the workload (1 of the 16 utterances,
    T = 861), throughput is per audio-second so the ratio to the GPU arm stands."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    import torch
    from oracle import vocoder7_oracle as O
    torch.set_num_threads(cores)
    cfg = O.OracleConfig(use_attention=False)
    sd = O.make_generator(cfg, seed=1234).state_dict()
    mel, pros, sty, emo = O.synthetic_inputs(1, T_FRAMES, seed=4321)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    with torch.no_grad():
        for _ in range(warm):
            O.generator_forward(sd, cfg, mel, pros, sty, emo)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.generator_forward(sd, cfg, mel, pros, sty, emo)
        dt = time.perf_counter() - t0
    audio_s = 1 * HOP * T_FRAMES / SR
    value = audio_s * steps / dt
    sample = (f"each step = 1 utterance x T={T_FRAMES} (10 s) of the B=16 workload, fp32, torch CPU oracle port of "
              f"vocoder7/generator.py, {cores} threads, attention off")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {
        "workload": f"vocoder7 Generator forward, default GANConfig + hidden_dim=512, B={B_PER_GPU} x T={T_FRAMES} "
                    f"(10 s) per GPU, random-init weights (seed 1234), synthetic randn mels",
        "batch_per_gpu": B_PER_GPU, "frames": T_FRAMES, "audio_seconds_per_step_per_gpu": B_PER_GPU * HOP * T_FRAMES / SR,
        "precision_plan": "fp16 operands (tcgen05 kind::f16: the same instruction and rate as bf16), fp32 accumulate; "
                          "BASELINE configs[1] says bf16 -- plain bf16 operands cannot meet the 1e-3 max-abs gate on this "
                          "network (SURVEY D4), the bf16 plan is reported under `precision_plans`",
        "attention": "off for the headline value (the conv hot path north_star names); the same step with the "
                     "SelfAttention layer on (builder-defined D3) is reported under with_attention (global) and "
                     "with_attention_windowed (attn_window 4096)",
        "l2": "per-layer activations (0.9 GB) exceed the 126 MB L2; no explicit flush",
        "parallelism": f"dp{n_gpus} (independent utterance shards, no collective)",
    }


def secondary_measurements(dev, dev_in, B, T):
    """Not the headline: (a) the same workload with the SelfAttention layer evaluated (global
    attention); (b) BASELINE configs[2], the STFT family on 1024 x 4 s waveforms."""
    import torch
    import b200voc
    from b200voc import GANConfig, Generator, _lib
    lib = _lib.load()
    peaks = measured_peaks()
    out = {}
    try:
        torch.manual_seed(1234)
        gen_a = Generator(GANConfig(use_attention=True)).eval().to(dev)
        with torch.no_grad():
            gen_a(*dev_in)
            torch.cuda.synchronize()
            lib.b200voc_gen_profile_enable(gen_a._handle, 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_att = 3
            e0.record()
            for _ in range(n_att):
                gen_a(*dev_in)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_att
        att = None
        for i in range(lib.b200voc_gen_profile_count(gen_a._handle)):
            if lib.b200voc_gen_profile_name(gen_a._handle, i).decode() == "attn":
                att = dict(ms=float(lib.b200voc_gen_profile_ms(gen_a._handle, i)),
                           flops=float(lib.b200voc_gen_profile_flops(gen_a._handle, i)))
        out["with_attention"] = {
            "value": B * HOP * T / SR / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": n_att,
            "note": "global single-head attention over L=128*T=110208 positions (97 % of all FLOPs); every 6th exponential on the "
                    "FMA pipe (B200VOC_ATTN_POLY); attn_* = the attention kernels of the last step",
            "attn_kernels_ms": att["ms"] if att else None,
            "attn_tflops": att["flops"] / (att["ms"] * 1e-3) / 1e12 if att else None}
        del gen_a
    except Exception as e:  # keep the headline line even if the secondary run fails
        out["with_attention"] = {"error": str(e)[:200]}
    try:
        torch.manual_seed(1234)
        gen_w = Generator(GANConfig(use_attention=True, attn_window=4096)).eval().to(dev)
        with torch.no_grad():
            for _ in range(2):
                gen_w(*dev_in)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                gen_w(*dev_in)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out["with_attention_windowed"] = {
            "value": B * HOP * T / SR / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": 5, "attn_window": 4096,
            "note": "the reference's as-constructed graph (generator.py:42-44) with the SelfAttention layer evaluated "
                    "block-locally over windows of 4096 positions (32 mel frames): the usable, streaming-exact form"}
        del gen_w
    except Exception as e:
        out["with_attention_windowed"] = {"error": str(e)[:200]}
    try:
        # the other precision plans next to the fp16 default: max-abs / SNR against the fp32 oracle on one utterance, ms / step
        from oracle import vocoder7_oracle as O
        ocfg = O.OracleConfig(use_attention=False)
        ora = O.make_generator(ocfg, seed=1234)
        ins1 = O.synthetic_inputs(1, 200, seed=77)
        with torch.no_grad():
            ref = O.generator_forward(ora.state_dict(), ocfg, *ins1)
        plans = {}
        for plan in ("fp16", "mixed", "bf16"):
            gp = Generator(GANConfig(use_attention=False, precision=plan)).eval()
            gp.load_state_dict(ora.state_dict())
            gp = gp.to(dev)
            with torch.no_grad():
                w1 = gp(*[t.to(dev) for t in ins1]).cpu()
                for _ in range(2):
                    gp(*dev_in)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    gp(*dev_in)
                e1.record()
                torch.cuda.synchronize()
            plans[plan] = {"max_abs": float((w1 - ref).abs().max()), "snr_db": float(O.snr_db(ref, w1)),
                           "ms_per_step": e0.elapsed_time(e1) / 5}
            del gp
        out["precision_plans"] = {"gate": "max-abs <= 1e-3 and SNR >= 40 dB vs the fp32 oracle (1 x T=200)", **plans}
    except Exception as e:
        out["precision_plans"] = {"error": str(e)[:200]}
    try:
        # callers / wire formats either side of the path: GlobalStyleTokens on the
        # time-major mel, Generator fed the same time-major mel, 16-bit PCM out with a length mask
        from b200voc import GlobalStyleTokens
        torch.manual_seed(1234)
        gst = GlobalStyleTokens(GANConfig()).eval().to(dev)
        gen_f = Generator(GANConfig(use_attention=False)).eval().to(dev)
        mel_btc = dev_in[0].transpose(1, 2).contiguous()
        lens = torch.full((B,), T, device=dev, dtype=torch.int32)
        pcm = torch.empty(B, 1, HOP * T, device=dev, dtype=torch.int16)

        def pipeline():
            style = gst(mel_btc, mel_layout="BTC")
            return gen_f(mel_btc, dev_in[1], style, dev_in[3], mel_layout="BTC", out_dtype=torch.int16,
                         frame_lengths=lens, out=pcm)
        with torch.no_grad():
            for _ in range(2):
                pipeline()
            torch.cuda.synchronize()
            a, b_, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            for _ in range(5):
                gst(mel_btc, mel_layout="BTC")
            b_.record()
            for _ in range(5):
                pipeline()
            c.record()
            torch.cuda.synchronize()
        out["frontend"] = {
            "gst_ms": a.elapsed_time(b_) / 5, "gst_mel_bytes": int(mel_btc.numel() * 4),
            "mel_btc_gst_generator_pcm16": {"ms_per_step": b_.elapsed_time(c) / 5,
                                             "value": B * HOP * T / SR / (b_.elapsed_time(c) / 5 / 1e3), "unit": UNIT},
            "note": "GlobalStyleTokens (vocoder7/gst.py) + Generator on [B,T,80] mels, int16 PCM out, length mask"}
        del gst, gen_f
    except Exception as e:
        out["frontend"] = {"error": str(e)[:200]}
    try:
        # BASELINE configs[0] shape (B = 1, T = 172 = 2 s): launch-bound, so the serving form is one CUDA graph per
        # request shape (b200voc.scheduler.GraphedSynthesizer); eager call beside it
        from b200voc.scheduler import GraphedSynthesizer
        torch.manual_seed(1234)
        gen_s = Generator(GANConfig(use_attention=False)).eval().to(dev)
        g0 = torch.Generator().manual_seed(3)
        small = [torch.randn(1, 80, 172, generator=g0).to(dev), torch.randn(1, 172, 18, generator=g0).to(dev),
                 torch.randn(1, 128, generator=g0).to(dev), torch.softmax(torch.randn(1, 6, generator=g0), -1).to(dev)]
        gs = GraphedSynthesizer(gen_s)
        o_small = torch.empty(1, 1, HOP * 172, device=dev)
        with torch.no_grad():
            for _ in range(3):
                gs(*small)
                gen_s(*small, out=o_small)
            torch.cuda.synchronize()
            a, b_, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            n_small = 50
            a.record()
            for _ in range(n_small):
                gs(*small)
            b_.record()
            for _ in range(n_small):
                gen_s(*small, out=o_small)
            c.record()
            torch.cuda.synchronize()
            same = bool(torch.equal(gs(*small), gen_s(*small)))
        g_ms, e_ms = a.elapsed_time(b_) / n_small, b_.elapsed_time(c) / n_small
        out["small_request"] = {
            "workload": "BASELINE configs[0] shape: 1 x T=172 (2 s), attention off", "graph_ms": g_ms, "eager_ms": e_ms,
            "graph_audio_s_per_s": HOP * 172 / SR / (g_ms / 1e3), "eager_audio_s_per_s": HOP * 172 / SR / (e_ms / 1e3),
            "launches_per_forward": gen_s.launch_count(), "bit_identical": same,
            "note": "one CUDA graph per exact request shape, 4 input copies + 1 graph launch per call"}
        del gs, gen_s
    except Exception as e:
        out["small_request"] = {"error": str(e)[:200]}
    try:
        # SURVEY 8f rank 4 (forward half): the three critics the trainer runs on every waveform
        # (vocoder7/trainer.py:86-92)
        from b200voc import MultiPeriodDiscriminator, MultiScaleDiscriminator, MultiBandDiscriminator
        Bc, Tc = 4, SR
        wavc = torch.rand(Bc, 1, Tc, device=dev) * 2 - 1
        res = {"batch": Bc, "samples": Tc, "note": "critic forwards (vocoder7/discriminators.py) on 4 x 1 s, all "
               "feature maps written; MSD's stride-1 64->256 / 256->1024 layers (97 % of the FLOPs) on tcgen05 with split-bf16 "
               "operands, the narrow / strided layers as fp32 direct convolution; train_fwd_bwd_ms = .train() forward (one power iteration per layer) "
               "+ backward of a loss over every score and feature map down to every weight and the waveform (csrc/disc_bwd.cu: wide "
               "layers' dgrad / wgrad on tcgen05), host-side launch overhead included"}
        for name, cls in (("mpd", MultiPeriodDiscriminator), ("msd", MultiScaleDiscriminator),
                          ("mbd", MultiBandDiscriminator)):
            torch.manual_seed(1234)
            crit = cls(GANConfig()).eval().to(dev)
            with torch.no_grad():
                crit(wavc)
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            with torch.no_grad():
                for _ in range(3):
                    crit(wavc)
            b_.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b_) / 3
            fl = crit.forward_flops(Bc, Tc)
            res[name] = {"ms": ms, "gflop": fl / 1e9, "tflops": fl / (ms * 1e-3) / 1e12}
            # the critic half of a training step (vocoder7/trainer.py:86-115): .train() forward (power iteration) +
            # backward of a loss over every score and feature map down to the weights and the waveform
            try:
                crit.train()
                wg = wavc.clone().requires_grad_(True)

                def d_step():
                    crit.zero_grad(set_to_none=True)
                    with torch.enable_grad():
                        o, f = crit(wg)
                        loss = sum((s_ ** 2).mean() for s_ in o) + sum(m.abs().mean() for fs in f for m in fs)
                        loss.backward()
                for _ in range(3):
                    d_step()
                torch.cuda.synchronize()
                a.record()
                for _ in range(5):
                    d_step()
                b_.record()
                torch.cuda.synchronize()
                ms_t = a.elapsed_time(b_) / 5
                res[name]["train_fwd_bwd_ms"] = ms_t
                res[name]["train_fwd_bwd_tflops"] = 3 * fl / (ms_t * 1e-3) / 1e12
            except Exception as e:
                res[name]["train_fwd_bwd_error"] = str(e)[:200]
            # the same step captured in ONE CUDA graph (forward, loss, backward; static input): what the narrow critics need,
            # their eager step is bound by ~600 launches from the host
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):
                        d_step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                crit.zero_grad(set_to_none=True)
                wg.grad = None
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph):
                    with torch.enable_grad():
                        o, f = crit(wg)
                        loss = sum((s_ ** 2).mean() for s_ in o) + sum(m.abs().mean() for fs in f for m in fs)
                        loss.backward()
                gph.replay()
                torch.cuda.synchronize()
                a.record()
                for _ in range(5):
                    gph.replay()
                b_.record()
                torch.cuda.synchronize()
                res[name]["train_fwd_bwd_graph_ms"] = a.elapsed_time(b_) / 5
                del gph
            except Exception as e:
                res[name]["train_fwd_bwd_graph_error"] = str(e)[:200]
            del crit
        out["critics"] = res
    except Exception as e:
        out["critics"] = {"error": str(e)[:200]}
    try:
        Bw, Nw = 1024, 88200
        x = torch.rand(Bw, Nw, device=dev) * 2 - 1
        res = {}

        def timeit(fn, reps=5):
            # two LIVE warm-up results: the loop below holds the previous output while the next one is allocated, and a
            # 0.7 GB cudaMalloc inside the timed region would be measured as kernel time
            w0 = fn()
            w1 = fn()
            del w0, w1
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                r = fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps, r
        frames = 1 + Nw // 256
        ms, lm = timeit(lambda: b200voc.log_mel(x))
        by = Bw * Nw * 4 + Bw * 80 * frames * 4
        res["stft_logmel"] = {"ms": ms, "algorithmic_gb": by / 1e9, "gbs": by / ms / 1e6, "frac_of_hbm": by / ms / 1e6 / peaks["hbm_gbs"]}
        ms, mg = timeit(lambda: b200voc.stft_magnitude(x, 1024, 256))
        by = Bw * Nw * 4 + Bw * 513 * frames * 4
        res["stft_mag"] = {"ms": ms, "algorithmic_gb": by / 1e9, "gbs": by / ms / 1e6, "frac_of_hbm": by / ms / 1e6 / peaks["hbm_gbs"]}
        del mg
        ms, sp = timeit(lambda: b200voc.stft(x, 1024, 256))
        by = Bw * Nw * 4 + Bw * 513 * frames * 8
        res["stft_complex"] = {"ms": ms, "algorithmic_gb": by / 1e9, "gbs": by / ms / 1e6, "frac_of_hbm": by / ms / 1e6 / peaks["hbm_gbs"]}
        ms, y = timeit(lambda: b200voc.istft(sp, 1024, 256, Nw))
        res["istft"] = {"ms": ms, "algorithmic_gb": by / 1e9, "gbs": by / ms / 1e6, "frac_of_hbm": by / ms / 1e6 / peaks["hbm_gbs"],
                        "roundtrip_maxabs": float((y - x).abs().max())}
        # training-side consumer: STFTLoss forward + backward, 3 resolutions, 256 x 4 s
        try:
            lm_mod = b200voc.STFTLoss(GANConfig()).to(dev)
            xf = x[:256].clone().requires_grad_(True)
            xr = torch.rand(256, Nw, device=dev) * 2 - 1

            def fb():
                with torch.enable_grad():
                    lm_mod(xf, xr).backward()
                return xf.grad
            ms, _ = timeit(fb, reps=3)
            res["stft_loss_fwd_bwd"] = {"ms": ms, "batch": 256, "audio_seconds": 256 * Nw / SR,
                                        "note": "STFTLoss(cfg) forward + backward (wav and gain gradients), n_fft 512/1024/2048"}
        except Exception as e:
            res["stft_loss_fwd_bwd"] = {"error": str(e)[:200]}
        res["audio_seconds"] = Bw * Nw / SR
        res["workload"] = "BASELINE configs[2]: 1024 x 4 s uniform(-1,1) waveforms, n_fft 1024, hop 256, 80 HTK mels"
        out["stft"] = res
    except Exception as e:
        out["stft"] = {"error": str(e)[:200]}
    return out


def multi_gpu_jobs(gen, dev, rank, world, barrier):
    """BASELINE configs[3] and [4], outside the headline timing, strong-scaled over the N ranks (N = 1 gives the
    single-GPU figure of the same jobs):
      sharded512     512 x 10 s utterances as pinned HOST tensors, sharded over the ranks, every rank streaming its
                     shard host -> device -> Generator -> host (b200voc.scheduler.sharded_synthesize_streaming); no
                     collective.  The optional final gather (all 451 MB of waveforms to rank 0, ragged point-to-point,
                     scheduler.gather_flat) is timed on its own.
      longform64x60  64 x 60 s (T = 5167) cut into 512-frame chunks + 8-frame halo = 704 independent units spread over
                     the ranks (scheduler.sharded_synthesize_long: an utterance spans GPUs), then the gather.
    Times are wall clock between barriers + device synchronisation, max over ranks (every rank waits at the barrier)."""
    import torch
    import torch.distributed as dist
    from b200voc import scheduler as S
    res = {}

    def timed(fn):
        barrier()
        t0 = time.perf_counter()
        r = fn()
        barrier()
        return time.perf_counter() - t0, r

    def gather_ms(flat, sizes):
        if world == 1:
            return None
        S.gather_flat(flat, sizes, 0)                       # warm-up (NCCL connection set-up)
        dt, _ = timed(lambda: S.gather_flat(flat, sizes, 0))
        return dt * 1e3

    try:
        n, T = 512, T_FRAMES
        g = torch.Generator().manual_seed(99)
        per = (n + world - 1) // world
        lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
        # every rank materialises only its own shard of the host work list (the other rows are never touched)
        mels = torch.empty(n, 80, T).pin_memory()
        pros = torch.empty(n, T, 18).pin_memory()
        mels[lo:hi].normal_(generator=g)
        pros[lo:hi].normal_(generator=g)
        sty = torch.randn(n, 128, generator=g).pin_memory()
        emo = torch.softmax(torch.randn(n, 6, generator=g), -1).pin_memory()
        host_wavs = torch.empty(hi - lo, 1, HOP * T).pin_memory()       # allocated once, as a serving loop would
        S.sharded_synthesize_streaming(gen, dev, mels, pros, sty, emo, max_batch=B_PER_GPU, out=host_wavs)   # warm-up
        dt, (lo2, wavs) = timed(lambda: S.sharded_synthesize_streaming(gen, dev, mels, pros, sty, emo, max_batch=B_PER_GPU,
                                                                       out=host_wavs))
        audio = n * HOP * T / SR
        sizes = [(min(n, (r + 1) * per) - min(n, r * per)) * HOP * T for r in range(world)]
        gms = gather_ms(wavs.to(dev).reshape(-1), sizes)
        res["sharded512"] = {"value": audio / dt, "unit": UNIT, "seconds": dt, "utterances": n, "frames": T,
                             "per_rank": hi - lo, "h2d_bytes": int((hi - lo) * (80 * T + 18 * T + 134) * 4),
                             "d2h_bytes": int((hi - lo) * HOP * T * 4), "collective_in_timed_region": "none",
                             "gather_to_rank0_ms": gms, "gather_bytes": int(sum(sizes[1:]) * 4) if world > 1 else 0}
        del mels, pros, wavs
    except Exception as e:
        res["sharded512"] = {"error": str(e)[:300]}
    try:
        Bl, Tl = 64, 5167
        g = torch.Generator().manual_seed(7)
        mel = torch.randn(Bl, 80, Tl, generator=g).to(dev)
        pr = torch.randn(Bl, Tl, 18, generator=g).to(dev)
        st = torch.randn(Bl, 128, generator=g).to(dev)
        em = torch.softmax(torch.randn(Bl, 6, generator=g), -1).to(dev)
        run = lambda: S.sharded_synthesize_long(gen, mel, pr, st, em, chunk_frames=512, halo=8, max_batch=B_PER_GPU)
        run()
        dt, (units, pieces) = timed(run)
        all_units = S.long_units(Bl, Tl, 512, 8)
        shards = S.plan_shards([u[2] - u[1] for u in all_units], world)
        sizes = [sum(HOP * (all_units[k][4] - all_units[k][3]) for k in sh) for sh in shards]
        gms = gather_ms(torch.cat([x.reshape(-1) for x in pieces]), sizes)
        audio = Bl * HOP * Tl / SR
        res["longform64x60"] = {"value": audio / dt, "unit": UNIT, "seconds": dt, "utterances": Bl, "frames": Tl,
                                "chunk_frames": 512, "halo_frames": 8, "units": len(all_units), "units_this_rank": len(units),
                                "recompute_overhead": sum(u[2] - u[1] for u in all_units) / (Bl * Tl) - 1.0,
                                "collective_in_timed_region": "none", "gather_to_rank0_ms": gms,
                                "gather_bytes": int(sum(sizes[1:]) * 4) if world > 1 else 0}
        del mel, pr, pieces
    except Exception as e:
        res["longform64x60"] = {"error": str(e)[:300]}
    torch.cuda.empty_cache()
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist
    from b200voc import GANConfig, Generator, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = GANConfig(use_attention=False)
    torch.manual_seed(1234)
    gen = Generator(cfg).eval().to(dev)
    g = torch.Generator().manual_seed(4321 + rank)
    B, T = B_PER_GPU, T_FRAMES
    host = [torch.randn(B, 80, T, generator=g).pin_memory(), torch.randn(B, T, 18, generator=g).pin_memory(),
            torch.randn(B, 128, generator=g).pin_memory(),
            torch.softmax(torch.randn(B, 6, generator=g), -1).pin_memory()]
    dev_in = [h.to(dev) for h in host]
    out = torch.empty(B, 1, HOP * T, device=dev)
    host_out = torch.empty(B, 1, HOP * T).pin_memory()
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            gen(*dev_in, out=out)
        barrier()
        # ---------------- kernel-resident timing (inputs already in HBM) ----------------
        lib.b200voc_gen_profile_enable(gen._handle, 1)
        sampler = ClockSampler(local)
        sampler.start()
        time.sleep(0.3)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per_layer = {}
        barrier()
        ev0.record()
        for _ in range(args.steps):
            gen(*dev_in, out=out)
        ev1.record()
        barrier()
        ms_total = ev0.elapsed_time(ev1)
        # per-launch events of the last timed step
        n = lib.b200voc_gen_profile_count(gen._handle)
        for i in range(n):
            nm = lib.b200voc_gen_profile_name(gen._handle, i).decode()
            per_layer[nm] = dict(ms=float(lib.b200voc_gen_profile_ms(gen._handle, i)),
                                 flops=float(lib.b200voc_gen_profile_flops(gen._handle, i)),
                                 bytes=float(lib.b200voc_gen_profile_bytes(gen._handle, i)))
        lib.b200voc_gen_profile_enable(gen._handle, 0)
        # ---------------- end-to-end: pinned host -> device -> forward -> pinned host ----------------
        # through the package's host-to-host serving loop (b200voc.scheduler.StreamingSynthesizer):
        # every step uploads its own inputs and downloads its own waveforms inside the timed region;
        # the copies of neighbouring steps overlap the kernels (double-buffered device tensors).
        from b200voc.scheduler import StreamingSynthesizer
        host_outs = [torch.empty(B, 1, HOP * T).pin_memory() for _ in range(2)]
        streamer = StreamingSynthesizer(gen, dev, depth=2)
        streamer.run([host] * 2, host_outs)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        streamer.run([host] * args.steps, [host_outs[i % 2] for i in range(args.steps)])
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
        host_out = host_outs[0]
        sampler.stop()
        # ---------------- soak: the same step back to back for >= 2 s (the "sustained" regime the peak refers to) ------
        soak_steps = max(args.steps, int(2.2e3 / (ms_total / args.steps)))
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(soak_steps):
            gen(*dev_in, out=out)
        s1.record()
        barrier()
        ms_soak = s0.elapsed_time(s1)
        # ---------------- BASELINE configs[3] / [4]: sharded batch and long-form jobs over all ranks ------------------
        jobs = multi_gpu_jobs(gen, dev, rank, world, barrier)
        extra = {}
        if rank == 0:
            extra = secondary_measurements(dev, dev_in, B, T)

    if world > 1:
        t = torch.tensor([ms_total, ms_e2e, ms_soak], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, ms_soak = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    audio_s_step = world * B * HOP * T / SR
    value = audio_s_step * args.steps / (ms_total / 1e3)
    e2e_value = audio_s_step * args.steps / (ms_e2e / 1e3)
    peaks = measured_peaks()
    # dominant kernel: the fused residual-block kernel at C=128 (stage 1), 3 launches per step
    dom = [v for k, v in per_layer.items() if k.startswith("res1.")]
    dom_ms = sum(v["ms"] for v in dom) / max(len(dom), 1)
    dom_flops = dom[0]["flops"] if dom else 0.0
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    total_flops = sum(v["flops"] for v in per_layer.values())
    conv_ms = sum(v["ms"] for v in per_layer.values())
    traffic_path = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    traffic = None
    if os.path.exists(traffic_path):
        try:
            traffic = json.load(open(traffic_path)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    cores = os.cpu_count() or 1
    cpu_val, cpu_times = cpu_oracle_throughput(1, T_FRAMES, reps=16, threads=cores)   # ~10 s of CPU work; batch 1 is the CPU path's best case
    library = None
    if world == 1:
        try:   # the library-path bar of SURVEY 8(d); never allowed to cost the headline line
            lib_val, lib_ms = library_gpu_throughput(dev, B_PER_GPU, T_FRAMES)
            library = {"value": lib_val, "unit": UNIT, "ms_per_step": lib_ms, "dtype": "bf16 autocast",
                       "what": "the oracle module under torch eager on this GPU (cuDNN / cuBLAS), same B x T, attention off"}
        except Exception as e:
            library = {"error": str(e)[:200]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": workload_config(world),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "kernel": "resblock3_kernel<128,...,PAIR> (stage-1 fused residual block on CTA pairs, tcgen05.mma.cta_group::2: conv k3 + GLU + FiLM + 1x1 + residual, 3 launches/step)",
                     "peak_source": f"{peaks['source']} bf16 sustained (kernel timed inside a long step)",
                     "flops_per_launch": dom_flops, "ms_per_launch": dom_ms,
                     "share_of_step": (sum(v["ms"] for v in dom) / conv_ms) if conv_ms > 0 else None,
                     # the longest SINGLE launch is the fused last stage (instruction / shared-memory-pipe bound)
                     "longest_single_launch": (lambda v: {"kernel": "stage_fused_kernel<32,...> (stage 3: ConvT + 3 residual blocks + band_merge + tanh)",
                                                          "ms": v["ms"], "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12,
                                                          "frac_of_tensor_peak": v["flops"] / (v["ms"] * 1e-3) / 1e12 / peak if peak else None,
                                                          "bound": "shared-memory data pipe (ncu: LSU 53 % + tensor-core operand reads 46 %) and instruction issue (62 %)"})(
                         per_layer["stage3+merge"]) if "stage3+merge" in per_layer and per_layer["stage3+merge"]["ms"] > 0 else None},
        "whole_step": {"algorithmic_tflop": total_flops / 1e12, "sum_kernel_ms": conv_ms,
                       "tflops_over_step": total_flops / (ms_total / args.steps * 1e-3) / 1e12 * 1.0,
                       "frac_of_tensor_peak": total_flops / (ms_total / args.steps * 1e-3) / 1e12 / peak},
        "layers": {k: {"ms": round(v["ms"], 4),
                       "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["ms"] > 0 else None,
                       "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None}
                   for k, v in per_layer.items()},
        "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"best of {len(cpu_times)} forwards of 1 utterance x T={T_FRAMES} (10 s) of the same workload, "
                                   f"fp32 torch CPU oracle, attention off; {['%.2f' % t for t in cpu_times]} s",
                         "library_gpu": library},
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(sum(h.numel() * 4 for h in host)),
                "d2h_bytes_per_step": int(host_out.numel() * 4), "ms_per_step": ms_e2e / args.steps},
        "soak": {"value": audio_s_step * soak_steps / (ms_soak / 1e3), "unit": UNIT, "steps": soak_steps,
                 "seconds": ms_soak / 1e3, "ms_per_step": ms_soak / soak_steps},
        **jobs,
        "gpu_launches": gen.launch_count() * args.steps,
        "clocks": sampler.summary(),
        **extra,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
