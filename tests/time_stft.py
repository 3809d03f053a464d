"""Development script (GPU): CUDA-event timing of the STFT family at BASELINE configs[2] (1024 x 4 s, n_fft 1024, hop 256);
run it under different B200VOC_STFT_* switches for A/B."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tts-core-remastered-1_b200")]
import b200voc  # noqa: E402

x = torch.rand(1024, 88200, device="cuda") * 2 - 1
b200voc.stft_prepare(1024, 80, 22050)


def timeit(fn, reps=10):
    a = fn(); b = fn(); del a, b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


ms_lm, lm = timeit(lambda: b200voc.log_mel(x))
ms_mag, _ = timeit(lambda: b200voc.stft_magnitude(x, 1024, 256))
ms_c, sp = timeit(lambda: b200voc.stft(x, 1024, 256))
ms_i, y = timeit(lambda: b200voc.istft(sp, 1024, 256, 88200))
print(f"logmel {ms_lm:.3f}  mag {ms_mag:.3f}  complex {ms_c:.3f}  istft {ms_i:.3f} ms   checksum {float(lm.double().mean()):.9f} "
      f"roundtrip {float((y - x).abs().max()):.2e}", flush=True)
