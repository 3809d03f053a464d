"""CPU: the Stockham FFT building blocks (csrc/fft_core.cuh) are __host__ __device__; this compiles the
host-side harness with nvcc (no GPU needed) and checks index math / twiddles / real-FFT split
against a naive double-precision DFT."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_fft_core_on_host():
    src = os.path.join(ROOT, "tests", "host", "fft_core_host.cu")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "fft_core_host")
        subprocess.run(["nvcc", "-O1", "-std=c++17", "-w", "-o", exe, src], check=True, capture_output=True)
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "FAIL" not in out.stdout and out.stdout.count("ok") >= 10


def test_attention_exp2_polynomial_constants():
    """csrc/attention.cu ex2_poly3 (every 6th exponential of the attention softmax runs on the FMA pipe): the same fp32
    operations in numpy -- clamp, magic-number rounding, degree-3 Horner, exponent-field add -- against 2^x over the whole
    range the kernel can see.  Bound: 8e-5 relative (the polynomial's 7.5e-5 + fp32 rounding), a sixth of the 16-bit
    rounding P gets right after."""
    import re
    import numpy as np
    src = open(os.path.join(ROOT, "tts-core-remastered-1_b200", "csrc", "attention.cu")).read()
    body = src[src.index("float ex2_poly3(float x)"):]
    body = body[:body.index("\n}\n")]
    c3, c2, c1, c0 = (np.float32(v) for v in re.findall(r"(0\.\d+)f", body)[:4])
    assert "12582912.0f" in body and "-126.0f" in body
    x = np.concatenate([np.linspace(-126.0, 0.0, 400001), -np.logspace(-8, 2, 20001), [-200.0, -1e30]]).astype(np.float32)
    xc = np.maximum(x, np.float32(-126.0))
    t = (xc + np.float32(12582912.0)).astype(np.float32)
    f = (xc - (t - np.float32(12582912.0)).astype(np.float32)).astype(np.float32)
    assert float(np.abs(f).max()) <= 0.5
    q = (c3 * f + c2).astype(np.float32)
    q = (q * f + c1).astype(np.float32)
    q = (q * f + c0).astype(np.float32)
    bits = (q.view(np.int32).astype(np.int64) + ((t.view(np.int32).astype(np.int64) << 23) & 0xFFFFFFFF)) & 0xFFFFFFFF
    p = bits.astype(np.uint32).view(np.float32)
    ref = np.exp2(xc.astype(np.float64))
    ok = xc > -125.0                                    # above the last binade the result is a normal number
    rel = np.abs(p[ok].astype(np.float64) / ref[ok] - 1.0)
    assert float(rel.max()) <= 8e-5, float(rel.max())
    assert float(np.abs(p[~ok]).max()) <= 2.0 ** -124   # clamped tail: (sub)normal dust that a 16-bit P rounds to zero
