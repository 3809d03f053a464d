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
