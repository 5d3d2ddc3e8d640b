"""The tagged 16-byte partial of the persistent kernel's exchange, checked on the CPU: tests/host/pay16_check.cu includes the
header the kernel is compiled from (tfdiffeq_b200/csrc/b2ode_pay16.cuh, __host__ __device__ functions) and runs pack / validate
/ unpack round trips, stale and torn messages, cleared memory and the poison pattern.  Needs nvcc only (no GPU)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="needs nvcc")
def test_pay16_round_trips_on_the_host(tmp_path):
    exe = str(tmp_path / "pay16_check")
    subprocess.run([NVCC, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "tfdiffeq_b200", "csrc"), "-o", exe,
                    os.path.join(HERE, "host", "pay16_check.cu")], check=True, timeout=300)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout
    assert "pay16: ok" in out.stdout
