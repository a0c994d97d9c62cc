"""Run the FFT-400 unit routines (csrc/fft400.cuh) thread by thread on the CPU against a float64 DFT."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fft400_unit_math_on_host(tmp_path):
    exe = tmp_path / "fft400_host_check"
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", str(exe),
                           os.path.join(ROOT, "tests", "host", "fft400_host_check.cu")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.strip().endswith("OK")
