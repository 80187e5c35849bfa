"""Runs the C++ host mirror's known-answer tests (krabmaga_b200/host/host_kat.cpp) on the device."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_host_mirror_known_answers():
    exe = os.path.join(ROOT, "krabmaga_b200", "host", "host_kat")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "-s"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all passed" in out.stdout


def test_cpp_host_mirror_builds_and_fails_loudly_without_a_device():
    exe = os.path.join(ROOT, "krabmaga_b200", "host", "host_kat")
    subprocess.check_call(["make", "-C", os.path.dirname(exe), "-s"])
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 2 and "no CUDA device" in out.stdout
