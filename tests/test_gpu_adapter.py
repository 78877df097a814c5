"""Runs the C++ drop-in adapter (reference-shaped FastCorrelativeScanMatcher2D/3D,
RealTime…2D and ConstraintBuilder2D/3D classes over the C ABI) on the device."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "cartographer_b200", "adapter")


@pytest.mark.gpu
def test_adapter_selftest_runs_on_device():
    exe = os.path.join(ADAPTER, "adapter_selftest")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ADAPTER, "-s"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "link check only" not in out.stdout, "no CUDA device seen by the adapter"
    assert "ConstraintBuilder2D 1 callback(s), 2 constraints" in out.stdout
    assert "ConstraintBuilder3D" in out.stdout


def test_adapter_headers_compile_and_link():
    """CPU: the adapter builds against the stand-in headers and links the C ABI; without
    a device the binary reports that and exits 0 (no CPU fallback to exercise)."""
    subprocess.check_call(["make", "-C", ADAPTER, "-s"])
    out = subprocess.run([os.path.join(ADAPTER, "adapter_selftest")], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    # the builder's gating / WhenDone cycle and the sampler sequence run without a device
    assert "host-only checks passed" in out.stdout
