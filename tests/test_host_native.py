"""Host-side C++ of the library checked without a GPU: ``tests/native/amg_host_check.cu`` includes
``csrc/amg_setup.cpp`` directly (its static helpers become visible), is compiled with nvcc as host code and run."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
@pytest.mark.parametrize("name,expect", [("amg_host_check", "coarse inverse residual"),
                                         ("sell_format_check", "sell format check: 0 failures"),
                                         ("halo_geom_check", "halo geometry check: 0 failures")])
def test_host_native_checks(tmp_path, name, expect):
    """amg_host_check: pivoted dense inverse + one complete hierarchy; sell_format_check: every exact matrix format
    read back with the kernels' own decode functions (csrc/sell_format.h); halo_geom_check: all ranks of a partition
    played in one process, distributed V-cycle through the exchange geometry against the serial cycle (csrc/halo_geom.h)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / name)
    subprocess.check_call([nvcc, "-std=c++17", "-O2", "-x", "cu", "-w", "-I", os.path.join(ROOT, "control_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "native", name + ".cu")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert expect in out.stdout
