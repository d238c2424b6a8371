"""Drives the tile kernels' per-thread strip arithmetic (pde_surrogate_b200/csrc/stencil_core.cuh)
thread-by-thread on the CPU and checks it against the C oracle.  No GPU needed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "host_emul")


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_stencil_strip_emulation():
    bld = os.path.join(EMUL, "_build")
    os.makedirs(bld, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-c", os.path.join(ROOT, "oracle", "darcy_oracle.c"), "-o",
                           os.path.join(bld, "darcy_oracle.o")])
    subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-c",
                           os.path.join(EMUL, "stencil_emul.cu"), "-o", os.path.join(bld, "stencil_emul.o")])
    subprocess.check_call(["nvcc", "-Wno-deprecated-gpu-targets", os.path.join(bld, "stencil_emul.o"),
                           os.path.join(bld, "darcy_oracle.o"), "-o", os.path.join(bld, "stencil_emul")])
    r = subprocess.run([os.path.join(bld, "stencil_emul")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "EMUL PASSED" in r.stdout
