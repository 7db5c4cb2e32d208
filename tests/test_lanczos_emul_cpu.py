"""CPU tier: the DEVICE code of kry_lanczos_diag / kry_lanczos_diag_dist (krypy_b200/csrc/kry_lanczos_kernels.cuh:
the whole Lanczos step for a diagonal inner-product matrix in ONE cooperative kernel -- BASELINE config C5, MINRES
with ip_B, krypy/utils.py:1000-1045), compiled unchanged for the host over the CUDA execution emulator
(tests/csrc/cuda_emul), single rank and row-partitioned over 2 and 3 emulated ranks, against a long-double
reference with the kernel's rounding points (B q is rounded to the storage type before it enters a dot, as the
generic seven-launch path stores it; tests/csrc/lanczos_emul_host.cpp)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "lanczos_emul_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread",
                           "-I", os.path.join(HERE, "csrc", "cuda_emul"), "-I", os.path.join(ROOT, "krypy_b200", "csrc"),
                           "-o", out, os.path.join(HERE, "csrc", "lanczos_emul_host.cpp")])

    def run(*args):
        p = subprocess.run([out] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0 and p.stdout.startswith("ok") and "ranks identical 1" in p.stdout, (args, p.stdout, p.stderr)
        return p.stdout
    return run


@pytest.mark.parametrize("dtype,vec", [("f64", 2), ("f64", 1), ("f32", 4), ("f32", 1)])
def test_lanczos_diag_kernel_emulated(emul, dtype, vec):
    """(dtype, vector width, rows, CTAs, ranks, pre-subtraction, normalised store): first step (no v_{k-1}),
    later steps, odd lengths (scalar tail), one to three CTAs"""
    emul(dtype, vec, 3001, 2, 1, 1, 1)
    emul(dtype, vec, 1000, 3, 1, 0, 1)
    emul(dtype, vec, 7, 1, 1, 1, 0)


@pytest.mark.parametrize("ranks", [2, 3])
def test_lanczos_diag_kernel_emulated_row_partitioned(emul, ranks):
    """both reductions are completed across the emulated ranks inside the kernel: alpha and beta bitwise identical
    on all ranks, two epochs per step"""
    emul("f64", 2, 1500, 2, ranks, 1, 1)
    emul("f32", 4, 2001, 2, ranks, 0, 1)
