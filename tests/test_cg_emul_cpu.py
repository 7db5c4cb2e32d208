"""CPU tier: a Jacobi-preconditioned CG iteration chain (krypy/linsys.py:593-689; BASELINE config C3) with the DEVICE
code of its kernels, compiled unchanged for the host over the CUDA execution emulator (tests/csrc/cuda_emul,
tests/csrc/cg_emul_host.cpp): SpMV with the <p, Ap> epilogue -> kry_cg_update_dev -> kry_cg_scalars -> kry_xpby_dev,
the scalars of the recurrence never leaving "device" memory, against a long-double CG (residual history, alpha,
the iterate); and kry_cg_scalars' global sum over emulated ranks."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "cg_emul_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread",
                           "-I", os.path.join(HERE, "csrc", "cuda_emul"), "-I", os.path.join(ROOT, "krypy_b200", "csrc"),
                           "-o", out, os.path.join(HERE, "csrc", "cg_emul_host.cpp")])

    def run(*args):
        p = subprocess.run([out] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0 and p.stdout.startswith("ok"), (args, p.stdout, p.stderr)
        return p.stdout
    return run


@pytest.mark.parametrize("nx,ny,its,grid", [(13, 101, 8, 2), (7, 37, 12, 3), (3, 5, 4, 1)])
def test_cg_iteration_chain_emulated(emul, nx, ny, its, grid):
    emul("chain", nx, ny, its, grid)


@pytest.mark.parametrize("ranks", [2, 4])
def test_cg_scalars_over_emulated_ranks(emul, ranks):
    emul("scalars", ranks)
