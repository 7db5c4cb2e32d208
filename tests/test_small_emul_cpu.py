"""CPU tier: the small device recurrences (krypy_b200/csrc/kry_small_kernels.cuh), compiled unchanged for the host
over the CUDA execution emulator (tests/csrc/cuda_emul, tests/csrc/small_emul_host.cpp), in the chains the solvers
run them in:
  * GMRES: kry_givens_update column after column on a random Hessenberg matrix with R left on the "device", then
    kry_tri_solve_t / kry_tri_solve -- the residual norms of min ||beta e_1 - H_k y|| (krypy/linsys.py:982-993) and
    the solution (linsys.py:946) against long-double normal equations;
  * MINRES for config C5 (linsys.py:791-853, diagonal ip_B): SpMV -> kry_lanczos_diag -> kry_minres_recur ->
    kry_minres_update -- the residual norm the recurrence reports is the TRUE ||b - A x_k||_B of the iterate the
    update kernel accumulates, and decreases monotonically;
  * kry_small_qr_apply (R^-1 Q^H c of the row-partitioned projector)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "small_emul_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread",
                           "-I", os.path.join(HERE, "csrc", "cuda_emul"), "-I", os.path.join(ROOT, "krypy_b200", "csrc"),
                           "-o", out, os.path.join(HERE, "csrc", "small_emul_host.cpp")])

    def run(*args):
        p = subprocess.run([out] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0 and p.stdout.startswith("ok"), (args, p.stdout, p.stderr)
        return p.stdout
    return run


@pytest.mark.parametrize("m", [1, 9, 30])
def test_gmres_recurrences_emulated(emul, m):
    emul("gmres", m)


@pytest.mark.parametrize("n,its,grid", [(601, 12, 2), (300, 20, 1), (1000, 6, 3)])
def test_minres_iteration_chain_emulated(emul, n, its, grid):
    emul("minres", n, its, grid)


@pytest.mark.parametrize("d", [1, 20, 32])
def test_small_qr_apply_emulated(emul, d):
    emul("qr", d)
