"""CPU tier: a CholQR2 orthonormalisation -- the set-up of the deflation projector (config C4; krypy/utils.py:680-707,
deflation.py:33-56 as utils._cholqr2 runs it) -- with the DEVICE code of its two kernels, kry_gram (cp.async
double-buffered staging, 4x4 register blocks per warp, only the upper block triangle for X^H X, deterministic
last-CTA reduction) and kry_block_trsm, compiled unchanged for the host over the CUDA execution emulator
(tests/csrc/cuda_emul, tests/csrc/block_emul_host.cpp): after two rounds |Q^T Q - I| is at rounding level; and
X^H Y for two different blocks against a long-double reference."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "block_emul_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread",
                           "-I", os.path.join(HERE, "csrc", "cuda_emul"), "-I", os.path.join(ROOT, "krypy_b200", "csrc"),
                           "-o", out, os.path.join(HERE, "csrc", "block_emul_host.cpp")])

    def run(*args):
        p = subprocess.run([out] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0 and p.stdout.startswith("ok"), (args, p.stdout, p.stderr)
        return p.stdout
    return run


@pytest.mark.parametrize("dtype,n,d,grid", [("f64", 3001, 20, 2), ("f64", 777, 5, 3), ("f64", 500, 17, 2), ("f64", 130, 13, 1),
                                            ("f32", 2000, 9, 2)])
def test_cholqr2_with_its_own_kernels_emulated(emul, dtype, n, d, grid):
    """(d <= 20: kry_gram holds ceil(kx/4) * ceil(ky/4) <= 32 output blocks, the host falls back to MGS beyond)"""
    emul("cholqr2", dtype, n, d, grid)


@pytest.mark.parametrize("n,kx,ky,grid", [(1000, 7, 13, 2), (130, 20, 20, 1), (129, 1, 1, 2), (4000, 4, 33, 3)])
def test_gram_of_two_blocks_emulated(emul, n, kx, ky, grid):
    emul("gram", "f64", n, kx, ky, grid)
