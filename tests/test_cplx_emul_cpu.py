"""CPU tier: the DEVICE code of the native complex kernels (krypy_b200/csrc/kry_zorth.cuh: the complex sweeps and
zorth_kernel of kry_orth_fused_z; kry_zspmv.cuh: the TMA-staged and the warp-per-row kernel of kry_spmv_csr_z),
compiled unchanged for the host with g++ over a small CUDA execution emulator (tests/csrc/cuda_emul: one OS
thread per CUDA thread, one process per CTA, pthread barriers for __syncthreads / shuffles / grid.sync, an
mbarrier + bulk-copy model that checks alignment, byte counts and the ring protocol) and compared with
extended-precision references inside tests/csrc/cplx_emul_host.cpp.

What this proves: indexing, tiling / remainder specialisation, reduction plumbing and the synchronisation
protocol of the kernels on multi-CTA grids.  The GPU tier (tests/test_zzz_native_gpu.py) runs the same kernels
through the C ABI on a B200."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "cplx_emul_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread",
                           "-I", os.path.join(HERE, "csrc", "cuda_emul"), "-I", os.path.join(ROOT, "krypy_b200", "csrc"),
                           "-o", out, os.path.join(HERE, "csrc", "cplx_emul_host.cpp")])

    def run(*args):
        p = subprocess.run([out] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0 and p.stdout.startswith("ok"), (args, p.stdout, p.stderr)
        return p.stdout
    return run


ORTH = [  # nv, j0, n: every dots tile size 1..8, every update remainder 0..7, one / two / four tiles, unrolled loops
    (0, 0, 100), (1, 0, 1), (1, 0, 7), (2, 0, 3001), (3, 0, 3001), (4, 0, 5000), (5, 0, 300), (6, 5, 999),
    (7, 0, 2500), (8, 0, 2049), (9, 0, 2500), (15, 0, 1000), (16, 0, 1000), (17, 0, 600), (24, 2, 800),
    (31, 0, 500), (32, 0, 700), (3, 3, 450),
]


@pytest.mark.parametrize("algo", [0, 1], ids=["cgs", "mgs"])
@pytest.mark.parametrize("passes", [1, 2])
def test_zorth_kernel_emulated(emul, algo, passes):
    # (exact MGS has no tiles -- one sweep and one grid-wide reduction per vector: a few counts cover it)
    for nv, j0, n in (ORTH if algo == 0 else [s for s in ORTH if s[0] <= 9 or s[0] == 17]):
        emul("orth", algo, passes, nv, j0, n, 2, 0)


@pytest.mark.parametrize("algo", [0, 1], ids=["cgs", "mgs"])
@pytest.mark.parametrize("grid", [1, 3, 5])
def test_zorth_kernel_emulated_grids_and_update_basis(emul, algo, grid):
    """other grid sizes (partials indexing, fixed-order final sums) and a separate update basis P"""
    emul("orth", algo, 1, 9, 0, 4100, grid, 1)
    emul("orth", algo, 2, 4, 1, 1300, grid, 1)


@pytest.mark.parametrize("cplx", [1, 0], ids=["complex_vals", "real_vals"])
@pytest.mark.parametrize("kind", ["stencil5", "band7", "rand12", "ragged", "tiny", "long"])
def test_zspmv_kernels_emulated(emul, kind, cplx):
    out = emul("spmv", kind, cplx, 2)
    if kind != "long":
        # staged path: every row bit-identical to the sum in storage order with separately rounded products
        rows = int(out.split("rows=")[1].split()[0])
        assert int(out.rsplit(" ", 1)[1]) == rows
    emul("spmv", kind, cplx, 3)


def test_the_product_build_never_defines_the_emulation_switch():
    """KRY_EMUL only swaps PTX wrappers / the dynamic shared-memory declaration for host versions in the test
    harnesses; the library is built without it (and has no CPU path)"""
    for rel in ("krypy_b200/csrc/Makefile", "__graft_entry__.py", "krypy_b200/_lib.py", "krypy_b200/_device.py"):
        assert "KRY_EMUL" not in open(os.path.join(ROOT, rel)).read(), rel
    import glob
    for path in glob.glob(os.path.join(ROOT, "krypy_b200", "csrc", "*.cu")):
        assert "define KRY_EMUL" not in open(path).read(), path
