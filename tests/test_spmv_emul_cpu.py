"""CPU tier: the DEVICE code of kry_spmv_csr / kry_spmv_csr_mdot (krypy_b200/csrc/kry_spmv_kernels.cuh: the
TMA-staged, warp-specialised kernel with its <w, y> and multi-vector dot epilogues, the warp-per-row kernel),
compiled unchanged for the host over the CUDA execution emulator (tests/csrc/cuda_emul: the mbarrier / bulk-copy
model checks alignment, byte counts and the phases of the producer / consumer ring) and compared with
extended-precision references (tests/csrc/spmv_emul_host.cpp).

The staged kernel sums every row in storage order with separately rounded products and sums -- scipy's
csr_matvec -- so its rows must be BIT-IDENTICAL to that ordered sum: the bit-exactness the GPU tier asserts
against scipy (tests/test_kernels_gpu.py::test_spmv_stencil_bitexact) is checked here on the same device code."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "spmv_emul_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread",
                           "-I", os.path.join(HERE, "csrc", "cuda_emul"), "-I", os.path.join(ROOT, "krypy_b200", "csrc"),
                           "-o", out, os.path.join(HERE, "csrc", "spmv_emul_host.cpp")])

    def run(*args):
        p = subprocess.run([out] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0 and p.stdout.startswith("ok"), (args, p.stdout, p.stderr)
        return p.stdout
    return run


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("kind", ["stencil5", "band7", "rand12", "ragged", "tiny", "long"])
def test_spmv_kernels_emulated(emul, kind, dtype):
    """5 / 7 / ~12 entries per row (stage capacities 6 / 8 / 16), empty rows, a tile beyond the stage capacity
    (direct loads), nnz % 4 != 0 (the tail the 16-byte bulk copies cannot cover), more tiles than CTAs (the ring
    wraps), long rows (warp-per-row kernel)"""
    out = emul(kind, dtype, "plain", 2)
    if kind != "long":
        rows = int(out.split("rows=")[1].split()[0])
        assert int(out.rsplit(" ", 1)[1]) == rows          # every row bit-identical to the ordered sum
    emul(kind, dtype, "plain", 3)


@pytest.mark.parametrize("kind", ["stencil5", "band7", "ragged", "long"])
def test_spmv_dot_epilogue_emulated(emul, kind):
    """<w, A x> in the epilogue (linsys.py:634, CG's <p, Ap>): deterministic last-CTA reduction, ticket left reset"""
    emul(kind, "f64", "dot", 2)
    emul(kind, "f32", "dot", 3)


@pytest.mark.parametrize("kind,nb", [("stencil5", 1), ("stencil5", 8), ("band7", 5), ("rand12", 8), ("ragged", 3)])
def test_spmv_multi_dot_epilogue_emulated(emul, kind, nb):
    """c_j = <B_j, A x>, j < nb, and <A x, A x> in the epilogue (kry_spmv_csr_mdot)"""
    emul(kind, "f64", "mdot", 2, nb)
