"""CPU tier: the DEVICE code of the N-sized streaming kernels (krypy_b200/csrc/kry_vec_kernels.cuh: kry_axpby,
kry_axpy_dev, kry_scale_dev, kry_diag_mul, kry_rot90, kry_block_dot with its sqrt / accumulate epilogues,
kry_block_axpy, kry_block_combine, kry_gemv_dense), compiled unchanged for the host over the CUDA execution emulator
(tests/csrc/cuda_emul) against long-double references (tests/csrc/vec_emul_host.cpp): fp64 / fp32 storage, aligned
/ unaligned paths, scalar tails, tile remainders of the block kernels, 1-3 CTAs."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "vec_emul_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread",
                           "-I", os.path.join(HERE, "csrc", "cuda_emul"), "-I", os.path.join(ROOT, "krypy_b200", "csrc"),
                           "-o", out, os.path.join(HERE, "csrc", "vec_emul_host.cpp")])

    def run(*args):
        p = subprocess.run([out] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0 and p.stdout.startswith("ok"), (args, p.stdout, p.stderr)
        return p.stdout
    return run


@pytest.mark.parametrize("dtype,vec,n,nv,grid", [("f64", 2, 3001, 5, 2), ("f64", 1, 777, 9, 3), ("f32", 4, 4003, 17, 2),
                                                 ("f32", 1, 1001, 4, 2), ("f64", 2, 7, 1, 1), ("f64", 2, 2048, 8, 2),
                                                 ("f64", 2, 1500, 33, 2)])
def test_streaming_kernels_emulated(emul, dtype, vec, n, nv, grid):
    emul(dtype, vec, n, nv, grid)
