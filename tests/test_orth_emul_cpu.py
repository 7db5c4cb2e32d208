"""CPU tier: the DEVICE code of the headline kernels -- orth_kernel (kry_orth_fused / kry_orth_fused_dist: block
classical and exact modified Gram-Schmidt, Lanczos pre-subtraction, norm, normalised store) and proj_kernel
(kry_project) of krypy_b200/csrc/kry_orth_kernels.cuh -- compiled unchanged for the host with g++ over the CUDA
execution emulator (tests/csrc/cuda_emul) and compared with extended-precision references
(tests/csrc/orth_emul_host.cpp).  Row-partitioned runs are emulated as 2 or 3 ranks of CTAs that exchange their
partial sums through shared slot / flag arrays exactly as the GPUs do over NVLink peer memory: the N > 1 device
path on the CPU tier, beside the gloo world-2 test of the host-side partition planning.

The kernels are the ones the GPU tier runs (device code byte-identical to the validated build, tools/sass_identity.py)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "orth_emul_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread",
                           "-I", os.path.join(HERE, "csrc", "cuda_emul"), "-I", os.path.join(ROOT, "krypy_b200", "csrc"),
                           "-o", out, os.path.join(HERE, "csrc", "orth_emul_host.cpp")])

    def run(*args):
        p = subprocess.run([out] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0 and p.stdout.startswith("ok"), (args, p.stdout, p.stderr)
        return p.stdout
    return run


# nv, j0, n: every dots tile size 1..16 and update remainder 0..7 occurs, two- and three-tile passes, the
# stride-unrolled loops (n > 4 * grid * 256 packs), odd lengths (scalar tail), an empty range
SHAPES = [(0, 0, 100), (1, 0, 1), (1, 0, 4099), (2, 0, 3001), (3, 0, 3001), (4, 0, 5001), (5, 0, 300), (6, 5, 999),
          (7, 0, 2500), (8, 0, 2049), (9, 0, 2500), (10, 0, 700), (11, 0, 700), (12, 0, 700), (13, 0, 700), (14, 0, 700),
          (15, 0, 1000), (16, 0, 1000), (17, 0, 600), (31, 2, 800), (33, 0, 500), (40, 3, 450)]


# exact MGS has no tiles (one sweep and one grid-wide reduction per vector): a few counts cover it
SHAPES_MGS = [(0, 0, 100), (1, 0, 1), (1, 0, 4099), (2, 0, 3001), (3, 0, 3001), (5, 0, 300), (6, 5, 999), (9, 0, 2500),
              (17, 0, 600)]


@pytest.mark.parametrize("algo", [0, 1], ids=["cgs", "mgs"])
@pytest.mark.parametrize("passes", [1, 2])
def test_orth_kernel_emulated_f64(emul, algo, passes):
    for nv, j0, n in (SHAPES if algo == 0 else SHAPES_MGS):
        emul("orth", "f64", 2, algo, passes, nv, j0, n, 2, 1, 0, 0)


@pytest.mark.parametrize("algo", [0, 1], ids=["cgs", "mgs"])
def test_orth_kernel_emulated_variants(emul, algo):
    """unaligned path (VEC = 1), fp32 storage (VEC = 4), Lanczos-style pre-subtraction, separate update basis,
    other grid sizes"""
    emul("orth", "f64", 1, algo, 1, 9, 0, 1001, 2, 1, 0, 1)
    emul("orth", "f64", 1, algo, 2, 3, 1, 333, 3, 1, 1, 0)
    emul("orth", "f32", 4, algo, 1, 9, 0, 4003, 2, 1, 0, 0)
    emul("orth", "f32", 4, algo, 2, 17 if algo == 0 else 5, 0, 2002, 2, 1, 1, 1)
    emul("orth", "f32", 1, algo, 1, 5, 0, 777, 2, 1, 0, 0)
    emul("orth", "f64", 2, algo, 1, 6, 5, 4100, 1, 1, 1, 0)      # Lanczos: one vector, pre-subtraction
    emul("orth", "f64", 2, algo, 1, 12, 0, 4100, 5, 1, 0, 1)


@pytest.mark.parametrize("algo", [0, 1], ids=["cgs", "mgs"])
@pytest.mark.parametrize("ranks", [2, 3])
def test_orth_kernel_emulated_row_partitioned(emul, algo, ranks):
    """kry_orth_fused_dist: the reductions are completed across the emulated ranks inside the kernel (publish to
    every peer's slots, release the flag, acquire all flags, rank-order sum): the coefficients are bitwise
    identical on all ranks and the epoch counters advance in step"""
    out = emul("orth", "f64", 2, algo, 1, 7, 0, 2000, 2, ranks, 0, 0)
    assert "ranks identical 1" in out
    emul("orth", "f64", 2, algo, 2, 4, 1, 1500, 2, ranks, 1, 0)
    emul("orth", "f64", 2, algo, 1, 20 if algo == 0 else 6, 0, 900, 1, ranks, 0, 1)
    emul("orth", "f32", 4, algo, 1, 5, 0, 1200, 2, ranks, 0, 0)


@pytest.mark.parametrize("with_qr", [1, 0])
def test_proj_kernel_emulated(emul, with_qr):
    """a <- (I - V R^-1 Q^H W^H)^iterations a (krypy/utils.py:604-627): one to three applications, d = 1 .. 33"""
    for d, its, n, grid in [(1, 1, 50, 1), (5, 2, 3001, 2), (20, 2, 3001, 2), (33, 3, 700, 3), (16, 1, 4100, 2), (17, 2, 999, 2)]:
        emul("proj", "f64", 2, d, its, n, grid, with_qr)
    emul("proj", "f64", 1, 7, 2, 1001, 2, with_qr)
    emul("proj", "f32", 4, 5, 1, 2000, 3, with_qr)
