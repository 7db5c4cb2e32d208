"""CPU tier: a ROW-PARTITIONED Jacobi-preconditioned CG (BASELINE config C3 at 8 GPUs; krypy/linsys.py:593-689 on
krypy_b200/dist.py's row blocks) over EMULATED ranks with the DEVICE code of every kernel of its iteration,
compiled unchanged for the host over the CUDA execution emulator (tests/csrc/cuda_emul,
tests/csrc/cgdist_emul_host.cpp): kry_xpby_dev -> kry_dist_halo (flag handshake + P2P gather of the remote entries
of p_k) -> kry_spmv_csr with the <p, Ap> epilogue on [local | halo] -> kry_peer_allreduce -> kry_cg_update_dev ->
kry_cg_scalars (global rho through the peer slots).  Three peer operations per iteration share one epoch counter;
the ranks run all iterations back to back with no synchronisation but the kernels' flag protocol.  Checked against
a long-double CG on the global system: rho / alpha of every iteration (bitwise identical on all ranks), the
iterate, the epoch counters -- also with one rank held back and random schedule jitter."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "cgdist_emul_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread",
                           "-I", os.path.join(HERE, "csrc", "cuda_emul"), "-I", os.path.join(ROOT, "krypy_b200", "csrc"),
                           "-o", out, os.path.join(HERE, "csrc", "cgdist_emul_host.cpp")])

    def run(*args, **env):
        p = subprocess.run([out] + [str(a) for a in args], capture_output=True, text=True, timeout=900,
                           env=dict(os.environ, **{k: str(v) for k, v in env.items()}))
        assert p.returncode == 0 and p.stdout.startswith("ok"), (args, p.stdout, p.stderr)
        assert "scalars bitwise identical on all ranks 1" in p.stdout
        return p.stdout
    return run


@pytest.mark.parametrize("ranks,nxl,ny,its,ctas", [(2, 5, 61, 6, 2), (3, 4, 37, 8, 2), (4, 3, 29, 5, 1)])
def test_row_partitioned_cg_over_emulated_ranks(emul, ranks, nxl, ny, its, ctas):
    emul(ranks, nxl, ny, its, ctas)


@pytest.mark.parametrize("slow", [0, 1, 2])
def test_row_partitioned_cg_with_a_rank_held_back(emul, slow):
    """the two alternating p buffers and the parity-double-buffered slots are enough however far a rank falls behind"""
    emul(3, 4, 37, 6, 2, EMUL_SLOW_RANK="%d:20000" % slow, EMUL_JITTER=3000)
