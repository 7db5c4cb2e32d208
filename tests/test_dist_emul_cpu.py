"""CPU tier: a row-partitioned Arnoldi process over EMULATED ranks with the DEVICE code of the one-wait step
(krypy_b200/csrc/kry_dist_kernels.cuh: dist_dot_kernel with <w, w>, dist_update_scale_kernel with the exact-norm
guard, the halo of v_{k+1} formed from the peers' w, the Givens update in an extra CTA), compiled unchanged for the
host over the CUDA execution emulator (tests/csrc/cuda_emul, tests/csrc/dist_emul_host.cpp).  Every rank is a
group of processes that runs all steps back to back with no synchronisation between ranks but the kernels' own
flag protocol (release / acquire on shared memory), so ranks run ahead of each other as GPUs do -- which
exercises the epoch-parity double buffering of the slot arrays and of w.  Checked: basis and Hessenberg columns
against a long-double reference, columns bitwise identical on all ranks, halo copies bitwise the owners' values,
Givens residuals, epoch counters (one more per guard step).

This is the N > 1 DEVICE path on the CPU tier (the gloo world-2 test covers the host-side partition / halo
planning; tests/test_dist_gpu.py runs 2 and 3 ranks on real hardware)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "dist_emul_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread",
                           "-I", os.path.join(HERE, "csrc", "cuda_emul"), "-I", os.path.join(ROOT, "krypy_b200", "csrc"),
                           "-o", out, os.path.join(HERE, "csrc", "dist_emul_host.cpp")])

    def run(*args, **env):
        p = subprocess.run([out] + [str(a) for a in args], capture_output=True, text=True, timeout=900,
                           env=dict(os.environ, **{k: str(v) for k, v in env.items()}))
        assert p.returncode == 0 and p.stdout.startswith("ok"), (args, p.stdout, p.stderr)
        assert "halo copies bitwise 1, H identical on all ranks 1" in p.stdout
        return p.stdout
    return run


@pytest.mark.parametrize("ranks", [2, 3, 4])
def test_one_wait_arnoldi_steps_over_emulated_ranks(emul, ranks):
    """(ranks, sweep CTAs, steps, rows per rank, Givens tail, guard step)"""
    emul(ranks, 2, 6, 600, 1, -1)


def test_exact_norm_guard_over_emulated_ranks(emul):
    """a step whose w = A v_k lies almost in span(V): <w, w> - sum c^2 keeps ~1e-6 of <w, w>, every CTA of every
    rank takes the guard path (exact local norm, a second exchange inside the kernel: one more epoch)"""
    out = emul(3, 2, 6, 600, 1, 3)
    assert "epochs 7 (want 7)" in out
    emul(2, 3, 5, 1000, 0, 2)        # without the Givens CTA, three sweep CTAs
    emul(2, 1, 3, 300, 1, 0)         # guard in the very first step, one sweep CTA


def test_without_givens_tail_and_odd_sizes(emul):
    emul(2, 2, 4, 701, 0, -1)        # odd local length (scalar tail of the sweeps)
    emul(3, 1, 9, 257, 1, -1)        # nine vectors: full block of 8 + remainder in the update sweep


@pytest.mark.parametrize("slow", [0, 1, 2])
def test_a_rank_held_back_and_schedule_jitter(emul, slow):
    """one rank is held back at every kernel boundary (the others run as far ahead as the flag protocol lets them:
    at most one step, so the two buffers of w and the two slot arrays are enough), with random sleeps in front of
    every release and in a fraction of the acquires: the results do not depend on the interleaving"""
    emul(3, 2, 6, 600, 1, 2, EMUL_SLOW_RANK="%d:30000" % slow, EMUL_JITTER=3000)
    emul(3, 2, 5, 400, 0, -1, EMUL_SLOW_RANK="%d:20000" % slow)
