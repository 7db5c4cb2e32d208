"""CPU tier twin of tests/test_zzz_native_gpu.py: the same checks driven over the numpy test double of the
device layer (tests/fake_device.py).  Validates the host logic around the native complex kernels (which
calls are made with which layout: even rows of the twin storage, interleaved coefficients, chunking, the
KRY_NATIVE_Z switch, the operator's device-format cache) and that the GPU test file itself is sound."""
import numpy as np
import pytest

import fake_device
import test_zzz_native_gpu as zn


@pytest.fixture()
def fake(monkeypatch):
    return fake_device.install(monkeypatch)


@pytest.mark.parametrize("algo", ["cgs", "mgs"])
@pytest.mark.parametrize("nv,j0,n,passes", [(1, 0, 1, 1), (5, 0, 7, 2), (9, 0, 300, 1), (12, 3, 4096, 2), (32, 0, 50, 1)])
def test_orth_fused_z_contract(fake, algo, nv, j0, n, passes):
    zn.check_orth_fused_z(fake, algo, passes, nv, j0, n)
    zn.check_orth_fused_z(fake, algo, passes, nv, j0, n, separate_P=True)


def test_orth_fused_z_edge_contracts(fake):
    zn.test_orth_fused_z_without_tail_and_empty_range(fake)
    zn.test_orth_fused_z_zero_vector(fake)
    zn.test_orth_fused_z_agrees_with_the_twin_kernel(fake)


@pytest.mark.parametrize("cplx", [True, False])
@pytest.mark.parametrize("kind", ["lap2d", "band7", "rand12", "long", "ragged", "tiny"])
def test_spmv_z_contract(fake, kind, cplx):
    zn.check_spmv_z(fake, kind, cplx)


def test_operator_formats(fake):
    zn.test_spmv_z_matches_the_embedded_operator(fake)


@pytest.mark.parametrize("ortho", ["mgs", "cgs2"])
@pytest.mark.parametrize("variant", ["plain", "restarted", "precond"])
def test_native_and_embedding_paths_agree(fake, ortho, variant):
    zn.test_native_and_embedding_paths_agree(fake, ortho, variant)


def test_switch_selects_the_kernels(fake):
    """native: the Arnoldi step of a complex GMRES is kry_spmv_csr_z + kry_orth_fused_z; KRY_NATIVE_Z=0: the
    real kernels on the embedding; complex Lanczos (real coefficients) stays on the real kernel either way"""
    zn.test_native_kernels_are_the_default_and_take_fewer_passes(fake)
    fake.reset_launch_count()
    zn._solve_complex(True, "mgs")
    c = dict(fake.calls)
    assert c.get("orth_fused_z", 0) > 0 and c.get("spmv_z", 0) > 0
    assert c.get("orth_fused", 0) == 0 and c.get("spmv", 0) == 0
    fake.reset_launch_count()
    zn._solve_complex(False, "mgs")
    c = dict(fake.calls)
    assert c.get("orth_fused_z", 0) == 0 and c.get("spmv_z", 0) == 0
    assert c.get("orth_fused", 0) > 0 and c.get("spmv", 0) > 0
    # complex MINRES: Lanczos with real coefficients on the real kernel, the matrix in the native format
    import test_zcomplex_gpu as z
    fake.reset_launch_count()
    z.test_complex_cases_match_reference_fixture_and_oracle("z_minres_herm")
    c = dict(fake.calls)
    assert c.get("spmv_z", 0) > 0 and c.get("orth_fused", 0) > 0 and c.get("orth_fused_z", 0) == 0


def test_complex_cg_and_minres_with_native_matrix(fake):
    zn.test_complex_cg_and_minres_with_native_matrix(fake)


def test_chunked_native_calls(fake):
    fake.reset_launch_count()
    zn.test_arnoldi_more_vectors_than_one_native_call(fake)
    # cgs2 with up to 40 vectors: steps 32..39 take two calls each; dmgs / mgs one call per step
    assert fake.calls["orth_fused_z"] == (40 + 8) + 40 + 40


def test_bench_cplx_tool_dry_run(monkeypatch, capsys):
    """tools/bench_cplx.py over the test double at a tiny grid (Python-level soundness of the measurement tool:
    keys, byte models, the KernelTimer legs); says nothing about performance"""
    import json
    import os
    import runpy
    import sys
    import time

    import torch

    fake_device.install(monkeypatch)

    class Ev(object):
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return 1e3 * (other.t - self.t) + 1e-3

        def synchronize(self):
            pass

    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "bench_cplx.py")
    monkeypatch.setattr(sys, "argv", [tool, "20"])
    runpy.run_path(tool, run_name="__main__")
    d = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][-1])
    assert set(d["runs"]) == {"native_cgs", "embedding_cgs", "native_mgs", "embedding_mgs"}
    for key, r in d["runs"].items():
        assert r["iterations"] > 0 and r["it_per_s"] > 0
        assert r["kernels"]["spmv"]["launches"] > 0 and r["kernels"]["orth"]["GBs"] > 0
    for o in ("cgs", "mgs"):
        assert d["runs"]["native_" + o]["history_max_rel_diff_vs_embedding"] < 1e-7   # (tiny grid: the solve reaches the rounding floor)
        assert d["runs"]["native_" + o]["algorithmic_GBs"] > 0


def _stub_cuda(monkeypatch):
    import contextlib
    import time

    import torch

    class Ev(object):
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return 1e3 * (other.t - self.t) + 1e-3

        def synchronize(self):
            pass

    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)


def test_l2_window_host_logic(fake, monkeypatch):
    """KRY_L2_WINDOW: the window is set once per restarted solve on the Arnoldi vector w = A v_k (the workspace
    buffer, so later cycles find it in place) and removed when the solve ends; off: never set"""
    from krypy_b200 import utils
    monkeypatch.setattr(utils, "_L2_WINDOW_MIN_BYTES", 1)
    for on in (True, False):
        monkeypatch.setattr(utils, "_L2_WINDOW", on)
        fake.l2_calls = []
        zn._solve_complex(True, "cgs", restarted=True)
        sets = [c for c in fake.l2_calls if c is not None]
        if on:
            assert len(sets) >= 1 and len(set(sets)) == 1        # one buffer, the same in every cycle
            assert fake.l2_calls[-1] is None                      # removed at the end
        else:
            assert sets == []


def test_bench_l2window_tool_dry_run(monkeypatch, capsys):
    import json
    import os
    import runpy
    import sys

    fake_device.install(monkeypatch)
    _stub_cuda(monkeypatch)
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "bench_l2window.py")
    monkeypatch.setattr(sys, "argv", [tool, "c2", "c5", "24"])
    runpy.run_path(tool, run_name="__main__")
    d = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][-1])
    assert len(d["runs"]) == 12
    for k, r in d["runs"].items():
        assert r["iterations"] > 0
        if k.endswith("window1"):
            assert r["bitwise_identical_history"] is True


def test_bench_extra_config_c2z_dry_run(monkeypatch):
    """bench_configs.run_device('c2z') (the complex twin of C2 as an extra key of the bench line) over the test
    double at a tiny grid, with reference parity against the unmodified reference / the oracle port"""
    import bench_configs

    fake_device.install(monkeypatch)
    _stub_cuda(monkeypatch)
    out = bench_configs.run_device("c2z", 6538.9, n=16, ref_steps=3, maxiter=10)
    assert "error" not in out and out["iterations"] > 0 and out["it_per_s"] > 0
    assert out["algorithmic_bytes_per_iteration"] > 0 and 0 < out["frac_of_measured_peak"]
    par = out["parity_vs_reference"]
    assert par["entries"] == 4 and par["max_rel_updated"] < 1e-10 and par["rel_last_explicit"] < 1e-10
    assert bench_configs.problem_size("c2z") == 2236 ** 2
