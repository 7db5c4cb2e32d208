"""CPU tier: the bench.py JSON contract.  The reference arm runs here (it times the oracle port on the
host cores; small grid so it takes seconds) and the committed B200 lines under profiles/ are checked
for the keys the driver and the judge read."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "48",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "gmres_iterations_per_second" and d["value"] > 0
    assert d["unit"] == "iterations/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["gpu_launches"] == 0 and "workload" in d["config"] and "model" not in d["config"]
    # "reference" = the unmodified reference installed into baseline/_ref; "port" only without that install
    want = "reference" if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "krypy")) else "port"
    assert d["cpu_baseline"]["kind"] == want and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--n", "32", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and "{" not in out.stdout


def test_committed_b200_lines_carry_the_contract():
    for n in (1, 2, 4, 8):
        p = os.path.join(ROOT, "profiles", "r1_bench_n%d.json" % n)
        with open(p) as f:
            d = json.loads([ln for ln in f.read().splitlines() if ln.strip().startswith("{")][-1])
        assert (BASE_KEYS | {"roofline", "clocks"}) <= set(d), (n, (BASE_KEYS | {"roofline", "clocks"}) - set(d))
        assert d["n_gpus"] == n and d["gpu_launches"] > 0 and d["value"] > 0
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
            assert d["cpu_baseline"]["kind"] in ("port", "reference")


def test_b200_arm_assembles_its_line_over_the_test_double(monkeypatch, capsys):
    """Dry run of bench.run_b200 at a tiny grid: the device layer is the numpy test double and the few
    torch.cuda calls of the bench are stubbed.  Catches Python-level breakage of the bench (names, keys,
    the e2e / breakdown / per-solve legs) without a GPU; says nothing about performance."""
    import time
    import types

    import torch

    import bench
    import fake_device

    fake = fake_device.install(monkeypatch)

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
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "device"}))
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, impl="b200", n=24, ortho="cgs",
                                 no_cpu_baseline=False, no_e2e=False, cpu_iters=3, no_mgs=False,
                                 no_extra_configs=True, ref_iters=3)
    bench.run_b200(args, 0, 1, 0)
    out = capsys.readouterr().out
    lines = [ln for ln in out.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "gmres_iterations_per_second" and d["n_gpus"] == 1 and d["iterations"] == 60
    assert d["gpu_launches"] > 0 and d["value"] > 0 and d["dtype"] == "f64"
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert "error" not in d["e2e"]["per_solve"] and 0 < d["e2e"]["per_solve"]["iterations_per_upload"] <= 150
    assert "error" not in d["e2e"]["breakdown_ms"] and d["e2e"]["breakdown_ms"]["total"] > 0
    assert "error" not in d["mgs_value"] and d["mgs_value"]["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] > 0
    assert d["parity_vs_cpu_max_rel"] < 1e-10           # product (over the double) vs oracle, same inputs
    assert fake.launch_count() > 0


def test_smoke_logic_over_the_test_double(monkeypatch, capsys):
    """__graft_entry__.smoke() with the device layer replaced by the test double: its oracle comparison
    and thresholds are sound (the real run on cuda:0 is the driver's)."""
    import torch

    import fake_device
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    fake_device.install(monkeypatch)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    g.smoke()
    assert "smoke ok" in capsys.readouterr().out


def _stub_cuda(monkeypatch):
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
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)


def test_named_configurations_dry_run_with_reference_parity(monkeypatch):
    """bench_configs.run_device (C3, C4, C5) at toy sizes over the test double: the result keys the bench
    line / profiles carry, and the parity leg against the host reference (the unmodified reference when
    baseline/_ref exists, else the oracle port)."""
    import bench_configs
    import fake_device

    fake_device.install(monkeypatch)
    _stub_cuda(monkeypatch)
    for name, n, steps, tol in (("c3", 10, 6, 1e-10), ("c5", 24, 6, 1e-4), ("c4", 24, 0, None)):
        out = bench_configs.run_device(name, 6538.9, n=n, maxiter=30 if name == "c4" else 12, ref_steps=steps)
        assert out["iterations"] > 0 and out["it_per_s"] > 0 and out["algorithmic_GBs"] > 0
        assert 0 < out["frac_of_measured_peak"] and out["n_gpus"] == 1
        if steps:
            par = out["parity_vs_reference"]
            assert par["entries"] == steps + 1 and par["reference_kind"] in ("reference", "port")
            assert par["max_rel_updated"] <= tol, (name, par)
        else:
            assert out["projector_setup_s"] >= 0 and out["undeflated_it_per_s"] > 0
    # the byte model reproduces SURVEY 8(d): 192 N (C3, fp64), 144 N (C5, fp32) up to boundary terms
    N = 400 ** 3
    assert abs(bench_configs.bytes_per_iteration("c3", N, 7 * N, 8) / N - 192.0) < 1e-6
    N = 4000 ** 2
    assert abs(bench_configs.bytes_per_iteration("c5", N, 5 * N, 4) / N - 144.0) < 1e-5


def test_b200_arm_reports_the_l2_window(monkeypatch, capsys):
    """the same dry run with the L2 residency window in effect (threshold lowered to the tiny grid): the line says
    so and explains why the Gram-Schmidt kernel's algorithmic GB/s may exceed the DRAM copy peak"""
    import time
    import types

    import torch

    import bench
    import fake_device
    from krypy_b200 import utils

    fake_device.install(monkeypatch)
    monkeypatch.setattr(utils, "_L2_WINDOW", True)
    monkeypatch.setattr(utils, "_L2_WINDOW_MIN_BYTES", 1)

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
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "device"}))
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, impl="b200", n=24, ortho="cgs",
                                 no_cpu_baseline=True, no_e2e=True, cpu_iters=3, no_mgs=True,
                                 no_extra_configs=True, ref_iters=3)
    bench.run_b200(args, 0, 1, 0)
    d = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.strip().startswith("{")][-1])
    assert d["l2_window"]["enabled"] is True and d["l2_window"]["hit_ratio"] == 1.0
    assert "l2_resident_vector" in d["roofline"]
