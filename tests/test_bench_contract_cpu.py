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
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
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
