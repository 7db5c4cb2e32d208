"""CPU tier, build container only: the REFERENCE's own test-suite (/root/reference/test, unmodified,
real and complex cases) run against krypy_b200's host layer with the device layer replaced by the
numpy test double -- tools/reference_suite_on_host_layer.py.  Skipped where /root/reference does not
exist (the GPU box); nothing under -m gpu, smoke() or bench.py reads the reference."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/test"), reason="reference checkout not present")
def test_reference_suite_passes_against_the_host_layer(tmp_path):
    # quick mode: every 5th case of the two huge parametrisations (test_linsys 13,385, test_deflation
    # 8,160 cases); the full run (25,482 passed, ~6 min serially) is `python tools/reference_suite_on_host_layer.py`
    env = dict(os.environ, KRY_REFTEST_SCRATCH=str(tmp_path / "reftest"), KRY_REFTEST_STRIDE="5")
    env.pop("KRY_TEST_DOUBLE", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "reference_suite_on_host_layer.py"),
                          "--concurrent"], capture_output=True, text=True, timeout=2400, env=env, cwd=ROOT)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    assert " failed" not in out.stdout and " error" not in out.stdout, tail
    passed = sum(int(m) for m in re.findall(r"(\d+) passed", out.stdout))
    want = 28 + 3909 + (13385 + 4) // 5 + (8160 + 4) // 5   # small files complete, big ones strided
    assert passed == want, (passed, want, tail)
