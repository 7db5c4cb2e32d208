"""Build-container check (needs /root/reference; NOT used by the GPU tests, smoke() or bench.py):
run the REFERENCE's own test files against krypy_b200's host layer, with the package aliased as
``krypy`` and the device layer replaced by the numpy test double (tests/fake_device.py).

The whole suite runs, real and complex cases (complex: real embedding + twin storage).  Result
recorded in DESIGN.md section 8:
  test_convenience_wrappers.py + test_recycling.py 28 passed | test_linsys.py 13385 passed |
  test_deflation.py 8160 passed | test_utils.py 3909 passed   (= the reference's 25,482 tests)
usage: python tools/reference_suite_on_host_layer.py
"""
import os
import shutil
import subprocess
import sys

REF = "/root/reference/test"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRATCH = os.environ.get("KRY_REFTEST_SCRATCH", "/tmp/krypy_b200_reftest")

CONFTEST = '''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy
numpy.complex = complex; numpy.float = float; numpy.int = int; numpy.Inf = numpy.Infinity = numpy.inf
import krypy_b200
from krypy_b200 import _device
import fake_device
_fake = fake_device.FakeContext()
_device.Context.get = classmethod(lambda cls, device=None: _fake)
for name, mod in (("krypy", krypy_b200), ("krypy.utils", krypy_b200.utils), ("krypy.linsys", krypy_b200.linsys),
                  ("krypy.deflation", krypy_b200.deflation), ("krypy.recycling", krypy_b200.recycling)):
    sys.modules[name] = mod
import pytest

def pytest_collection_modifyitems(config, items):
    # KRY_REFTEST_STRIDE=k: keep every k-th case of the two huge parametrisations (quick mode of the CPU tier)
    import os
    k = int(os.environ.get("KRY_REFTEST_STRIDE", "1"))
    if k > 1:
        big = [it for it in items if it.fspath.basename in ("test_linsys.py", "test_deflation.py")]
        drop = set(id(it) for i, it in enumerate(big) if i %% k)
        items[:] = [it for it in items if id(it) not in drop]


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_call(item):
    outcome = yield
    exc = outcome.excinfo
    if exc is not None and issubclass(exc[0], NotImplementedError):
        outcome.force_exception(pytest.skip.Exception("device path: " + str(exc[1])[:80]))
''' % (ROOT, os.path.join(ROOT, "tests"))

def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference tests not present at " + REF)
    shutil.rmtree(SCRATCH, ignore_errors=True)
    os.makedirs(SCRATCH)
    for f in os.listdir(REF):
        if f.endswith(".py"):
            shutil.copy(os.path.join(REF, f), SCRATCH)
    open(os.path.join(SCRATCH, "conftest.py"), "w").write(CONFTEST)
    runs = [["test_convenience_wrappers.py", "test_recycling.py"], ["test_linsys.py"], ["test_deflation.py"],
            ["test_utils.py"]]
    # usage: ... [-n WORKERS] [--concurrent]
    #   -n WORKERS    pytest-xdist inside each run
    #   --concurrent  start the four pytest runs at once (the CPU tier's quick mode)
    extra = []
    if "-n" in sys.argv:
        extra = ["-n", sys.argv[sys.argv.index("-n") + 1]]
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-W", "ignore"] + extra
    rc = 0
    if "--concurrent" in sys.argv:
        procs = [(r, subprocess.Popen(cmd + r, cwd=SCRATCH, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True)) for r in runs]
        for r, pr in procs:
            out, _ = pr.communicate()
            print("==", " ".join(r))
            print(out[-1500:], flush=True)
            rc |= pr.returncode
    else:
        for r in runs:
            print("==", " ".join(r), flush=True)
            rc |= subprocess.call(cmd + r, cwd=SCRATCH)
    sys.exit(rc)


if __name__ == "__main__":
    main()
