"""CPU tier: the kernel launch plan of the real-valued path is pinned.  For every fixture case the
host layer, driven over the test double, must issue exactly the recorded number of launches per
C-ABI entry point (tests/golden/launch_plan.json).  A refactor of the host layer that changes what
is launched on the validated path shows up here before any GPU time is spent; a deliberate change
regenerates the file:  python tests/test_launch_plan_cpu.py --regenerate"""
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
PLAN = os.path.join(HERE, "golden", "launch_plan.json")
VARIANTS = [("lap2d_gmres30", dict(ortho=o)) for o in ("cgs", "cgs2", "dmgs")]


def _collect():
    import cases
    import fake_device
    import runners
    from krypy_b200 import _device
    fake = fake_device.FakeContext()
    saved = _device.Context.get
    _device.Context.get = classmethod(lambda cls, device=None: fake)
    out = {}
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for name, kw in [(n, {}) for n in cases.ALL_CASES] + VARIANTS:
                fake.calls = {}
                runners.run_product(name, **kw)
                key = name + ("" if not kw else "/" + ",".join("%s=%s" % it for it in sorted(kw.items())))
                out[key] = dict(sorted(fake.calls.items()))
    finally:
        _device.Context.get = saved
    return out


def test_launch_plan_of_the_real_path_is_unchanged():
    with open(PLAN) as f:
        want = json.load(f)
    got = _collect()
    assert set(got) == set(want)
    for name in sorted(want):
        assert got[name] == want[name], (name, {k: (want[name].get(k), got[name].get(k))
                                                for k in set(want[name]) | set(got[name])
                                                if want[name].get(k) != got[name].get(k)})


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    if "--regenerate" in sys.argv:
        with open(PLAN, "w") as f:
            json.dump(_collect(), f, indent=1, sort_keys=True)
        print("wrote", PLAN)
