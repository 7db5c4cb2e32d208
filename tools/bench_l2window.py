"""A/B measurement of the L2 residency window on w = A v_k (kry_l2_window, KRY_L2_WINDOW; krypy_b200/utils.py): config C2
(GMRES(30), 2-D 5-point Laplacian, N = 9,998,244, fp64) with block (cgs) and exact modified (mgs, the drop-in
default) Gram-Schmidt, and config C5 (MINRES, diagonal ip_B, fp32, N = 16M), each with the window off and on, on
torch's default stream and on a side stream, in ONE process on the same resident inputs.  Results must be
bitwise identical (the window is a cache hint); the figure of interest is iterations/s.

usage: python tools/bench_l2window.py [c2 c5] > gpurun_out/l2window.json"""
import json
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import krypy_b200 as kp
from krypy_b200 import _device, problems, utils

warnings.simplefilter("ignore")
which = [a for a in sys.argv[1:] if not a.isdigit()] or ["c2", "c5"]
small = [int(a) for a in sys.argv[1:] if a.isdigit()]          # a grid size for dry runs
ctx = _device.Context.get()
out = {"runs": {}, "what": __doc__.split("\n\n")[0]}


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    try:
        s = fn()
    except kp.utils.ConvergenceError as e:
        s = e.solver
    e1.record()
    torch.cuda.synchronize()
    return s, e0.elapsed_time(e1) * 1e-3


def ab(name, make, warm):
    res = {}
    for stream_mode in ("default", "side"):
        side = torch.cuda.Stream() if stream_mode == "side" else None
        for window in (False, True):
            utils._L2_WINDOW = window
            ctx.l2_window_error = None
            with torch.cuda.stream(side) if side is not None else torch.cuda.stream(torch.cuda.current_stream()):
                timed(warm)
                best = None
                for _ in range(3):
                    s, sec = timed(make)
                    its = len(s.resnorms) - 1
                    if best is None or sec < best[1]:
                        best = (its, sec, s)
                info = getattr(ctx, "_l2win", None)
                timer = _device.KernelTimer()
                ctx.timer = timer
                try:
                    timed(warm)
                    summ = timer.summary()
                finally:
                    ctx.timer = None
            its, sec, s = best
            key = "%s_%s_window%d" % (name, stream_mode, int(window))
            res[key] = {"iterations": its, "seconds": sec, "it_per_s": its / sec, "final_resnorm": float(s.resnorms[-1]),
                        "resnorm_checksum": float(np.sum(np.array(s.resnorms) * np.arange(1, len(s.resnorms) + 1))),
                        "kernels_avg_us": {t: float(np.mean(d["ms"]) * 1e3) for t, d in summ.items()},
                        "window_error": getattr(ctx, "l2_window_error", None)}
            print("%-28s %8.1f it/s  %s %s" % (key, its / sec, res[key]["kernels_avg_us"], res[key]["window_error"] or ""),
                  file=sys.stderr)
            del s, best
        ctx.l2_window(None)
        a, b = res["%s_%s_window0" % (name, stream_mode)], res["%s_%s_window1" % (name, stream_mode)]
        b["speedup_over_window_off"] = b["it_per_s"] / a["it_per_s"] if a["it_per_s"] > 0 else None
        b["bitwise_identical_history"] = bool(a["resnorm_checksum"] == b["resnorm_checksum"]
                                              and a["final_resnorm"] == b["final_resnorm"])
    utils._L2_WINDOW = False
    return res


if "c2" in which:
    n = small[0] if small else 3162
    A = problems.laplace2d(n)
    b = problems.rhs_normal(n * n)
    ls = kp.linsys.LinearSystem(A, b)
    # the device limits, once
    probe = torch.empty(n * n, dtype=torch.float64, device=ctx.device)
    out["device"] = dict(zip(("max_set_aside", "max_window", "set_aside", "window", "hit_ratio"), ctx.l2_window(probe) or ()))
    out["device"]["error"] = getattr(ctx, "l2_window_error", None)
    ctx.l2_window(None)
    del probe
    for ortho in ("cgs", "mgs"):
        out["runs"].update(ab("c2_" + ortho,
                              lambda: kp.linsys.RestartedGmres(ls, maxiter=30, max_restarts=4, tol=1e-14, ortho=ortho),
                              lambda: kp.linsys.RestartedGmres(ls, maxiter=30, max_restarts=0, tol=1e-14, ortho=ortho)))
    del ls, A, b
    torch.cuda.empty_cache()

if "c5" in which:
    import bench_configs
    P = bench_configs.problem("c5", n=small[0] if small else None)
    ls = kp.linsys.LinearSystem(P["A"], P["b"], dtype=np.float32, **P["ls"])
    out["runs"].update(ab("c5_minres",
                          lambda: kp.linsys.Minres(ls, maxiter=50, tol=1e-5),
                          lambda: kp.linsys.Minres(ls, maxiter=5, tol=1e-5)))
print(json.dumps(out))
