"""Where does the GPU idle between the restart cycles of a row-partitioned GMRES(30)?  Runs config C2 at the
PER-RANK size of an 8-GPU run on one GPU (world = 1: the peer protocol talks to itself, same host path, similar
kernel durations) under torch.profiler and lists every gap > 4 us between consecutive GPU activities of two
cycles with the activities on both sides.  ANALYSIS TOOL (times under a profiler are not bench values).
    python tools/cycle_boundary_profile.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29683")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    import warnings
    warnings.simplefilter("ignore")
    import krypy_b200 as kp
    from krypy_b200 import dist as kd, problems
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1118
    m = 30
    N = n * n
    kd.init()
    part = kd.RowPartition(N, 1, 0)
    ls = kd.DistLinearSystem(problems.laplace2d(n), problems.rhs_normal(N), part)
    ws = kp.utils.SolverWorkspace()
    state = {"x": None, "r": None}

    def cycle():
        try:
            s = kp.linsys.Gmres(ls, x0=state["x"], maxiter=m, tol=1e-13, ortho="cgs", _workspace=ws, _prelaunch=True,
                                _x0_residual=state["r"])
        except kp.utils.ConvergenceError as e:
            s = e.solver
        state["x"], state["r"] = s.__dict__["_xk_dev"].reshape(-1), s.__dict__.get("_last_residual")

    for _ in range(6):
        cycle()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        cycle()
    e1.record()
    torch.cuda.synchronize()
    print("N = %d (world 1): %.3f ms per cycle unprofiled (%.1f us per step)" % (N, e0.elapsed_time(e1) / 10,
                                                                                  e0.elapsed_time(e1) * 1e3 / 300))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            cycle()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if str(e.device_type).endswith("CUDA")]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    busy = sum(e.time_range.end - e.time_range.start for e in ev)
    print("profiled: %d GPU activities, busy %.3f ms of %.3f ms" % (len(ev), busy / 1e3, (ev[-1].time_range.end - t0) / 1e3))
    # the boundary between the 2nd and 3rd profiled cycle: everything that is not one of the cycle's step kernels
    step_names = ("spmv_staged", "dist_dot", "dist_update_scale", "dist_halo")
    prev = None
    gaps = []
    for e in ev:
        if prev is not None:
            gap = e.time_range.start - prev.time_range.end
            if gap > 4.0:
                gaps.append((gap, prev, e))
        prev = e
    tot = sum(g for g, _, _ in gaps)
    print("gaps > 4 us: %d, total %.3f ms" % (len(gaps), tot / 1e3))
    for gap, a, b in gaps:
        print("  %8.1f us idle at t=%9.1f us   after %-44s (%.1f us)   before %s" % (
            gap, a.time_range.end - t0, a.name[:44], a.time_range.end - a.time_range.start, b.name[:60]))
    # host-side view of the same boundary: CPU ops by total self time
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=60))
    kd.shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
