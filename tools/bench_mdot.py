"""Single-GPU timing of the SpMV with the multi-dot epilogue (kry_spmv_csr_mdot) against the plain SpMV + a
separate block dot, at the per-rank sizes of config C2 on 1 / 4 / 8 GPUs.  ANALYSIS TOOL.
    python tools/bench_mdot.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us


def main():
    from krypy_b200 import _device, problems
    ctx = _device.Context.get()
    td = torch.float64
    for n in (3162, 1581, 1118):
        A = problems.laplace2d(n)
        N = n * n
        Ad = ctx.upload_csr(A, td)
        x = torch.randn(N, dtype=td, device="cuda")
        y = torch.empty(N, dtype=td, device="cuda")
        V = torch.randn(32, N, dtype=td, device="cuda")
        out = ctx.scalars(64)
        t_sp = timeit(lambda: ctx.spmv(Ad, x, y))
        print("N=%d  spmv alone %.1f us (%.0f GB/s)" % (N, t_sp, 80.0 * N / t_sp / 1e3))
        for nb in (1, 4, 8, 12, 16, 24, 30):
            t_md = timeit(lambda: ctx.spmv_mdot(Ad, x, y, V, nb, 1, out))
            t_dot = timeit(lambda: ctx.block_dot(V, nb, y, out, 0, None))
            by = (80.0 + 8.0 * nb) * N
            print("   nb=%2d  mdot %.1f us (%.0f GB/s)   spmv + block_dot %.1f us   (dot alone %.1f us)"
                  % (nb, t_md, by / t_md / 1e3, t_sp + t_dot, t_dot))


if __name__ == "__main__":
    main()
