"""Micro-benchmarks of single kernels (CUDA events, warm, inputs >> L2) for tuning.
usage: python tools/bench_kernels.py [spmv] [orth] [cg] ..."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from krypy_b200 import _device, problems
from krypy_b200._lib import KRY_ORTH_CGS, KRY_ORTH_MGS

ctx = _device.Context.get()
what = sys.argv[1:] or ["spmv", "orth"]


def timeit(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


if "spmv" in what:
    for name, A in (("lap2d n=3162", problems.laplace2d(3162)), ("poisson3d n=200", problems.poisson3d(200))):
        N = A.shape[0]
        Ad = ctx.upload_csr(A, torch.float64)
        x = torch.randn(N, dtype=torch.float64, device="cuda")
        y = torch.empty_like(x)
        t = timeit(lambda: ctx.spmv(Ad, x, y))
        by = A.nnz * 12 + 4 * (N + 1) + 16 * N
        print("spmv %-16s stages=%s  %.1f us  %.0f GB/s (algorithmic)" % (name, os.environ.get("KRY_SPMV_STAGES", "def"), t * 1e6, by / t / 1e9))
        dot = ctx.scalars(1)
        t = timeit(lambda: ctx.spmv(Ad, x, y, w=x, dot_out=dot))
        print("spmv+dot %-12s %.1f us  %.0f GB/s" % (name, t * 1e6, (by + 8 * N) / t / 1e9))
        del Ad, x, y

if "orth" in what:
    N = 9998244
    ld = (N + 31) // 32 * 32
    Vs = torch.randn((32, ld), dtype=torch.float64, device="cuda") / np.sqrt(N)
    V = Vs[:, :N]
    q0 = torch.randn(N, dtype=torch.float64, device="cuda")
    q = q0.clone()
    h = ctx.scalars(40)
    for algo, an in ((KRY_ORTH_CGS, "cgs"), (KRY_ORTH_MGS, "mgs")):
        for nv in (1, 8, 16, 24, 31):
            def f():
                ctx.orth_fused(V, V, 0, nv, q, 1, algo, h, nrm=h[nv:], vnext=V[31])
            q.copy_(q0)
            t = timeit(f, reps=10, warm=2)
            by = ((2 * nv + 3 + 2) if algo == KRY_ORTH_CGS else (4 * nv + 2 + 2)) * N * 8
            print("orth %s nv=%2d  %.1f us  %.0f GB/s (algorithmic %.2f GB)" % (an, nv, t * 1e6, by / t / 1e9, by / 1e9))

if "stream" in what:
    N = 1 << 27
    a = torch.randn(N, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
    t = timeit(lambda: b.copy_(a)); print("torch copy %.0f GB/s" % (16 * N / t / 1e9))
    t = timeit(lambda: ctx.axpby(1.0, a, 0.0, None, b)); print("axpby copy %.0f GB/s" % (16 * N / t / 1e9))
    t = timeit(lambda: ctx.axpby(1.0, a, 2.0, b, b)); print("axpby 3-stream %.0f GB/s" % (24 * N / t / 1e9))
    o = ctx.scalars(1)
    t = timeit(lambda: ctx.block_dot(a.reshape(1, -1), 1, b, o)); print("dot %.0f GB/s" % (16 * N / t / 1e9))
