"""ncu target for the native complex kernels: two GMRES(30) cycles (block Gram-Schmidt) of the complex twin of C2
(tools/bench_cplx.py) and nothing else.  usage (tools/gpu_session_r2z.sh):
  ncu --set full --clock-control none --import-source on -k regex:'zorth_kernel|zspmv_staged' -s 100 -c 2 ... python tools/profile_cplx.py [n]"""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import torch

import krypy_b200 as kp
from krypy_b200 import problems

warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
N = n * n
A = (problems.laplace2d(n).astype(np.complex128) - (0.02 + 0.01j) * sp.identity(N, dtype=np.complex128)).tocsr()
rng = np.random.default_rng(0)
b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
ls = kp.linsys.LinearSystem(A, b)
try:
    s = kp.linsys.RestartedGmres(ls, maxiter=30, max_restarts=1, tol=1e-14, ortho="cgs")
except kp.utils.ConvergenceError as e:
    s = e.solver
torch.cuda.synchronize()
print("iterations", len(s.resnorms) - 1, "final", s.resnorms[-1])
