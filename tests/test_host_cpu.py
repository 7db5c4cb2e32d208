"""CPU-only tests: C-ABI surface (library loads, exports every declared symbol), host logic
that needs no kernel launch, and the world_size-2 gloo run of the row-partition planning."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from krypy_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "krypy_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(kry_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    # and the ctypes prototypes cover exactly the header
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert lib.kry_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import krypy_b200 as kp
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kp.gmres(np.eye(4), np.ones(4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kp.linsys.LinearSystem(np.eye(4), np.ones(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "krypy_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no CPU fallback", ""), fn


def test_problem_generators_match_kron_definitions():
    import scipy.sparse as sp
    from krypy_b200 import problems
    n = 7
    I = sp.identity(n)
    T = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(n, n))
    L = (sp.kron(I, T) + sp.kron(T, I)).tocsr()
    assert abs(problems.laplace2d(n) - L).max() == 0
    c = 0.1
    C = c * sp.diags([-1, 0, 1], [-1, 0, 1], shape=(n, n))
    CD = (sp.kron(I, T + C) + sp.kron(T + 0.5 * C, I)).tocsr()
    assert abs(problems.convdiff2d(n, c) - CD).max() < 1e-15
    L3 = (sp.kron(sp.kron(I, I), T) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(T, I), I)).tocsr()
    assert abs(problems.poisson3d(n) - L3).max() == 0
    A, B = problems.shifted_laplace_B(n, sigma=0.3, dtype=np.float64)
    K = (L - 0.3 * sp.identity(n * n)).toarray()
    assert np.allclose(B.toarray() @ A.toarray(), K)                 # A = B^-1 (L - sigma I)
    for M in (problems.laplace2d(n), problems.poisson3d(4), problems.convdiff2d(n)):
        assert M.indices.dtype == np.int32 and M.has_sorted_indices
    # row blocks with global column indices
    assert abs(problems.laplace2d(n, rows=(10, 30)) - L[10:30]).max() == 0


def test_host_operator_dispatch_and_dtype_rules():
    import scipy.sparse as sp
    from krypy_b200 import utils as u
    A = np.arange(9.0).reshape(3, 3)
    assert isinstance(u.get_linearoperator((3, 3), A), u.MatrixLinearOperator)
    assert isinstance(u.get_linearoperator((3, 3), None), u.IdentityLinearOperator)
    assert isinstance(u.get_linearoperator((3, 3), sp.diags([1.0, 2.0, 3.0]).tocsr()), u.DiagonalLinearOperator)
    assert isinstance(u.get_linearoperator((3, 3), sp.csr_array(A)), u.MatrixLinearOperator)   # F7 superset
    with pytest.raises(TypeError):
        u.get_linearoperator((3, 3), "nope")
    with pytest.raises(u.LinearOperatorError):
        u.get_linearoperator((4, 4), A)
    op = u.MatrixLinearOperator(A)
    I = u.IdentityLinearOperator((3, 3))
    assert (I * op) is op and (op * I) is op                         # utils.py:1409-1412
    assert isinstance(op * op, u._ProductLinearOperator) and isinstance(2 * op, u._ScaledLinearOperator)
    assert u.find_common_dtype(op, I, np.ones(3, dtype=np.float32)) == np.float64   # F3
    assert u.find_common_dtype(None, "x") == np.float64
    flat, (x,) = u.shape_vecs(np.ones(3))
    assert flat and x.shape == (3, 1)
    g = u.Givens(np.array([[3.0], [-4.0]]))
    assert (g.c, g.s, g.r) == (-0.6, 0.8, -5.0)
    g = u.Givens(np.array([[0.0], [0.0]]))
    assert (g.c, g.s, g.r) == (1.0, 0.0, 0.0)
    import torch
    assert u._compute_dtype(np.complex64) == torch.complex128        # complex runs by real embedding
    assert u._compute_dtype(np.int32) == torch.float64 and u._compute_dtype(np.float32) == torch.float32


def test_givens_host_twin_matches_reference_table():
    import runners
    from krypy_b200 import utils as u
    tab = runners.load_golden("givens_table")["table"]
    for a, b, c, s, r in tab:
        g = u.Givens(np.array([[a], [b]]))
        np.testing.assert_allclose([g.c, g.s, g.r], [c, s, r], rtol=2e-15, atol=0)


def test_row_partition_planning_world2_gloo():
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(ROOT, "tests", "_dist_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


@pytest.mark.parametrize("world", [2, 3, 5, 8])
def test_peer_addressing_contract_in_numpy(world):
    """The multi-GPU exchange addresses a peer's basis row as peer_base + row*ld + halo_off with ONE
    leading dimension for all ranks (block + largest halo).  Emulate the peer-mapped regions with
    numpy arrays and check every rank's extended-vector SpMV against the global product -- middle
    ranks have two-sided halos of a different size than the edge ranks."""
    from krypy_b200 import problems
    from krypy_b200.dist import HaloPlan, RowPartition, local_rows
    A = problems.convdiff2d(19, c=0.1)
    N = A.shape[0]
    parts = [RowPartition(N, world, r) for r in range(world)]
    plans = [HaloPlan(local_rows(A, p), p) for p in parts]
    assert len({p.block for p in parts}) == 1 and sum(p.nloc for p in parts) == N
    ld = parts[0].block + max(pl.nhalo for pl in plans)          # what DistCsrOperator._ext_len enforces
    rows = 4
    rng = np.random.default_rng(0)
    X = rng.standard_normal((rows, N))
    regions = [np.full(rows * ld, np.nan) for _ in range(world)]     # one "cudaMalloc" per rank
    for r, p in enumerate(parts):
        for k in range(rows):
            regions[r][k * ld: k * ld + p.nloc] = X[k, p.lo:p.hi]
    for r, (p, pl) in enumerate(zip(parts, plans)):
        for k in range(rows):
            off = k * ld                                           # identical on every rank
            for i in range(pl.nhalo):                              # kry_halo_gather
                regions[r][off + p.block + i] = regions[pl.halo_peer[i]][off + pl.halo_off[i]]
            xe = regions[r][off: off + pl.ext]
            y = pl.local_matrix() @ np.nan_to_num(xe, nan=0.0)
            assert not np.isnan(xe[: p.nloc]).any() and not np.isnan(xe[p.block: p.block + pl.nhalo]).any()
            assert np.array_equal(y, (A @ X[k])[p.lo:p.hi])


def test_every_entry_point_rejects_a_null_context_without_touching_the_device():
    """C-ABI error convention: int status (0 ok, negative error class) + kry_last_error();
    a NULL context is refused by every entry point before any CUDA call -- safe to exercise without a GPU."""
    import ctypes
    from krypy_b200 import _lib
    lib = _lib.load()
    zero = {ctypes.c_int: 0, ctypes.c_longlong: 0, ctypes.c_double: 0.0}
    called = 0
    for name, (res, args) in sorted(_lib.PROTOTYPES.items()):
        if res is not ctypes.c_int or not args or args[0] is not ctypes.c_void_p or name == "kry_ctx_destroy":
            continue
        rc = getattr(lib, name)(*[zero.get(a, None) for a in args])
        msg = lib.kry_last_error()
        assert rc < 0, (name, rc)
        assert msg and (b"NULL" in msg or b"ctx" in msg or b"requirement failed" in msg), (name, msg)
        called += 1
    assert called >= 30
    assert lib.kry_ctx_destroy(None) == 0 and lib.kry_launch_count(None) == -1
    assert lib.kry_mailbox_host(None) is None and lib.kry_mailbox_dev(None) is None


def test_context_creation_fails_loudly_without_a_device():
    import ctypes
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from krypy_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.kry_ctx_create(0, None, ctypes.byref(h))
    assert rc < 0 and not h.value and lib.kry_last_error()


def test_plain_c_consumer_of_the_abi_compiles_and_links():
    """examples/gmres_c_abi.c: GMRES(m) driven from C through include/krypy_b200.h (no Python, no torch).
    Compile + link only here; tools/gpu_session.sh runs it on the GPU box."""
    import shutil
    import tempfile
    cuda = "/usr/local/cuda"
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("gcc / CUDA runtime headers not available")
    out = os.path.join(tempfile.mkdtemp(), "gmres_c_abi")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "gmres_c_abi.c"), "-o", out, "-L", os.path.join(ROOT, "krypy_b200"),
           "-lkrypy_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + os.path.join(ROOT, "krypy_b200")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and os.path.exists(out), res.stderr[-2000:]
