"""Native complex128 kernels (csrc/kry_cplx.cu: kry_orth_fused_z, kry_spmv_csr_z) on the GPU: each entry point
against numpy / scipy on the same seeded inputs, the solver classes with the native kernels against the
real-embedding kernels (KRY_NATIVE_Z switch, krypy_b200/utils.py) and against the reference's complex fixtures.

The fixtures themselves run with the native kernels in tests/test_zcomplex_gpu.py (native is the default); this
file adds the kernel-level checks and the native-versus-embedding comparison.  tests/test_znative_host_cpu.py
drives the same functions over the numpy test double on the CPU tier (soundness of this file + host logic).

(File name sorts last: written with a single GPU run left in the round.)"""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

import cases
from krypy_b200 import problems

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    from krypy_b200 import _device
    assert torch.cuda.is_available()
    return _device.Context.get()


def T(ctx, a):
    import torch
    return torch.from_numpy(np.array(a, order="C", copy=True)).to(ctx.device)   # (a copy also over the CPU test double)


def crandn(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


# ------------------------------------------------------------------ kry_orth_fused_z
def _zcgs_ref(V, P, q, j0, passes):
    """block classical Gram-Schmidt with the reference's inner product (krypy/utils.py:183: X^H Y)"""
    q = q.astype(np.complex128).copy()
    h = np.zeros(V.shape[0], dtype=np.complex128)
    for _ in range(passes):
        c = V[j0:].conj() @ q
        h[j0:] += c
        q = q - c @ P[j0:]
    return q, h, float(np.linalg.norm(q))


def _zmgs_ref(V, P, q, j0, passes):
    """the reference's modified Gram-Schmidt loop on complex data (krypy/utils.py:1012-1029)"""
    q = q.astype(np.complex128).copy()
    h = np.zeros(V.shape[0], dtype=np.complex128)
    for _ in range(passes):
        for j in range(j0, V.shape[0]):
            c = np.vdot(V[j], q)
            h[j] += c
            q = q - c * P[j]
    return q, h, float(np.linalg.norm(q))


def _twin_block(ctx, Z):
    """complex (k, n) block in twin storage (krypy_b200.utils._Twin): the layout the solvers hand to the kernel"""
    from krypy_b200 import utils
    return utils._Twin.of(ctx, T(ctx, Z))


def check_orth_fused_z(ctx, algo, passes, nv, j0, n, separate_P=False):
    import torch
    from krypy_b200._lib import KRY_ORTH_CGS, KRY_ORTH_MGS
    rng = np.random.default_rng(1000 * nv + 10 * n + passes + (7 if separate_P else 0))
    V = crandn(rng, max(nv, 1), n) / np.sqrt(2 * n)
    P = crandn(rng, max(nv, 1), n) / np.sqrt(2 * n) if separate_P else V
    q = crandn(rng, n)
    tv = _twin_block(ctx, V)
    tp = _twin_block(ctx, P) if separate_P else tv
    qd = T(ctx, q)
    assert qd.dtype == torch.complex128
    h0 = 0.5 - 0.25j
    h = T(ctx, np.full(nv + 1, h0).view(np.float64))               # interleaved, pre-filled: the kernel accumulates
    vnext = ctx.empty((n,), torch.complex128)
    code = KRY_ORTH_CGS if algo == "cgs" else KRY_ORTH_MGS
    ctx.orth_fused_z(tv.T, tp.T, tv.T.stride(0), j0, nv, qd, passes, code, h.data_ptr(), nrm=h[2 * nv:], vnext=vnext)
    ctx.sync()
    qr, hr, nr = (_zcgs_ref if algo == "cgs" else _zmgs_ref)(V[:nv], P[:nv], q, j0, passes)
    hh = h.cpu().numpy().view(np.complex128)
    np.testing.assert_allclose(hh[j0:nv], h0 + hr[j0:nv], rtol=1e-12, atol=1e-12)
    assert np.all(hh[:j0] == h0)                                    # untouched below j0
    np.testing.assert_allclose(hh[nv].real, nr, rtol=1e-12)
    assert hh[nv].imag == h0.imag                                   # the norm is one double
    np.testing.assert_allclose(qd.cpu().numpy(), qr, rtol=1e-11, atol=1e-12 * np.abs(q).max())
    np.testing.assert_allclose(vnext.cpu().numpy(), qr / nr, rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("algo", ["cgs", "mgs"])
@pytest.mark.parametrize("passes", [1, 2])
@pytest.mark.parametrize("nv,j0", [(1, 0), (2, 0), (3, 0), (5, 0), (8, 0), (9, 0), (17, 0), (31, 0), (32, 0), (12, 3),
                                   (6, 5)])
@pytest.mark.parametrize("n", [1, 7, 4096, 150001])
def test_orth_fused_z(ctx, algo, passes, nv, j0, n):
    check_orth_fused_z(ctx, algo, passes, nv, j0, n)


@pytest.mark.parametrize("algo", ["cgs", "mgs"])
@pytest.mark.parametrize("nv", [1, 2, 4, 7, 16])
def test_orth_fused_z_unrolled_sweeps(ctx, algo, nv):
    """long enough for the stride-unrolled loops of the few-vector sweeps (4 x grid x 256 elements)"""
    check_orth_fused_z(ctx, algo, 1, nv, 0, 4 * ctx.sm_count * 2 * 256 + 12345)


@pytest.mark.parametrize("algo", ["cgs", "mgs"])
def test_orth_fused_z_separate_update_basis(ctx, algo):
    """dots against V, update with P (the preconditioned Arnoldi process, krypy/utils.py:1015, 1027)"""
    check_orth_fused_z(ctx, algo, 1, 6, 0, 30011, separate_P=True)
    check_orth_fused_z(ctx, algo, 2, 11, 2, 5000, separate_P=True)


def test_orth_fused_z_without_tail_and_empty_range(ctx):
    """nrm / vnext absent (preconditioned path), and nv == j0 (norm and normalised store only)"""
    import torch
    from krypy_b200._lib import KRY_ORTH_CGS, KRY_ORTH_MGS
    rng = np.random.default_rng(3)
    n, nv = 20011, 4
    V = crandn(rng, nv, n) / np.sqrt(2 * n)
    q = crandn(rng, n)
    tv = _twin_block(ctx, V)
    for code, ref in ((KRY_ORTH_CGS, _zcgs_ref), (KRY_ORTH_MGS, _zmgs_ref)):
        qd = T(ctx, q)
        h = ctx.scalars(2 * nv + 2)
        ctx.orth_fused_z(tv.T, tv.T, tv.T.stride(0), 0, nv, qd, 1, code, h.data_ptr())
        ctx.sync()
        qr, hr, _ = ref(V, V, q, 0, 1)
        np.testing.assert_allclose(h.cpu().numpy()[: 2 * nv].view(np.complex128), hr, rtol=1e-12, atol=1e-12)
        assert np.all(h.cpu().numpy()[2 * nv:] == 0.0)
        np.testing.assert_allclose(qd.cpu().numpy(), qr, rtol=1e-11, atol=1e-12 * np.abs(q).max())
        qd = T(ctx, q)
        vnext = ctx.empty((n,), torch.complex128)
        ctx.orth_fused_z(tv.T, tv.T, tv.T.stride(0), 2, 2, qd, 1, code, h.data_ptr(), nrm=h[2 * nv:], vnext=vnext)
        ctx.sync()
        assert np.array_equal(qd.cpu().numpy(), q)
        np.testing.assert_allclose(h.cpu().numpy()[2 * nv], np.linalg.norm(q), rtol=1e-13)
        np.testing.assert_allclose(vnext.cpu().numpy(), q / np.linalg.norm(q), rtol=1e-13, atol=1e-15)


def test_orth_fused_z_zero_vector(ctx):
    """norm 0: vnext is zeros, not NaN (as kry_orth_fused)"""
    import torch
    from krypy_b200._lib import KRY_ORTH_CGS
    n = 1000
    rng = np.random.default_rng(4)
    tv = _twin_block(ctx, crandn(rng, 2, n))
    qd = T(ctx, np.zeros(n, dtype=np.complex128))
    h = ctx.scalars(6)
    vnext = T(ctx, np.ones(n, dtype=np.complex128))
    ctx.orth_fused_z(tv.T, tv.T, tv.T.stride(0), 0, 2, qd, 1, KRY_ORTH_CGS, h.data_ptr(), nrm=h[4:], vnext=vnext)
    ctx.sync()
    assert np.all(h.cpu().numpy() == 0.0)
    assert np.all(vnext.cpu().numpy() == 0.0)


def test_orth_fused_z_agrees_with_the_twin_kernel(ctx):
    """the native kernel and the real kernel on the twin rows produce the same coefficients and vectors"""
    import torch
    from krypy_b200._lib import KRY_ORTH_CGS, KRY_ORTH_MGS
    rng = np.random.default_rng(5)
    n, nv = 50021, 9
    V = crandn(rng, nv, n) / np.sqrt(2 * n)
    q = crandn(rng, n)
    tv = _twin_block(ctx, V)
    for code in (KRY_ORTH_CGS, KRY_ORTH_MGS):
        qa, qb = T(ctx, q), T(ctx, q)
        ha, hb = ctx.scalars(2 * nv + 2), ctx.scalars(2 * nv + 2)
        va, vb = ctx.empty((n,), torch.complex128), ctx.empty((n,), torch.complex128)
        ctx.orth_fused_z(tv.T, tv.T, tv.T.stride(0), 0, nv, qa, 1, code, ha.data_ptr(), nrm=ha[2 * nv:], vnext=va)
        ctx.orth_fused(tv.T, tv.T, 0, 2 * nv, qb, 1, code, hb, nrm=hb[2 * nv:], vnext=vb)
        ctx.sync()
        np.testing.assert_allclose(ha.cpu().numpy(), hb.cpu().numpy(), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(va.cpu().numpy(), vb.cpu().numpy(), rtol=1e-11, atol=1e-13)


# ------------------------------------------------------------------ kry_spmv_csr_z
def _zmat(kind, rng, cplx):
    if kind == "lap2d":                                   # 5 entries per row -> staged, 6 per row capacity
        A = problems.laplace2d(37).astype(np.complex128 if cplx else np.float64).tocsr()
        if cplx:
            A = (A + sp.diags(1j * rng.standard_normal(A.shape[0]))).tocsr()
            A.data = A.data * np.exp(1j * rng.standard_normal(A.nnz))
    elif kind == "band7":                                 # 7 per row -> 8 per row capacity
        n = 3001
        A = sp.diags([rng.standard_normal(n - abs(o)) for o in (-30, -2, -1, 0, 1, 2, 30)], (-30, -2, -1, 0, 1, 2, 30)).tocsr()
    elif kind == "rand12":                                # ~12 per row -> 16 per row capacity
        A = sp.random(2500, 2500, density=12 / 2500, random_state=np.random.RandomState(1), format="csr")
    elif kind == "long":                                  # ~60 per row -> warp-per-row kernel
        A = sp.random(700, 900, density=60 / 900, random_state=np.random.RandomState(2), format="csr")
    elif kind == "ragged":                                # empty rows, one heavy tile beyond the stage capacity
        n = 5000
        A = sp.lil_matrix((n, n))
        A.setdiag(rng.standard_normal(n))
        for r in range(300, 320):
            A[r, ::2] = rng.standard_normal(n // 2)       # 2500 entries in each of 20 rows of one tile
        A = A.tocsr()
        A[10:40] = 0
        A.eliminate_zeros()
    elif kind == "tiny":
        A = sp.csr_matrix(np.array([[1.0, 2.0, 0.0], [0.0, 0.0, 0.0], [0.0, 3.0, 4.0]]))
    else:
        raise KeyError(kind)
    A = sp.csr_matrix(A)
    A.sort_indices()
    if cplx and not np.iscomplexobj(A.data):
        A = A.astype(np.complex128)
        A.data = A.data * np.exp(1j * rng.standard_normal(A.nnz))
    return A


def check_spmv_z(ctx, kind, cplx):
    import torch
    rng = np.random.default_rng(hash(kind) % 1000 + (1 if cplx else 0))
    A = _zmat(kind, rng, cplx)
    assert np.iscomplexobj(A.data) == bool(cplx)
    x = crandn(rng, A.shape[1])
    Ad = ctx.upload_csr_z(A)
    assert Ad.vals.dtype == (torch.complex128 if cplx else torch.float64)
    xd = T(ctx, x)
    yd = T(ctx, np.full(A.shape[0], np.nan + 0j))
    ctx.spmv_z(Ad, xd, yd)
    ctx.sync()
    got = yd.cpu().numpy()
    want = A @ x
    scale = abs(A) @ np.abs(x) + 1e-300
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - want) / scale) <= 4e-16 * max(1, int(np.diff(A.indptr).max()))
    return float(np.mean(got == want))


@pytest.mark.parametrize("cplx", [True, False])
@pytest.mark.parametrize("kind", ["lap2d", "band7", "rand12", "long", "ragged", "tiny"])
def test_spmv_z_matches_scipy(ctx, kind, cplx):
    check_spmv_z(ctx, kind, cplx)


def test_spmv_z_many_tiles(ctx):
    """more tiles than resident CTAs: the TMA ring wraps (stage reuse, mbarrier phases)"""
    import torch
    n = 1100                                               # 1.21M rows = 4727 tiles
    A = (problems.laplace2d(n) + 0.3j * sp.identity(n * n)).tocsr()
    rng = np.random.default_rng(8)
    x = crandn(rng, n * n)
    Ad = ctx.upload_csr_z(A)
    yd = ctx.empty((n * n,), torch.complex128)
    ctx.spmv_z(Ad, T(ctx, x), yd)
    ctx.sync()
    want = A @ x
    np.testing.assert_allclose(yd.cpu().numpy(), want, rtol=0, atol=1e-14 * np.abs(want).max())


def test_spmv_z_matches_the_embedded_operator(ctx):
    """MatrixLinearOperator on a complex block: native CSR and real embedding give the same product"""
    from krypy_b200 import utils
    rng = np.random.default_rng(9)
    A = _zmat("lap2d", rng, True)
    X = crandn(rng, A.shape[0], 3)
    got = {}
    for native in (True, False):
        old = utils._NATIVE_Z
        utils._NATIVE_Z = native
        try:
            op = utils.MatrixLinearOperator(A)
            Xd = ctx.to_block(X, utils._compute_dtype(np.complex128))
            got[native] = ctx.to_numpy(op._apply_dev(Xd))
            got[native, "adj"] = ctx.to_numpy(op._apply_dev(Xd, adj=True))
        finally:
            utils._NATIVE_Z = old
    np.testing.assert_allclose(got[True], A @ X, rtol=1e-14, atol=1e-14)
    np.testing.assert_allclose(got[True], got[False], rtol=1e-14, atol=1e-14)
    np.testing.assert_allclose(got[True, "adj"], A.T.conj() @ X, rtol=1e-14, atol=1e-14)


# ------------------------------------------------------------------ solvers: native against embedding
def _solve_complex(native, ortho, restarted=False, precond=False):
    import krypy_b200 as kp
    from krypy_b200 import utils
    rng = np.random.default_rng(17)
    n = 24
    A = (problems.laplace2d(n) - (0.7 + 0.4j) * sp.identity(n * n)).tocsr()        # shifted (Helmholtz-like), complex
    b = crandn(rng, n * n)
    old = utils._NATIVE_Z
    utils._NATIVE_Z = native
    try:
        kw = {}
        if precond:
            kw["M"] = sp.diags(1.0 / (4.0 + 0.1 * rng.random(n * n))).tocsr()
        ls = kp.linsys.LinearSystem(A, b, **kw)
        ctx = kp._device.Context.get()
        ctx.reset_launch_count()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                if restarted:
                    s = kp.linsys.RestartedGmres(ls, maxiter=15, max_restarts=5, tol=1e-9, ortho=ortho)
                else:
                    s = kp.linsys.Gmres(ls, maxiter=60, tol=1e-9, ortho=ortho)
            except kp.utils.ConvergenceError as e:
                s = e.solver
        return np.array(s.resnorms), s.xk, A, b
    finally:
        utils._NATIVE_Z = old


@pytest.mark.parametrize("ortho", ["mgs", "dmgs", "cgs", "cgs2"])
@pytest.mark.parametrize("variant", ["plain", "restarted", "precond"])
def test_native_and_embedding_paths_agree(ctx, ortho, variant):
    kw = dict(restarted=variant == "restarted", precond=variant == "precond")
    ra, xa, A, b = _solve_complex(True, ortho, **kw)
    rb, xb, _, _ = _solve_complex(False, ortho, **kw)
    assert ra.shape == rb.shape
    assert np.all(np.abs(ra - rb) <= 1e-10 * rb + 1e-13)
    np.testing.assert_allclose(xa, xb, rtol=0, atol=1e-9 * np.abs(xb).max())
    if variant != "precond":
        assert np.linalg.norm(b.reshape(-1, 1) - A @ xa) <= 1.001 * ra[-1] * np.linalg.norm(b) + 1e-12


def test_native_kernels_are_the_default_and_take_fewer_passes(ctx):
    """the switch defaults to native, and the native step launches the complex kernels (by launch count: the
    SpMV and the Gram-Schmidt kernel are one launch each on both paths, so equal counts with different kernels
    is what a correct switch gives; the kernel names are in the ncu launch list of the profile tool)"""
    from krypy_b200 import utils
    import os
    assert utils._NATIVE_Z == (os.environ.get("KRY_NATIVE_Z", "1") not in ("0", ""))


def test_complex_cg_and_minres_with_native_matrix(ctx):
    """CG / MINRES use real coefficients only (alpha.real): with the native matrix format the SpMV is
    kry_spmv_csr_z and the fused <p, Ap> epilogue of the real kernel is replaced by a block dot"""
    name_cg = [c for c in cases.COMPLEX_CASES if "cg" in c]
    name_mr = [c for c in cases.COMPLEX_CASES if "minres" in c]
    assert name_cg and name_mr
    import test_zcomplex_gpu as z
    for name in name_cg[:1] + name_mr[:1]:
        z.test_complex_cases_match_reference_fixture_and_oracle(name)


def test_arnoldi_more_vectors_than_one_native_call(ctx):
    """k + 1 > 32 complex vectors with block Gram-Schmidt: chunked kry_orth_fused_z calls"""
    from krypy_b200 import utils
    rng = np.random.default_rng(21)
    N, m = 400, 40
    A = sp.random(N, N, density=0.02, random_state=np.random.RandomState(3), format="csr").astype(np.complex128)
    A.data = A.data * np.exp(1j * rng.standard_normal(A.nnz))
    A = (A + 3 * sp.identity(N)).tocsr()
    v = crandn(rng, N, 1)
    for ortho in ("cgs2", "dmgs", "mgs"):
        arn = utils.Arnoldi(A, v, maxiter=m, ortho=ortho)
        while arn.iter < m and not arn.invariant:
            arn.advance()
        V, H = arn.get()
        k = H.shape[1]
        np.testing.assert_allclose(A @ V[:, :k], V @ H, rtol=0, atol=1e-12 * np.abs(H).max())
        if ortho != "mgs":     # (one-pass MGS loses orthogonality on this non-normal matrix on every path, 4e-5)
            np.testing.assert_allclose(V.conj().T @ V, np.eye(V.shape[1]), rtol=0, atol=1e-11)
