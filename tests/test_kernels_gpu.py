"""Per-entry-point parity of the C ABI (include/krypy_b200.h) against numpy /
the oracle, on the GPU.  Integer/index work and the fp64 SpMV are bit-exact;
floating-point reductions are compared to 1e-13 relative (fp64) / 2e-6 (fp32),
the tolerance of a reordered sum of the sizes used here."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

RT64 = 1e-13
RT32 = 3e-6


@pytest.fixture(scope="module")
def ctx():
    import torch
    from krypy_b200 import _device
    assert torch.cuda.is_available()
    return _device.Context.get()


def T(ctx, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)


def tdt(dt):
    import torch
    return torch.float64 if dt == np.float64 else torch.float32


def basis(ctx, rng, nv, n, dt, pad=True):
    """(nv, n) basis with padded leading dimension like the product uses."""
    ld = (n + 31) // 32 * 32 if pad else n
    host = rng.standard_normal((nv, n)).astype(dt)
    store = ctx.zeros((nv, ld), tdt(dt))
    store[:, :n].copy_(T(ctx, host))
    return host, store[:, :n]


# ---------------------------------------------------------------- SpMV
@pytest.mark.parametrize("n", [1, 3, 17, 100, 257])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_spmv_stencil_bitexact(ctx, n, dt):
    from krypy_b200 import problems
    from oracle import krylov_oracle as ko
    A = problems.convdiff2d(n, c=0.1, dtype=dt)
    N = n * n
    rng = np.random.default_rng(n)
    x = rng.standard_normal(N).astype(dt)
    Ad = ctx.upload_csr(A, tdt(dt))
    xd, yd = T(ctx, x), ctx.empty((N,), tdt(dt))
    ctx.spmv(Ad, xd, yd)
    y = yd.cpu().numpy()
    # oracle: sequential per-row sum, separately rounded (scipy csr_matvec order), in double
    ref = ko.csr_matvec(N, A.indptr, A.indices, A.data.astype(np.float64), x.astype(np.float64))
    if dt == np.float64:
        assert np.array_equal(y, ref)
        assert np.array_equal(y, A.dot(x))
    else:
        assert np.array_equal(y, ref.astype(np.float32))


@pytest.mark.parametrize("avg,n", [(3, 1000), (12, 3000), (25, 2000), (70, 1500)])
def test_spmv_random_modes(ctx, avg, n):
    rng = np.random.default_rng(avg)
    A = sp.random(n, n, density=avg / n, format="csr", random_state=rng, dtype=np.float64)
    A.sort_indices()
    x = rng.standard_normal(n)
    Ad = ctx.upload_csr(A, tdt(np.float64))
    yd = ctx.empty((n,), tdt(np.float64))
    w = rng.standard_normal(n)
    dot = ctx.scalars(1)
    ctx.spmv(Ad, T(ctx, x), yd, w=T(ctx, w), dot_out=dot)
    ref = A.dot(x)
    np.testing.assert_allclose(yd.cpu().numpy(), ref, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(dot.item(), np.dot(w, ref), rtol=1e-11, atol=1e-11)


def test_spmv_long_row_in_short_matrix(ctx):
    """a tile whose nnz exceeds the shared-memory stage falls back to direct loads"""
    n = 5000
    rng = np.random.default_rng(5)
    A = sp.eye(n, format="lil", dtype=np.float64)
    A[123, :] = rng.standard_normal(n)          # 5000-entry row inside a diagonal matrix
    A[4000, :3000] = 1.5
    A = A.tocsr()
    x = rng.standard_normal(n)
    Ad = ctx.upload_csr(A, tdt(np.float64))
    yd = ctx.empty((n,), tdt(np.float64))
    ctx.spmv(Ad, T(ctx, x), yd)
    assert np.array_equal(yd.cpu().numpy(), A.dot(x))


def test_spmv_empty_rows_and_rect(ctx):
    rng = np.random.default_rng(6)
    A = sp.random(700, 300, density=0.004, format="csr", random_state=rng)
    A.sort_indices()
    x = rng.standard_normal(300)
    Ad = ctx.upload_csr(A, tdt(np.float64))
    yd = ctx.empty((700,), tdt(np.float64))
    ctx.spmv(Ad, T(ctx, x), yd)
    assert np.array_equal(yd.cpu().numpy(), A.dot(x))


def test_spmv_dot_epilogue_stencil(ctx):
    from krypy_b200 import problems
    A = problems.poisson3d(21)
    N = A.shape[0]
    p = np.random.default_rng(7).standard_normal(N)
    Ad = ctx.upload_csr(A, tdt(np.float64))
    pd, Apd, out = T(ctx, p), ctx.empty((N,), tdt(np.float64)), ctx.scalars(1)
    ctx.spmv(Ad, pd, Apd, w=pd, dot_out=out)
    Ap = A.dot(p)
    assert np.array_equal(Apd.cpu().numpy(), Ap)
    np.testing.assert_allclose(out.item(), p.dot(Ap), rtol=RT64)


@pytest.mark.parametrize("nb", [0, 1, 5, 8, 9, 16, 17, 31, 32])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_spmv_mdot_epilogue(ctx, nb, dt):
    """kry_spmv_csr_mdot: y bit-identical to the plain SpMV, c[j] = <B[j], y>, c[nb] = <y, y> (all three
    accumulator tiles, all three staged row-length classes)"""
    from krypy_b200 import problems
    rng = np.random.default_rng(100 + nb)
    rt = RT64 if dt == np.float64 else RT32
    for A in (problems.laplace2d(61, dtype=dt), problems.poisson3d(17, dtype=dt),
              sp.random(3000, 3000, density=12.0 / 3000, format="csr", random_state=rng, dtype=dt)):
        A = sp.csr_matrix(A)
        A.sort_indices()
        N = A.shape[0]
        x = rng.standard_normal(N).astype(dt)
        B = rng.standard_normal((max(nb, 1), N)).astype(dt)
        Ad = ctx.upload_csr(A, tdt(dt))
        xd, yd, y0 = T(ctx, x), ctx.empty((N,), tdt(dt)), ctx.empty((N,), tdt(dt))
        Bd = T(ctx, B)
        out = ctx.scalars(nb + 1)
        assert ctx.spmv_mdot(Ad, xd, yd, Bd, nb, 1, out) is True
        ctx.spmv(Ad, xd, y0)
        y = yd.cpu().numpy()
        assert np.array_equal(y, y0.cpu().numpy())
        y64 = y.astype(np.float64)
        ref = np.concatenate([B[:nb].astype(np.float64) @ y64, [y64 @ y64]])
        scale = np.concatenate([np.linalg.norm(B[:nb].astype(np.float64), axis=1) * np.linalg.norm(y64),
                                [y64 @ y64]])
        assert np.all(np.abs(out.cpu().numpy() - ref) <= 50 * RT64 * scale), (nb, dt)
        # deterministic: a second launch gives the same bits
        out2 = ctx.scalars(nb + 1)
        ctx.spmv_mdot(Ad, xd, yd, Bd, nb, 1, out2)
        assert np.array_equal(out.cpu().numpy(), out2.cpu().numpy())
    # long rows: no staged path, the caller is told (no launch, no fallback inside the library)
    Al = sp.random(500, 500, density=0.2, format="csr", random_state=rng, dtype=dt)
    Ad = ctx.upload_csr(Al, tdt(dt))
    assert ctx.spmv_mdot(Ad, T(ctx, x[:500].copy()), ctx.empty((500,), tdt(dt)), T(ctx, B[:, :500].copy()),
                         min(nb, 1), 1, ctx.scalars(2)) is False


def test_gemv_diag(ctx):
    rng = np.random.default_rng(8)
    for dt, rt in ((np.float64, RT64), (np.float32, RT32)):
        A = rng.standard_normal((37, 53)).astype(dt)
        x = rng.standard_normal(53).astype(dt)
        yd = ctx.empty((37,), tdt(dt))
        ctx.gemv(T(ctx, A), T(ctx, x), yd)
        np.testing.assert_allclose(yd.cpu().numpy(), A.astype(np.float64) @ x, rtol=rt * 10, atol=rt * 10)
        d = rng.standard_normal(53).astype(dt)
        zd = ctx.empty((53,), tdt(dt))
        ctx.diag_mul(T(ctx, d), T(ctx, x), zd)
        assert np.array_equal(zd.cpu().numpy(), d * x)


# ---------------------------------------------------------------- elementwise
@pytest.mark.parametrize("n", [1, 2, 5, 1023, 4099])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_elementwise(ctx, n, dt):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(dt)
    y = rng.standard_normal(n).astype(dt)
    rt = RT64 if dt == np.float64 else RT32
    xd, yd, zd = T(ctx, x), T(ctx, y), ctx.empty((n,), tdt(dt))
    ctx.axpby(1.0, xd, -1.0, yd, zd)
    assert np.array_equal(zd.cpu().numpy(), x - y)
    ctx.axpby(0.75, xd, 2.5, yd, zd)
    np.testing.assert_allclose(zd.cpu().numpy(), 0.75 * x.astype(np.float64) + 2.5 * y, rtol=rt, atol=rt)
    ctx.axpby(-3.0, xd, 0.0, None, zd)
    assert np.array_equal(zd.cpu().numpy(), (-3.0 * x).astype(dt))
    c = T(ctx, np.array([0.3, 7.0]))
    y2 = T(ctx, y)
    ctx.axpy_dev(c, -1.0, xd, y2)
    np.testing.assert_allclose(y2.cpu().numpy(), y - 0.3 * x.astype(np.float64), rtol=rt, atol=rt)
    ctx.scale_dev(c[1:], 1, -1.0, xd, zd)
    assert np.array_equal(zd.cpu().numpy(), (-x.astype(np.float64) / 7.0).astype(dt))
    # unaligned views take the scalar path
    if n > 3:
        xs, zs = xd[1:], zd[1:]
        ctx.axpby(2.0, xs, 0.0, None, zs)
        assert np.array_equal(zs.cpu().numpy(), (2.0 * x[1:]).astype(dt))


# ---------------------------------------------------------------- tall-skinny
@pytest.mark.parametrize("nv", [1, 7, 8, 9, 33, 70])
@pytest.mark.parametrize("n", [5, 1000, 70001])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_block_dot_axpy_combine(ctx, nv, n, dt):
    rng = np.random.default_rng(nv * 1000 + n)
    rt = RT64 if dt == np.float64 else RT32
    Vh, Vd = basis(ctx, rng, nv, n, dt)
    q = rng.standard_normal(n).astype(dt)
    qd = T(ctx, q)
    out = ctx.scalars(nv)
    acc = T(ctx, np.ones(nv))
    ctx.block_dot(Vd, nv, qd, out, 0, acc)
    ref = Vh.astype(np.float64) @ q.astype(np.float64)
    scale = np.abs(Vh.astype(np.float64)) @ np.abs(q.astype(np.float64)) + 1e-300
    assert np.max(np.abs(out.cpu().numpy() - ref) / scale) < rt
    np.testing.assert_allclose(acc.cpu().numpy(), 1.0 + out.cpu().numpy(), rtol=1e-15)
    coef = rng.standard_normal(nv)
    cd = T(ctx, coef)
    ctx.block_axpy(Vd, nv, cd, -1.0, qd)
    refq = q.astype(np.float64) - coef @ Vh.astype(np.float64)
    mag = np.abs(q) + np.abs(coef) @ np.abs(Vh.astype(np.float64))
    assert np.max(np.abs(qd.cpu().numpy() - refq) / mag) < (rt if dt == np.float64 else 2e-7)
    x0 = rng.standard_normal(n).astype(dt)
    od = ctx.empty((n,), tdt(dt))
    ctx.block_combine(Vd, nv, cd, T(ctx, x0), od)
    refc = x0 + coef @ Vh.astype(np.float64)
    assert np.max(np.abs(od.cpu().numpy() - refc) / (mag + np.abs(x0))) < (rt if dt == np.float64 else 2e-7)
    ctx.block_combine(Vd, nv, cd, None, od)
    assert np.max(np.abs(od.cpu().numpy() - coef @ Vh.astype(np.float64)) / mag) < (rt if dt == np.float64 else 2e-7)


def test_block_dot_sqrt_post(ctx):
    rng = np.random.default_rng(3)
    x = rng.standard_normal(12345)
    xd = T(ctx, x)
    out = ctx.scalars(1)
    ctx.block_dot(xd.reshape(1, -1), 1, xd, out, 1, None)
    np.testing.assert_allclose(out.item(), np.linalg.norm(x), rtol=RT64)


# ---------------------------------------------------------------- fused Gram-Schmidt
def _mgs_ref(V, P, q, j0, passes, pre=None):
    """krypy/utils.py:1000-1034 in numpy (double)."""
    q = q.astype(np.float64).copy()
    h = np.zeros(V.shape[0])
    if pre is not None:
        q -= pre[0] * pre[1]
    for _ in range(passes):
        for j in range(j0, V.shape[0]):
            a = V[j] @ q
            h[j] += a
            q -= a * P[j]
    return q, h, np.linalg.norm(q)


def _cgs_ref(V, P, q, j0, passes, pre=None):
    q = q.astype(np.float64).copy()
    h = np.zeros(V.shape[0])
    if pre is not None:
        q -= pre[0] * pre[1]
    for _ in range(passes):
        c = V[j0:] @ q
        h[j0:] += c
        q -= c @ P[j0:]
    return q, h, np.linalg.norm(q)


@pytest.mark.parametrize("algo", ["cgs", "mgs"])
@pytest.mark.parametrize("passes", [1, 2])
@pytest.mark.parametrize("nv,j0", [(1, 0), (5, 0), (17, 0), (31, 0), (40, 3), (6, 5)])
@pytest.mark.parametrize("n", [7, 4096, 150001])
def test_orth_fused(ctx, algo, passes, nv, j0, n):
    from krypy_b200._lib import KRY_ORTH_CGS, KRY_ORTH_MGS
    rng = np.random.default_rng(nv * 7 + n + passes)
    Vh, Vd = basis(ctx, rng, nv, n, np.float64)
    Vh /= np.sqrt(n)
    Vd.copy_(T(ctx, Vh))
    q = rng.standard_normal(n)
    qd = T(ctx, q)
    h = T(ctx, np.full(nv + 1, 0.5))
    vnext = ctx.empty((n,), tdt(np.float64))
    code = KRY_ORTH_CGS if algo == "cgs" else KRY_ORTH_MGS
    ctx.orth_fused(Vd, Vd, j0, nv, qd, passes, code, h, nrm=h[nv:], vnext=vnext)
    ref = (_cgs_ref if algo == "cgs" else _mgs_ref)(Vh, Vh, q, j0, passes)
    qr, hr, nr = ref
    hh = h.cpu().numpy()
    np.testing.assert_allclose(hh[j0:nv], 0.5 + hr[j0:nv], rtol=1e-12, atol=1e-12)
    assert np.all(hh[:j0] == 0.5)
    np.testing.assert_allclose(hh[nv], nr, rtol=1e-12)
    np.testing.assert_allclose(qd.cpu().numpy(), qr, rtol=1e-11, atol=1e-12 * np.abs(q).max())
    np.testing.assert_allclose(vnext.cpu().numpy(), qr / nr, rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_orth_lanczos_with_P_basis(ctx, dt):
    """Lanczos step: pre-subtraction, one dot against V[k], update with P[k] (utils.py:1000-1029)"""
    from krypy_b200._lib import KRY_ORTH_MGS
    n, k = 30011, 4
    rng = np.random.default_rng(11)
    Vh, Vd = basis(ctx, rng, k + 2, n, dt)
    Ph, Pd = basis(ctx, rng, k + 2, n, dt)
    q = rng.standard_normal(n).astype(dt)
    qd = T(ctx, q)
    lz = T(ctx, np.array([0.37, 0.0, 0.0]))
    h_ptr = lz.data_ptr() + 8 * (1 - k)
    ctx.orth_fused(Vd, Pd, k, k + 1, qd, 1, KRY_ORTH_MGS, None, nrm=None, vnext=None,
                   pre_vec=Pd[k - 1], pre_coef=lz, h_ptr=h_ptr)
    qr, hr, _ = _mgs_ref(Vh[:k + 1].astype(np.float64), Ph[:k + 1].astype(np.float64), q, k, 1,
                         pre=(0.37, Ph[k - 1].astype(np.float64)))
    rt = 1e-12 if dt == np.float64 else 2e-5
    np.testing.assert_allclose(lz.cpu().numpy()[1], hr[k], rtol=rt * 100, atol=rt * 100)
    np.testing.assert_allclose(qd.cpu().numpy(), qr, rtol=rt * 1e3, atol=rt * 10 * np.abs(qr).max())


def test_orth_fused_f32_cgs(ctx):
    from krypy_b200._lib import KRY_ORTH_CGS
    n, nv = 100003, 12
    rng = np.random.default_rng(12)
    Vh, Vd = basis(ctx, rng, nv, n, np.float32)
    Vh = (Vh / np.sqrt(n)).astype(np.float32)
    Vd.copy_(T(ctx, Vh))
    q = rng.standard_normal(n).astype(np.float32)
    qd = T(ctx, q)
    h = ctx.scalars(nv + 1)
    vnext = ctx.empty((n,), tdt(np.float32))
    ctx.orth_fused(Vd, Vd, 0, nv, qd, 1, KRY_ORTH_CGS, h, nrm=h[nv:], vnext=vnext)
    qr, hr, nr = _cgs_ref(Vh.astype(np.float64), Vh.astype(np.float64), q, 0, 1)
    np.testing.assert_allclose(h.cpu().numpy()[:nv], hr, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(h.cpu().numpy()[nv], nr, rtol=1e-6)
    np.testing.assert_allclose(vnext.cpu().numpy(), qr / nr, rtol=1e-4, atol=1e-6)


# ---------------------------------------------------------------- deflation projector
@pytest.mark.parametrize("d", [1, 5, 20, 33])
@pytest.mark.parametrize("n", [50, 20011])
@pytest.mark.parametrize("iterations", [1, 2, 3])
def test_project_matches_oracle(ctx, d, n, iterations):
    import scipy.linalg
    if d > n:
        pytest.skip("d > n")
    rng = np.random.default_rng(d + n)
    W, _ = np.linalg.qr(rng.standard_normal((n, d)))
    V, _ = np.linalg.qr(W + 0.3 * rng.standard_normal((n, d)))
    Q, R = scipy.linalg.qr(W.T @ V)
    a = rng.standard_normal(n)
    # oracle: krypy/utils.py:604-627 with _apply :540-549
    z = a.copy()
    c_first = None
    for it in range(iterations):
        c = W.T @ z
        if it == 0:
            c_first = c.copy()
        c = scipy.linalg.solve_triangular(R, Q.T @ c)
        z = z - V @ c
    Wd = ctx.zeros((d, (n + 31) // 32 * 32), tdt(np.float64))[:, :n]
    Vd = ctx.zeros((d, (n + 31) // 32 * 32), tdt(np.float64))[:, :n]
    Wd.copy_(T(ctx, W.T.copy()))
    Vd.copy_(T(ctx, V.T.copy()))
    ad = T(ctx, a)
    cf = ctx.scalars(d)
    ctx.project(Wd, Vd, d, ad, T(ctx, Q), T(ctx, R), iterations, cf)
    np.testing.assert_allclose(cf.cpu().numpy(), c_first, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(ad.cpu().numpy(), z, rtol=1e-10, atol=1e-11)


# ---------------------------------------------------------------- small recurrences
def test_givens_update_matches_oracle_and_table(ctx):
    from oracle import krylov_oracle as ko
    import runners
    rng = np.random.default_rng(21)
    m = 40
    H = np.triu(rng.standard_normal((m + 1, m)), -1)
    y = np.zeros((m + 1, 1)); y[0] = 3.7
    R = np.zeros((m + 1, m)); G = []
    yd = ctx.scalars(m + 2); yd[0:1].fill_(3.7)
    cs = ctx.scalars(2 * m + 2); rcol = ctx.scalars(m + 2); hcol = ctx.scalars(m + 2)
    for k in range(m):
        R[:k + 2, k] = H[:k + 2, k]
        for i in range(k):
            R[i:i + 2, k] = G[i].dot(R[i:i + 2, k])
        G.append(ko.givens(R[k:k + 2, [k]])[3])
        R[k:k + 2, k] = G[k].dot(R[k:k + 2, k])
        y[k:k + 2] = G[k].dot(y[k:k + 2])
        hcol[:k + 2].copy_(T(ctx, H[:k + 2, k].copy()))
        ctx.givens_update(k, hcol, rcol, cs, yd, 0)
        ctx.sync()
        mb = ctx.mailbox
        np.testing.assert_allclose(mb[0], abs(y[k + 1, 0]), rtol=1e-12, atol=1e-15)
        assert np.array_equal(mb[1:k + 3], H[:k + 2, k])
        np.testing.assert_allclose(mb[k + 3:2 * k + 5], R[:k + 2, k], rtol=1e-11, atol=1e-13)
        assert np.all(hcol[:k + 2].cpu().numpy() == 0)
    # drotg table generated by the reference (tests/golden/givens_table.npz)
    tab = runners.load_golden("givens_table")["table"]
    for a, b, c, s, r in tab:
        hcol[:2].copy_(T(ctx, np.array([a, b])))
        yd.zero_(); yd[0:1].fill_(1.0)
        ctx.givens_update(0, hcol, rcol, cs, yd, 0)
        ctx.sync()
        got = cs[:2].cpu().numpy()
        np.testing.assert_allclose(got, [c, s], rtol=4e-16, atol=0)
        np.testing.assert_allclose(rcol[0].item(), r, rtol=2e-15, atol=0)   # r = c*a + s*b: 2 roundings


def test_tri_solve(ctx):
    import scipy.linalg
    rng = np.random.default_rng(22)
    for k in (1, 2, 30, 200):
        R = np.triu(rng.standard_normal((k, k))) + 5 * np.eye(k)
        y = rng.standard_normal(k)
        out = ctx.scalars(k)
        ctx.tri_solve(k, T(ctx, R), T(ctx, y), out)
        np.testing.assert_allclose(out.cpu().numpy(), scipy.linalg.solve_triangular(R, y), rtol=1e-11)
        # column-after-column storage with a padded leading dimension (the device-resident R of a restart cycle):
        # the same operations in the same order, so the same bits
        Rt = np.zeros((k, k + 2))
        Rt[:, :k] = R.T
        out_t = ctx.scalars(k)
        ctx.tri_solve_t(k, T(ctx, Rt), T(ctx, y), out_t)
        assert np.array_equal(out_t.cpu().numpy(), out.cpu().numpy())


def test_minres_recur_and_update(ctx):
    from oracle import krylov_oracle as ko
    rng = np.random.default_rng(23)
    n, steps = 5003, 12
    al = rng.standard_normal(steps); be = np.abs(rng.standard_normal(steps + 1)) + 0.1
    st = ctx.scalars(16); st[6:7].fill_(2.5)
    h3 = ctx.scalars(3)
    y = [2.5, 0]; G1 = G2 = None
    W = np.zeros((n, 2)); yk = np.zeros(n)
    W0, W1, ykd = ctx.zeros((n,), tdt(np.float64)), ctx.zeros((n,), tdt(np.float64)), ctx.zeros((n,), tdt(np.float64))
    for k in range(steps):
        v = rng.standard_normal(n)
        Rr = np.zeros((4, 1)); Rr[1] = be[k] if k > 0 else 0.0
        if G1 is not None: Rr[:2] = G1.dot(Rr[:2])
        Rr[2:4, 0] = [al[k], be[k + 1]]
        if G2 is not None: Rr[1:3] = G2.dot(Rr[1:3])
        G1 = G2
        c_, s_, r_, G2 = ko.givens(Rr[2:4]); Rr[2] = r_; Rr[3] = 0
        y = G2.dot(y)
        z = (v - Rr[0, 0] * W[:, 0] - Rr[1, 0] * W[:, 1]) / Rr[2, 0]
        W = np.column_stack([W[:, 1], z]); yk = yk + y[0] * z; y = [y[1], 0]
        h3[1:3].copy_(T(ctx, np.array([al[k], be[k + 1]])))
        ctx.minres_recur(k, h3, st, 1, 0)
        ctx.minres_update(T(ctx, v), W0, W1, ykd, st)
        W0, W1 = W1, W0
        ctx.sync()
        mb = ctx.mailbox
        np.testing.assert_allclose(mb[0], abs(y[0]), rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose(mb[1:4], Rr[:3, 0], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(ykd.cpu().numpy(), yk, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(h3.cpu().numpy()[:2], [be[k + 1], 0.0])


@pytest.mark.parametrize("n", [3, 1000, 99991])
@pytest.mark.parametrize("use_d", [False, True])
def test_cg_update(ctx, n, use_d):
    rng = np.random.default_rng(n)
    p, Ap, yk, r = (rng.standard_normal(n) for _ in range(4))
    d = np.abs(rng.standard_normal(n)) + 0.5
    rho, pap = 2.3, 0.7
    alpha = rho / pap
    pd, Apd, ykd, rd = T(ctx, p), T(ctx, Ap), T(ctx, yk), T(ctx, r)
    zd = ctx.empty((n,), tdt(np.float64))
    ctx.cg_update(Apd, pd, ykd, rd, zd if use_d else None, T(ctx, d) if use_d else None, rho,
                  T(ctx, np.array([pap])), 0)
    ctx.sync()
    r2 = r - alpha * Ap
    z2 = d * r2 if use_d else r2
    np.testing.assert_allclose(ykd.cpu().numpy(), yk + alpha * p, rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(rd.cpu().numpy(), r2, rtol=1e-12, atol=1e-14)
    if use_d:
        np.testing.assert_allclose(zd.cpu().numpy(), z2, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ctx.mailbox[0], r2 @ z2, rtol=1e-12)
    np.testing.assert_allclose(ctx.mailbox[1], alpha, rtol=1e-15)


@pytest.mark.parametrize("n", [3, 1000, 99991])
@pytest.mark.parametrize("use_d", [False, True])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_cg_device_scalars_recurrence(ctx, n, use_d, dt):
    """kry_cg_update_dev + kry_cg_scalars + kry_xpby_dev: one CG step whose scalars never leave the
    device, against numpy with the same rounding points (linsys.py:627-665)"""
    import torch
    rng = np.random.default_rng(n + use_d)
    p, Ap, yk, r, pold = (rng.standard_normal(n).astype(dt) for _ in range(5))
    d = (np.abs(rng.standard_normal(n)) + 0.5).astype(dt)
    rho, pap = 2.3, 0.7
    st = torch.zeros(8, dtype=torch.float64, device=ctx.device)
    st[1], st[2] = rho, pap
    pd, Apd, ykd, rd = T(ctx, p), T(ctx, Ap), T(ctx, yk), T(ctx, r)
    zd = ctx.empty((n,), tdt(dt))
    ctx.cg_update_dev(Apd, pd, ykd, rd, zd if use_d else None, T(ctx, d) if use_d else None, st)
    ctx.cg_scalars(st, 0)
    out = ctx.empty((n,), tdt(dt))
    ctx.xpby_dev(zd if use_d else rd, st[4:], T(ctx, pold), out)
    ctx.sync()
    alpha = rho / pap
    f = np.float64
    r2 = (r.astype(f) - alpha * Ap.astype(f)).astype(dt)
    z2 = (d.astype(f) * r2.astype(f)).astype(dt) if use_d else r2
    rt = 1e-12 if dt == np.float64 else 2e-6
    np.testing.assert_allclose(ykd.cpu().numpy(), (yk.astype(f) + alpha * p.astype(f)).astype(dt), rtol=rt, atol=rt)
    np.testing.assert_allclose(rd.cpu().numpy(), r2, rtol=rt, atol=rt)
    s = float(r2.astype(f) @ z2.astype(f))
    stt = st.cpu().numpy()
    np.testing.assert_allclose(ctx.mailbox[0], s, rtol=1e-12 if dt == np.float64 else 1e-5)
    assert stt[0] == rho and stt[3] == alpha == ctx.mailbox[1] and ctx.mailbox[2] == pap
    raw = float(ctx.mailbox[0])
    assert stt[1] == np.sqrt(abs(raw)) ** 2                      # the reference squares the norm
    assert stt[4] == stt[1] / rho
    zin = (zd if use_d else rd).cpu().numpy()                               # (the device's own z: r used an FMA)
    want = ((zin.astype(f) + (stt[4] * pold.astype(f)))).astype(dt)         # numpy: z + (beta * p)
    got = out.cpu().numpy()
    if dt == np.float64:
        assert np.array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,kx,ky", [(1, 2, 2), (127, 3, 5), (4099, 20, 20), (70001, 7, 32), (33000, 32, 4)])
def test_gram_and_block_trsm(ctx, dt, n, kx, ky):
    """kry_gram (X^H Y and X^H X in one pass) and kry_block_trsm (X R^-1) against numpy/scipy"""
    import scipy.linalg
    import torch
    rng = np.random.default_rng(n + kx)
    Xh, X = basis(ctx, rng, kx, n, dt)
    Yh, Y = basis(ctx, rng, ky, n, dt)
    out = torch.zeros(kx * ky, dtype=torch.float64, device=ctx.device)
    ctx.gram(X, kx, Y, ky, out)
    ref = Xh.astype(np.float64) @ Yh.astype(np.float64).T
    scale = np.abs(Xh).astype(np.float64) @ np.abs(Yh).astype(np.float64).T + 1e-300
    assert np.all(np.abs(out.cpu().numpy().reshape(kx, ky) - ref) <= 1e-14 * scale * max(np.log2(n + 1), 1))
    assert ctx.gram_fits(kx, ky, False)
    if ctx.gram_fits(kx, kx, True):
        out2 = torch.zeros(kx * kx, dtype=torch.float64, device=ctx.device)
        ctx.gram(X, kx, X, kx, out2)                                  # same-block (syrk-like) staging
        G = out2.cpu().numpy().reshape(kx, kx)
        ref2 = Xh.astype(np.float64) @ Xh.astype(np.float64).T
        assert np.all(np.abs(G - ref2) <= 1e-14 * (np.abs(Xh).astype(np.float64) @ np.abs(Xh).astype(np.float64).T + 1e-300)
                      * max(np.log2(n + 1), 1))
        assert np.array_equal(out2.cpu().numpy(), ctx_gram_again(ctx, X, kx))      # deterministic
    R = np.triu(rng.standard_normal((kx, kx))) + 4.0 * np.eye(kx)
    Q = ctx.zeros(tuple(X.shape), tdt(dt))
    ctx.block_trsm(X, kx, T(ctx, R), Q)
    want = scipy.linalg.solve_triangular(R, Xh.astype(np.float64), trans="T", lower=False)
    np.testing.assert_allclose(Q.cpu().numpy(), want.astype(dt), rtol=RT64 * 100 if dt == np.float64 else RT32 * 10,
                               atol=(1e-12 if dt == np.float64 else 1e-5) * np.abs(want).max())
    ctx.block_trsm(X, kx, T(ctx, R), X)                           # in place
    assert np.array_equal(X.cpu().numpy(), Q.cpu().numpy())


def ctx_gram_again(ctx, X, kx):
    import torch
    o = torch.zeros(kx * kx, dtype=torch.float64, device=ctx.device)
    ctx.gram(X, kx, X, kx, o)
    return o.cpu().numpy()
