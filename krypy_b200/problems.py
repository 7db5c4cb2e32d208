"""Synthetic sparse systems of SURVEY.md section 8(d): direct stencil -> CSR
builders (sorted int32 column indices, no scipy.sparse.kron temporaries).

All builders return host ``scipy.sparse.csr_matrix`` objects (what the
reference accepts, krypy/utils.py:250-251) assembled from flat numpy arrays.
``rows=(lo, hi)`` returns only a contiguous block of rows with GLOBAL column
indices (shape ``(hi-lo, N)``), which is what a row-partitioned rank holds.
"""
import numpy as np
import scipy.sparse as sp

__all__ = [
    "stencil_csr", "laplace2d", "poisson3d", "convdiff2d", "shifted_laplace_B",
    "rhs_normal", "jacobi_csr",
]


def stencil_csr(dims, offsets, coeffs, dtype=np.float64, rows=None, rowscale=None,
                diag_shift=0.0):
    """CSR matrix of a constant-coefficient stencil on a ``dims`` grid
    (last index fastest), homogeneous Dirichlet boundaries.

    offsets: list of per-dimension integer offset tuples, listed in ascending
    column order; coeffs: matching values.
    """
    dims = tuple(int(d) for d in dims)
    N = int(np.prod(dims))
    lo, hi = (0, N) if rows is None else rows
    n = hi - lo
    idx = np.arange(lo, hi, dtype=np.int64)
    # per-dimension coordinates of every row
    coords = []
    rem = idx.copy()
    for d in reversed(dims):
        coords.append(rem % d)
        rem //= d
    coords = coords[::-1]
    strides = [int(np.prod(dims[i + 1:])) for i in range(len(dims))]
    ns = len(offsets)
    cols = np.empty((n, ns), dtype=np.int32)
    vals = np.empty((n, ns), dtype=dtype)
    mask = np.empty((n, ns), dtype=bool)
    for s, (off, c) in enumerate(zip(offsets, coeffs)):
        ok = np.ones(n, dtype=bool)
        lin = 0
        for d, o in enumerate(off):
            if o:
                cc = coords[d] + o
                ok &= (cc >= 0) & (cc < dims[d])
                lin += o * strides[d]
        mask[:, s] = ok
        cols[:, s] = (idx + lin).astype(np.int32)
        v = c + (diag_shift if all(o == 0 for o in off) else 0.0)
        vals[:, s] = v
    if rowscale is not None:
        vals *= np.asarray(rowscale, dtype=dtype)[lo:hi, None]
    indptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(mask.sum(axis=1), out=indptr[1:])
    A = sp.csr_matrix((vals[mask], cols[mask], indptr), shape=(n, N))
    A.has_sorted_indices = True
    return A


def laplace2d(n, dtype=np.float64, rows=None, shift=0.0, rowscale=None):
    """5-point Laplacian, diag 4, off-diagonals -1 (config C2; C5 with shift)."""
    offs = [(-1, 0), (0, -1), (0, 0), (0, 1), (1, 0)]
    return stencil_csr((n, n), offs, [-1.0, -1.0, 4.0, -1.0, -1.0], dtype, rows,
                       rowscale, diag_shift=-shift)


def poisson3d(n, dtype=np.float64, rows=None):
    """7-point Poisson, diag 6, off-diagonals -1 (config C3)."""
    offs = [(-1, 0, 0), (0, -1, 0), (0, 0, -1), (0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0)]
    return stencil_csr((n, n, n), offs, [-1.0, -1.0, -1.0, 6.0, -1.0, -1.0, -1.0],
                       dtype, rows)


def convdiff2d(n, c=0.1, dtype=np.float64, rows=None):
    """A = I (x) (T+C) + (T+C/2) (x) I with T=tridiag(-1,2,-1),
    C=c*tridiag(-1,0,1): nonsymmetric convection-diffusion (config C4)."""
    offs = [(-1, 0), (0, -1), (0, 0), (0, 1), (1, 0)]
    co = [-1.0 - 0.5 * c, -1.0 - c, 4.0, -1.0 + c, -1.0 + 0.5 * c]
    return stencil_csr((n, n), offs, co, dtype, rows)


def shifted_laplace_B(n, sigma=0.3, dtype=np.float32, rows=None):
    """Config C5: B = diag(linspace(1,2,N)) (SPD, CSR), A = B^{-1}(L - sigma I),
    self-adjoint in <x,y>_B.  Returns (A, B)."""
    N = n * n
    bdiag = np.linspace(1.0, 2.0, N)
    A = laplace2d(n, dtype=dtype, rows=rows, shift=sigma, rowscale=1.0 / bdiag)
    lo, hi = (0, N) if rows is None else rows
    m = hi - lo
    B = sp.csr_matrix((bdiag[lo:hi].astype(dtype), np.arange(lo, hi, dtype=np.int32),
                       np.arange(m + 1, dtype=np.int32)), shape=(m, N))
    return A, B


def jacobi_csr(A):
    """M = diag(A)^{-1} as a CSR diagonal matrix (config C3's preconditioner)."""
    d = A.diagonal()
    n = d.shape[0]
    return sp.csr_matrix((1.0 / d, np.arange(n, dtype=np.int32),
                          np.arange(n + 1, dtype=np.int32)), shape=(n, n))


def rhs_normal(N, dtype=np.float64, seed=0):
    """b = default_rng(seed).standard_normal((N,1))."""
    return np.random.default_rng(seed).standard_normal((N, 1)).astype(dtype)
