"""Complex arithmetic by real embedding.

(The two kernels of the Arnoldi hot loop have native complex versions -- kry_orth_fused_z works on the even rows
of the twin storage described here and never reads the twins, kry_spmv_csr_z takes the complex matrix as it is;
krypy_b200/utils.py ``_NATIVE_Z``, DESIGN.md section 4a.  Everything else in the complex path is what follows.)

The kernels are real.  A complex vector of length N is stored as 2N interleaved reals
``[re0, im0, re1, im1, ...]`` and a complex matrix entry ``a+ib`` becomes the 2x2 block
``[[a, -b], [b, a]]`` (so the expanded real operator applied to an embedded vector IS the complex
product, term by term).  Complex inner products need one extra ingredient: for every vector ``v``
the orthogonalisation works against, its twin ``i*v`` is stored next to it (``kry_rot90``), because

    <v, q>_C = <v, q>_R + i <i v, q>_R          and          (a + i b) v = a v + b (i v),

i.e. complex (modified or classical) Gram-Schmidt against ``k`` vectors is real Gram-Schmidt against
the ``2k`` real vectors ``v_0, i v_0, v_1, i v_1, ...`` and the real coefficient array it produces
is the interleaved complex coefficient array.  This keeps the Krylov iterates those of the complex
algorithm (unlike solving the 2N x 2N real-equivalent system).  Cost: the basis is read twice.
"""
import numpy as np

I2 = np.array([[1.0, 0.0], [0.0, 1.0]])
J2 = np.array([[0.0, -1.0], [1.0, 0.0]])


def expand_sparse(A):
    """N x M (complex or real) scipy sparse -> 2N x 2M real CSR of the embedding (int32 indices,
    sorted).  Direct O(nnz) construction -- ``scipy.sparse.kron`` needs several times the memory of
    the result, which matters at the sizes this library is for.  Row 2i holds ``(2j: re, 2j+1: -im)``,
    row 2i+1 ``(2j: im, 2j+1: re)`` for every entry (i, j)."""
    import scipy.sparse as sp
    A = sp.csr_matrix(A)
    if not A.has_sorted_indices:
        A = A.sorted_indices()
    N, M = A.shape
    nnz = A.nnz
    if 4 * nnz >= 2 ** 31 - 1:
        raise NotImplementedError("embedded matrix needs 64-bit row pointers (not built)")
    indptr = A.indptr.astype(np.int64)
    counts = np.diff(indptr)
    re = np.ascontiguousarray(A.data.real, dtype=np.float64)
    im = np.ascontiguousarray(A.data.imag, dtype=np.float64) if np.iscomplexobj(A.data) else np.zeros(nnz)
    rows = np.repeat(np.arange(N, dtype=np.int64), counts)
    t = np.arange(nnz, dtype=np.int64) - indptr[rows]            # position within the row
    even = 4 * indptr[rows] + 2 * t                                # destination of (2j) in row 2i
    odd = even + 2 * counts[rows]                                  # ... in row 2i+1
    cols = np.empty(4 * nnz, dtype=np.int32)
    vals = np.empty(4 * nnz, dtype=np.float64)
    c2 = (2 * A.indices.astype(np.int64)).astype(np.int32)
    cols[even], cols[even + 1], cols[odd], cols[odd + 1] = c2, c2 + 1, c2, c2 + 1
    vals[even], vals[even + 1], vals[odd], vals[odd + 1] = re, -im, im, re
    ip2 = np.zeros(2 * N + 1, dtype=np.int64)
    ip2[1::2] = 4 * indptr[:-1] + 2 * counts
    ip2[2::2] = 4 * indptr[1:]
    E = sp.csr_matrix((vals, cols, ip2.astype(np.int32)), shape=(2 * N, 2 * M))
    E.has_sorted_indices = True
    return E


def expand_dense(A):
    """N x M (complex or real) ndarray -> 2N x 2M real ndarray of the embedding"""
    A = np.asarray(A)
    E = np.kron(A.real.astype(np.float64), I2)
    if np.iscomplexobj(A):
        E = E + np.kron(A.imag.astype(np.float64), J2)
    return np.ascontiguousarray(E)


def to_pairs(z):
    """complex array (..., k) -> float64 array (..., 2k) interleaved"""
    z = np.ascontiguousarray(z, dtype=np.complex128)
    return z.view(np.float64).reshape(z.shape[:-1] + (2 * z.shape[-1],))


def from_pairs(r):
    """float64 array (..., 2k) interleaved -> complex array (..., k)"""
    r = np.ascontiguousarray(r, dtype=np.float64)
    return r.view(np.complex128).reshape(r.shape[:-1] + (r.shape[-1] // 2,))
