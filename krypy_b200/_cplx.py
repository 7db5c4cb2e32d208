"""Complex arithmetic by real embedding.

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
    """N x M (complex or real) scipy sparse -> 2N x 2M real CSR of the embedding"""
    import scipy.sparse as sp
    A = sp.csr_matrix(A)
    E = sp.kron(A.real.astype(np.float64), I2, format="csr")
    if np.iscomplexobj(A.data):
        E = E + sp.kron(sp.csr_matrix(A.imag.astype(np.float64)), J2, format="csr")
    E = sp.csr_matrix(E)
    E.sort_indices()
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
