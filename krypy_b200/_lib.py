"""ctypes binding of libkrypy_b200.so (the C ABI declared in include/krypy_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``make -C
krypy_b200/csrc``).  There is no CPU fallback: if the shared object is missing
or no CUDA device is present, the first use raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkrypy_b200.so")

KRY_F32 = 0
KRY_F64 = 1
KRY_ORTH_CGS = 0
KRY_ORTH_MGS = 1
KRY_MAILBOX_DOUBLES = 16384

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_double = ctypes.c_double

# name -> (restype, argtypes); one entry per symbol declared in include/krypy_b200.h
PROTOTYPES = {
    "kry_version": (c_int, []),
    "kry_last_error": (ctypes.c_char_p, []),
    "kry_ctx_create": (c_int, [c_int, c_void_p, ctypes.POINTER(c_void_p)]),
    "kry_ctx_destroy": (c_int, [c_void_p]),
    "kry_ctx_set_stream": (c_int, [c_void_p, c_void_p]),
    "kry_device_info": (c_int, [c_void_p, ctypes.POINTER(c_ll)]),
    "kry_mailbox_host": (c_void_p, [c_void_p]),
    "kry_mailbox_dev": (c_void_p, [c_void_p]),
    "kry_sync": (c_int, [c_void_p]),
    "kry_launch_count": (c_ll, [c_void_p]),
    "kry_reset_launch_count": (None, [c_void_p]),
    "kry_l2_window": (c_int, [c_void_p, c_void_p, c_ll, ctypes.POINTER(c_ll)]),
    "kry_spmv_csr": (c_int, [c_void_p, c_int, c_ll, c_ll, c_ll, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    "kry_gemv_dense": (c_int, [c_void_p, c_int, c_ll, c_ll, c_void_p, c_ll, c_void_p, c_void_p]),
    "kry_diag_mul": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p]),
    "kry_axpby": (c_int, [c_void_p, c_int, c_ll, c_double, c_void_p, c_double, c_void_p, c_void_p]),
    "kry_axpy_dev": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_double, c_void_p, c_void_p]),
    "kry_scale_dev": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_int, c_double, c_void_p, c_void_p]),
    "kry_rot90": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p]),
    "kry_block_dot": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_ll, c_int, c_void_p, c_void_p,
                              c_int, c_void_p]),
    "kry_block_axpy": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_ll, c_int, c_void_p, c_double,
                               c_void_p]),
    "kry_block_combine": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_ll, c_int, c_void_p, c_void_p,
                                  c_void_p]),
    "kry_orth_fused": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int,
                               c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "kry_orth_fused_z": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p]),
    "kry_spmv_csr_z": (c_int, [c_void_p, c_int, c_ll, c_ll, c_ll, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "kry_lanczos_diag": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "kry_lanczos_diag_dist": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "kry_gram": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_ll, c_int, c_void_p, c_ll, c_int, c_void_p]),
    "kry_block_trsm": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_ll]),
    "kry_project": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_void_p,
                            c_void_p, c_void_p, c_int, c_void_p]),
    "kry_givens_update": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "kry_tri_solve": (c_int, [c_void_p, c_int, c_void_p, c_ll, c_void_p, c_void_p]),
    "kry_tri_solve_t": (c_int, [c_void_p, c_int, c_void_p, c_ll, c_void_p, c_void_p]),
    "kry_givens_update_z": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "kry_tri_solve_z": (c_int, [c_void_p, c_int, c_void_p, c_ll, c_void_p, c_void_p]),
    "kry_minres_recur": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int]),
    "kry_minres_update": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "kry_cg_update": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_double, c_void_p, c_int]),
    "kry_cg_update_dev": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "kry_cg_scalars": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "kry_xpby_dev": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kry_peer_alloc": (c_int, [c_void_p, c_ll, ctypes.POINTER(c_void_p)]),
    "kry_peer_free": (c_int, [c_void_p, c_void_p]),
    "kry_ipc_export": (c_int, [c_void_p, c_void_p, ctypes.c_char_p]),
    "kry_ipc_open": (c_int, [c_void_p, ctypes.c_char_p, ctypes.POINTER(c_void_p)]),
    "kry_ipc_close": (c_int, [c_void_p, c_void_p]),
    "kry_halo_gather": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "kry_peer_allreduce": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p,
                                   c_void_p, c_void_p, c_int, c_void_p]),
    "kry_orth_fused_dist": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int,
                                    c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "kry_dist_dot": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_ll, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                             c_void_p, c_void_p]),
    "kry_dist_update": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_int,
                                c_int, c_void_p, c_void_p, c_void_p]),
    "kry_dist_scale": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                               c_void_p, c_void_p]),
    "kry_dist_scale_halo": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_ll,
                                    c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "kry_dist_scale_haloq": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_ll,
                                     c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "kry_spmv_csr_mdot": (c_int, [c_void_p, c_int, c_ll, c_ll, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_ll, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "kry_dist_update_scale": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_void_p, c_int,
                                      c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "kry_dist_halo": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_int, c_int,
                              c_void_p, c_void_p, c_void_p]),
    "kry_small_qr_apply": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kry_peer_barrier": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}


class KryError(RuntimeError):
    """A C-ABI call returned a negative status."""


_lib = None


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "krypy_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C krypy_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


KRY_ERR_UNSUPPORTED = -3      # include/krypy_b200.h


def check(rc):
    if rc != 0:
        msg = load().kry_last_error()
        raise KryError("krypy_b200 C-ABI error %d: %s" % (rc, msg.decode() if msg else "?"))
