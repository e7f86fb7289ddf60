"""ctypes binding of libkrylov_b200.so (the C ABI declared in include/krylov_b200.h).

This is the only place the Python host touches native code.  There is no CPU
fallback: if the shared library is missing the import fails loudly, and every
compute entry point raises :class:`KrylovDeviceError` when no CUDA device is
present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkrylov_b200.so")
if os.environ.get("KRY_B200_LIB"):        # A/B measurement builds of the same library (scripts/gpu_*.sh)
    LIB_PATH = os.path.join(_HERE, os.path.basename(os.environ["KRY_B200_LIB"]))


class KrylovDeviceError(RuntimeError):
    """A call into libkrylov_b200 failed (status code + kry_last_error text)."""

    def __init__(self, status, message):
        RuntimeError.__init__(self, "libkrylov_b200 error %d: %s" % (status, message))
        self.status = status
        self.message = message


KRY_OK = 0
KRY_ERR_INVALID = -1
KRY_ERR_SHAPE = -2
KRY_ERR_CUDA = -3
KRY_ERR_NOMEM = -4
KRY_ERR_UNSUPPORTED = -5
KRY_ERR_COMM = -6
KRY_ERR_STATE = -7

KRY_CSR_SYMMETRIC = 1
KRY_CSR_BUILD_TRANSPOSE = 2
KRY_SPMV_AUTO, KRY_SPMV_ROW, KRY_SPMV_STREAM, KRY_SPMV_TMA, KRY_SPMV_ROWB8, KRY_SPMV_ROWB4 = 0, 1, 2, 3, 4, 5
KRY_SPMV_ROWPF, KRY_SPMV_ROWPF2 = 6, 7
KRY_CG, KRY_BICGSTAB, KRY_CGS, KRY_TFQMR, KRY_MINRES = 1, 2, 3, 4, 5
KRY_NUM_SLOTS = 64
KRY_OPT_L2_HINTS, KRY_OPT_GRAPHS, KRY_OPT_P2P, KRY_OPT_CG_FUSE, KRY_OPT_CG_FUSE_SHARDS = 1, 2, 3, 4, 5
KRY_OPT_CG_ONE_CTA, KRY_OPT_MINRES_FUSE, KRY_OPT_MINRES_PERSISTENT, KRY_OPT_HALO_P2P = 6, 7, 8, 9
KRY_COMM_ID_BYTES = 128

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
handle = C.c_void_p


class SolverParams(C.Structure):
    _fields_ = [("abstol", C.c_double), ("reltol", C.c_double), ("matvec_max", C.c_int64),
                ("check_curvature", C.c_int32), ("guess_supplied", C.c_int32),
                ("shift", C.c_double), ("rtol", C.c_double), ("etol", C.c_double),
                ("window", C.c_int32), ("reserved", C.c_int32)]


class SolverStatus(C.Structure):
    _fields_ = [("done", C.c_int32), ("converged", C.c_int32), ("definite", C.c_int32),
                ("istop", C.c_int32), ("n_matvec", C.c_int64), ("n_iter", C.c_int64),
                ("hist_count", C.c_int64), ("resid_norm0", C.c_double),
                ("resid_norm", C.c_double), ("threshold", C.c_double), ("aux", C.c_double * 16)]


class LlsParams(C.Structure):
    _fields_ = [("window", C.c_int32), ("istop", C.c_int32), ("itn", C.c_int64), ("nmatvec", C.c_int64),
                ("itnlim", C.c_int64), ("damp", C.c_double), ("atol", C.c_double), ("btol", C.c_double),
                ("ctol", C.c_double), ("etol", C.c_double), ("rtol", C.c_double), ("shift", C.c_double),
                ("eps", C.c_double)]


class LlsStatus(C.Structure):
    _fields_ = [("done", C.c_int32), ("istop", C.c_int32), ("itn", C.c_int64), ("nmatvec", C.c_int64),
                ("hist_count", C.c_int64)]


KRY_LLS_LSQR, KRY_LLS_LSMR, KRY_LLS_CRAIG, KRY_LLS_CRAIGMR, KRY_LLS_SYMMLQ = 0, 1, 2, 3, 4
KRY_LLS_HIST_WIDTH = 4


class Axpby(C.Structure):
    _fields_ = [("z", handle), ("u", handle), ("w", handle), ("a", C.c_double), ("b", C.c_double),
                ("a_slot", C.c_int), ("b_slot", C.c_int), ("a_neg", C.c_int), ("b_neg", C.c_int)]


class DotSpec(C.Structure):
    _fields_ = [("u", handle), ("w", handle)]


# name -> (restype, argtypes); every symbol include/krylov_b200.h declares
PROTOTYPES = {
    "kry_abi_version": (C.c_int, []),
    "kry_last_error": (C.c_char_p, []),
    "kry_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "kry_ctx_create": (C.c_int, [C.c_int, C.POINTER(handle)]),
    "kry_ctx_destroy": (C.c_int, [handle]),
    "kry_ctx_sync": (C.c_int, [handle]),
    "kry_ctx_props": (C.c_int, [handle, c_i64p]),
    "kry_timer_start": (C.c_int, [handle]),
    "kry_timer_stop": (C.c_int, [handle, c_f64p]),
    "kry_flush_l2": (C.c_int, [handle]),
    "kry_launch_count": (C.c_int, [handle, c_i64p]),
    "kry_halo_trace_read": (C.c_int, [handle, C.POINTER(C.c_uint64)]),
    "kry_ctx_set_option": (C.c_int, [handle, C.c_int, C.c_int]),
    "kry_ctx_get_option": (C.c_int, [handle, C.c_int, C.POINTER(C.c_int)]),
    "kry_prof_enable": (C.c_int, [handle, C.c_int]),
    "kry_prof_read": (C.c_int, [handle, c_i64p, c_f64p]),
    "kry_host_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "kry_host_free": (C.c_int, [C.c_void_p]),
    "kry_vec_create": (C.c_int, [handle, C.c_int64, C.POINTER(handle)]),
    "kry_vec_create_cap": (C.c_int, [handle, C.c_int64, C.c_int64, C.POINTER(handle)]),
    "kry_vec_destroy": (C.c_int, [handle]),
    "kry_vec_size": (C.c_int, [handle, c_i64p]),
    "kry_vec_upload": (C.c_int, [handle, C.c_void_p, C.c_int64]),
    "kry_vec_download": (C.c_int, [handle, C.c_void_p, C.c_int64]),
    "kry_vec_read": (C.c_int, [handle, C.c_int64, C.c_int64, C.c_void_p]),
    "kry_vec_fill": (C.c_int, [handle, C.c_double]),
    "kry_vec_copy": (C.c_int, [handle, handle]),
    "kry_csr_create": (C.c_int, [handle, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_uint32, C.POINTER(handle)]),
    "kry_csr_destroy": (C.c_int, [handle]),
    "kry_csr_shape": (C.c_int, [handle, c_i64p, c_i64p, c_i64p]),
    "kry_csr_download": (C.c_int, [handle, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kry_csr_build_transpose": (C.c_int, [handle]),
    "kry_csr_diagonal": (C.c_int, [handle, C.c_void_p]),
    "kry_csr_create_poisson1d": (C.c_int, [handle, C.c_int64, C.c_int64, C.c_int64, C.c_uint32,
                                           C.POINTER(handle)]),
    "kry_csr_create_poisson2d": (C.c_int, [handle, C.c_int64, C.c_int64, C.c_int64, C.c_uint32,
                                           C.POINTER(handle)]),
    "kry_csr_create_coo": (C.c_int, [handle, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_uint32, C.POINTER(handle)]),
    "kry_csr_combine": (C.c_int, [handle, handle, C.c_double, handle, C.c_double, C.c_void_p, C.c_double,
                                  C.c_uint32, C.POINTER(handle)]),
    "kry_csr_to_dense": (C.c_int, [handle, C.c_void_p]),
    "kry_csr_create_convdiff3d": (C.c_int, [handle, C.c_int64, C.c_double, C.c_int64, C.c_int64,
                                            C.c_uint32, C.POINTER(handle)]),
    "kry_csr_set_kernel": (C.c_int, [handle, C.c_int, C.c_int, C.c_int]),
    "kry_spmv": (C.c_int, [handle, C.c_int, handle, handle]),
    "kry_spmv_dot": (C.c_int, [handle, C.c_int, handle, handle, C.c_int, C.POINTER(handle), C.c_int]),
    "kry_multi_axpy_dot": (C.c_int, [handle, C.c_int, C.POINTER(Axpby), C.c_int, C.POINTER(DotSpec),
                                     C.c_int]),
    "kry_spmv_axpby_dot": (C.c_int, [handle, C.c_int, handle, C.POINTER(Axpby), C.c_int, handle, C.c_int]),
    "kry_scalars_read": (C.c_int, [handle, C.c_int, C.c_int, c_f64p]),
    "kry_scalars_write": (C.c_int, [handle, C.c_int, C.c_int, c_f64p]),
    "kry_lls_create": (C.c_int, [handle, C.c_int, C.POINTER(handle)]),
    "kry_lls_destroy": (C.c_int, [handle]),
    "kry_lls_scalar_name": (C.c_char_p, [C.c_int, C.c_int]),
    "kry_lls_setup": (C.c_int, [handle, C.POINTER(LlsParams), c_f64p, C.c_int]),
    "kry_lls_step": (C.c_int, [handle, C.c_int]),
    "kry_lls_status": (C.c_int, [handle, C.POINTER(LlsStatus), c_f64p, C.c_int]),
    "kry_lls_history": (C.c_int, [handle, C.c_int64, C.c_int64, C.c_void_p]),
    "kry_lls_multi_axpy_dot": (C.c_int, [handle, C.c_int, C.c_int, C.POINTER(Axpby), C.c_int, C.POINTER(DotSpec)]),
    "kry_lls_spmv_axpby_dot": (C.c_int, [handle, C.c_int, handle, C.c_int, handle, C.POINTER(Axpby), handle]),
    "kry_lls_release_gate": (C.c_int, [handle]),
    "kry_graph_begin": (C.c_int, [handle, C.POINTER(handle)]),
    "kry_graph_end": (C.c_int, [handle]),
    "kry_graph_launch": (C.c_int, [handle, C.c_int]),
    "kry_graph_destroy": (C.c_int, [handle]),
    "kry_solver_create": (C.c_int, [handle, C.c_int, handle, C.POINTER(handle)]),
    "kry_solver_destroy": (C.c_int, [handle]),
    "kry_solver_set_precon_diag": (C.c_int, [handle, C.c_void_p, C.c_int]),
    "kry_solver_setup": (C.c_int, [handle, C.c_void_p, C.c_void_p, C.POINTER(SolverParams)]),
    "kry_solver_setup_dev": (C.c_int, [handle, handle, handle, C.POINTER(SolverParams)]),
    "kry_solver_iterate": (C.c_int, [handle, C.c_int64]),
    "kry_solver_status_read": (C.c_int, [handle, C.POINTER(SolverStatus)]),
    "kry_solver_history": (C.c_int, [handle, C.c_int64, C.c_int64, C.c_void_p, c_i32p]),
    "kry_solver_status_enqueue": (C.c_int, [handle, C.c_int]),
    "kry_solver_status_wait": (C.c_int, [handle, C.c_int, C.POINTER(SolverStatus)]),
    "kry_solver_history_nowait": (C.c_int, [handle, C.c_int64, C.c_int64, C.c_void_p, c_i32p]),
    "kry_solver_solution": (C.c_int, [handle, C.c_void_p]),
    "kry_solver_get_vector": (C.c_int, [handle, C.c_char_p, C.c_void_p]),
    "kry_solver_set_vector": (C.c_int, [handle, C.c_char_p, C.c_void_p]),
    "kry_solver_set_scalar": (C.c_int, [handle, C.c_char_p, C.c_double]),
    "kry_solver_get_scalar": (C.c_int, [handle, C.c_char_p, c_f64p]),
    "kry_comm_unique_id": (C.c_int, [C.c_void_p]),
    "kry_comm_init": (C.c_int, [handle, C.c_int, C.c_int, C.c_void_p]),
    "kry_comm_destroy": (C.c_int, [handle]),
    "kry_comm_size": (C.c_int, [handle, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "kry_comm_barrier": (C.c_int, [handle]),
    "kry_comm_allgather_host": (C.c_int, [handle, C.c_void_p, C.c_void_p, C.c_int64]),
    "kry_comm_allreduce_host": (C.c_int, [handle, c_f64p, C.c_int, C.c_int]),
    "kry_csr_shard_finalize": (C.c_int, [handle, C.c_int64, C.c_int64]),
}

_NO_CHECK = {"kry_abi_version", "kry_last_error"}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libkrylov_b200.so is missing (%s). Build it with "
            "`make -C pykrylov_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the Krylov hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)       # AttributeError here == ABI drift; let it propagate
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def last_error():
    msg = lib.kry_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status):
    """Raise for a negative kry_status (shape errors map to ValueError like the
    reference's LinearOperator._matvec, linop/linop.py:283-296)."""
    if status == KRY_OK:
        return
    msg = last_error()
    if status == KRY_ERR_SHAPE:
        raise ValueError(msg)
    raise KrylovDeviceError(status, msg)


def call(name, *args):
    check(getattr(lib, name)(*args))
