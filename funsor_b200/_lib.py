"""ctypes binding of include/funsor_b200.h.  There is no CPU fallback: if the shared library is
missing the import of any compute entry point raises, and every call checks the int status."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FB2_LIB_PATH") or os.path.join(_HERE, "lib", "libfunsor_b200.so")  # FB2_LIB_PATH: A/B builds of the same ABI

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE, ERR_NO_DEVICE, ERR_NUMERIC, ERR_ABORTED = range(8)
SEMIRING_LOG, SEMIRING_MAX = 0, 1


class Fb2Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__("funsor_b200 status {}: {}".format(status, message))
        self.status = status


class Fb2Unsupported(Fb2Error, NotImplementedError):
    pass


class Fb2Aborted(Fb2Error):
    """an earlier asynchronous dataflow pass on this device gave up: its outputs are NaN (include/funsor_b200.h, fb2_async_status)"""


_fp = c_void_p  # device/host pointers are passed as integers
_i64p = POINTER(c_int64)

# name -> (restype, argtypes); mirrors include/funsor_b200.h one to one
SIGNATURES = {
    "fb2_version": (c_int, []),
    "fb2_status_string": (c_char_p, [c_int]),
    "fb2_last_error": (c_char_p, []),
    "fb2_async_status": (c_int, [c_int]),
    "fb2_device_query": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_size_t)]),
    "fb2_semiring_matmul_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32, c_int32, c_int32]),
    "fb2_semiring_matmul_f32": (c_int, [c_int, _fp, _i64p, _fp, _i64p, _fp, _fp, c_int64, c_int64, c_int32, c_int32, c_int32, _fp, c_size_t, c_void_p]),
    "fb2_chain_product_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "fb2_chain_product_gemm_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "fb2_chain_product_f32": (c_int, [c_int, _fp, _i64p, c_int64, c_int64, c_int32, _fp, _fp, c_size_t, c_void_p]),
    "fb2_hmm_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64, c_int32]),
    "fb2_hmm_forward_f32": (c_int, [c_int, _fp, c_int64, _fp, c_int64, c_int64, _fp, c_int64, c_int64, c_int32, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_hmm_viterbi_f32": (c_int, [_fp, c_int64, _fp, c_int64, c_int64, _fp, c_int64, c_int64, c_int32, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_hmm_marginals_f32": (c_int, [_fp, c_int64, _fp, c_int64, c_int64, _fp, c_int64, c_int64, c_int32, _fp, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_hmm_chunk_summary_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "fb2_hmm_chunk_summary_f32": (c_int, [c_int, _fp, c_int64, c_int64, _fp, c_int64, c_int64, c_int32, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_hmm_chunk_summary_gemm_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int64]),
    "fb2_hmm_chunk_summary_gemm_f32": (c_int, [_fp, _fp, c_int64, c_int32, c_int64, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_hmm_backward_sample_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32, c_int64, c_int64]),
    "fb2_hmm_backward_sample_f32": (c_int, [_fp, c_int64, _fp, c_int64, c_int64, _fp, _fp, _fp, c_int64, c_int64, c_int32, c_int64, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_chain_adjoint_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "fb2_chain_adjoint_f32": (c_int, [c_int, _fp, _i64p, _fp, _fp, c_int64, c_int64, c_int32, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_gaussian_sqrt_to_info_f32": (c_int, [_fp, _fp, _fp, c_int64, c_int32, c_int32, _fp, _fp, _fp, c_void_p]),
    "fb2_gaussian_info_to_sqrt_f32": (c_int, [_fp, _fp, _fp, c_int64, c_int32, c_float, _fp, _fp, _fp, _fp, c_void_p]),
    "fb2_gaussian_moment_match_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "fb2_gaussian_moment_match_f32": (c_int, [_fp, _fp, _fp, _fp, c_int64, c_int32, c_int32, _fp, _fp, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_gaussian_pair_combine_f32": (c_int, [_fp, _fp, _fp, c_int64, _fp, _fp, _fp, c_int64, c_int64, c_int64, c_int64, c_int64, c_int32, _fp, _fp, _fp, _fp, c_void_p]),
    "fb2_gaussian_kalman_eta_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, c_int64, c_int64, c_int32, c_int32, _fp, _fp, c_void_p]),
    "fb2_gaussian_kalman_constants_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, c_int64, c_int32, c_int32, _fp, _fp, _fp, _fp, _fp, _fp, c_void_p]),
    "fb2_gaussian_kalman_chain_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "fb2_gaussian_kalman_chain_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, c_int64, c_int64, c_int32, c_int32, _fp, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_gaussian_eta_f32": (c_int, [_fp, _fp, _fp, c_int64, c_int64, c_int32, c_int32, _fp, _fp, c_void_p]),
    "fb2_gaussian_chain_homog_f32": (c_int, [_fp, _fp, _fp, c_int64, c_int64, c_int32, _fp, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_gaussian_chain_sample_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "fb2_gaussian_chain_homog_keep_f32": (c_int, [_fp, _fp, _fp, c_int64, c_int64, c_int32, _fp, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_gaussian_chain_homog_backward_sample_f32": (c_int, [_fp, c_int64, c_int64, c_int32, c_int64, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_gaussian_chain_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "fb2_gaussian_chain_f32": (c_int, [_fp, _fp, _fp, c_int64, c_int64, c_int32, _fp, _fp, _fp, _fp, c_size_t, c_void_p]),
    "fb2_hmm_forward_f32_host": (c_int, [c_int, _fp, c_int64, _fp, c_int64, c_int64, _fp, c_int64, c_int64, c_int32, _fp]),
    "fb2_chain_product_f32_host": (c_int, [c_int, _fp, c_int64, c_int64, c_int32, _fp]),
    "fb2_host_arena_release": (c_int, []),
    "fb2_launch_count": (c_int64, []),
    "fb2_launch_count_reset": (None, []),
}

_lib = None


def load():
    """Load (once) and return the ctypes library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "funsor_b200: {} is missing. Build it with `python -m funsor_b200.build` "
            "(there is no CPU fallback for this path).".format(LIB_PATH)
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the binary disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status):
    if status == OK:
        return
    lib = load()
    msg = lib.fb2_last_error().decode() or lib.fb2_status_string(status).decode()
    if status == ERR_INVALID:
        raise ValueError("funsor_b200: " + msg)
    if status == ERR_UNSUPPORTED:
        raise Fb2Unsupported(status, msg)
    if status == ERR_NUMERIC:
        raise ValueError("funsor_b200: " + msg)
    if status == ERR_ABORTED:
        lib.fb2_async_status(1)  # reported once; the caller decides whether to recompute
        raise Fb2Aborted(status, msg)
    raise Fb2Error(status, msg)


def strides4(values):
    return (c_int64 * 4)(*[int(v) for v in values])
