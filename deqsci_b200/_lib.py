"""ctypes binding of libdeqsci.so (include/deqsci.h).  There is NO fallback: if the CUDA library
is missing or a call fails, the product path raises."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdeqsci.so")

NET_FFDNET, NET_DNCNN = 0, 1
PREC_TC_SPLIT, PREC_FP32, PREC_TC_SINGLE = 0, 1, 2
PRECISIONS = {"tc_split": PREC_TC_SPLIT, "fp32": PREC_FP32, "tc_single": PREC_TC_SINGLE}


class DeqsciError(RuntimeError):
    pass


class SolverOpts(Structure):
    _fields_ = [("m", c_int), ("lam", c_float), ("beta", c_float), ("max_iter", c_int), ("tol", c_float),
                ("sigma0", c_float), ("sigma_decay", c_float), ("sigma_start_call", c_int),
                ("final_call", c_int), ("res_eps", ctypes.c_double)]


class SolverResult(Structure):
    _fields_ = [("residual", ctypes.c_double), ("iterations", c_int), ("f_calls", c_int), ("converged", c_int),
                ("sigma_next", c_float), ("min_sample_residual", ctypes.c_double)]


class BNParams(Structure):
    _fields_ = [("gamma", c_void_p), ("beta", c_void_p), ("running_mean", c_void_p), ("running_var", c_void_p)]


class SavedForward(Structure):
    _fields_ = [("acts", POINTER(c_void_p)), ("pre", POINTER(c_void_p)), ("bn_record", c_void_p), ("zprime", c_void_p)]


class ConvLayer(Structure):
    _fields_ = [("cin", c_int), ("cout", c_int), ("relu", c_int),
                ("weight_host", POINTER(c_float)), ("scale_host", POINTER(c_float)),
                ("bias_host", POINTER(c_float))]


# name -> (restype, argtypes); every symbol include/deqsci.h declares (tests check the list)
_P = c_void_p
SIGNATURES = {
    "deqsci_version": (c_int, []),
    "deqsci_last_error": (c_char_p, []),
    "deqsci_gap_forward": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "deqsci_gap_adjoint": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "deqsci_phi_sum": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "deqsci_gap_step": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "deqsci_gap_vjp": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "deqsci_denoiser_create": (c_int, [c_int, c_int, c_int, POINTER(ConvLayer), POINTER(_P)]),
    "deqsci_denoiser_destroy": (c_int, [_P]),
    "deqsci_denoiser_workspace_bytes": (c_size_t, [_P, c_int, c_int, c_int, c_int]),
    "deqsci_denoise_residual": (c_int, [_P, _P, c_float, _P, _P, c_size_t, c_int, c_int, c_int, c_int, _P]),
    "deqsci_iterate": (c_int, [_P, _P, _P, _P, _P, c_float, _P, _P, c_size_t, c_int, c_int, c_int, c_int, _P]),
    "deqsci_denoiser_update_weights": (c_int, [_P, c_int, POINTER(_P), _P]),
    "deqsci_iterate_train": (c_int, [_P, _P, _P, _P, _P, c_float, _P, _P, c_size_t, POINTER(BNParams), c_float, c_float,
                                     c_int, c_int, c_int, c_int, _P]),
    "deqsci_anderson_scratch_floats": (c_size_t, [c_int, c_int, c_longlong]),
    "deqsci_anderson_update": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_longlong, c_int, c_int,
                                       c_float, c_float, _P]),
    "deqsci_anderson_mix": (c_int, [_P, _P, _P, c_int, c_int, c_longlong, c_int, c_int, c_float, _P]),
    "deqsci_residual": (c_int, [_P, _P, _P, _P, c_longlong, c_float, _P]),
    "deqsci_reconstruct_workspace_bytes": (c_size_t, [_P, c_int, c_int, c_int, c_int, c_int]),
    "deqsci_reconstruct": (c_int, [_P, _P, _P, _P, _P, _P, POINTER(SolverOpts), _P, c_size_t, POINTER(SolverResult),
                                   c_int, c_int, c_int, c_int, _P]),
    "deqsci_reconstruct_train": (c_int, [_P, _P, _P, _P, _P, _P, POINTER(SolverOpts), POINTER(BNParams), c_float,
                                         c_float, _P, c_size_t, POINTER(SolverResult), c_int, c_int, c_int, c_int, _P]),
    "deqsci_adjoint_solve_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "deqsci_adjoint_solve": (c_int, [_P, _P, _P, _P, POINTER(SolverOpts), _P, c_size_t, POINTER(SolverResult),
                                     c_int, c_int, c_int, c_int, _P]),
    "deqsci_denoiser_activation_bytes": (c_size_t, [_P, c_int, c_int, c_int, c_int]),
    "deqsci_iterate_save": (c_int, [_P, _P, _P, _P, _P, c_float, _P, _P, c_size_t, POINTER(SavedForward), c_int, c_int,
                                    c_int, c_int, _P]),
    "deqsci_iterate_train_save": (c_int, [_P, _P, _P, _P, _P, c_float, _P, _P, c_size_t, POINTER(BNParams), c_float,
                                          c_float, POINTER(SavedForward), c_int, c_int, c_int, c_int, _P]),
    "deqsci_backward_workspace_bytes": (c_size_t, [_P, c_int, c_int, c_int, c_int]),
    "deqsci_backward_weights": (c_int, [_P, _P, POINTER(SavedForward), POINTER(_P), _P, c_float, c_float, POINTER(_P),
                                        POINTER(_P), POINTER(_P), _P, c_size_t, c_int, c_int, c_int, c_int, _P]),
    "deqsci_denoise_residual_masked": (c_int, [_P, _P, _P, _P, c_size_t, POINTER(_P), c_int, c_int, c_int, c_int, _P]),
    "deqsci_adjoint_solve_denoiser": (c_int, [_P, POINTER(_P), _P, _P, _P, _P, POINTER(SolverOpts), c_float, _P, c_size_t,
                                              POINTER(SolverResult), c_int, c_int, c_int, c_int, _P]),
    "deqsci_comm_bytes": (c_size_t, [c_longlong]),
    "deqsci_comm_alloc": (c_int, [c_longlong, POINTER(_P), _P]),
    "deqsci_comm_open": (c_int, [_P, POINTER(_P)]),
    "deqsci_comm_close": (c_int, [_P]),
    "deqsci_comm_free": (c_int, [_P]),
    "deqsci_comm_error": (c_int, [_P, c_longlong, POINTER(c_int)]),
    "deqsci_adam_allreduce_step": (c_int, [_P, _P, _P, POINTER(_P), c_int, c_int, c_longlong, c_float, c_float, c_float,
                                           c_float, c_int, c_float, ctypes.c_uint, _P]),
    "deqsci_profile_begin": (c_int, [c_int]),
    "deqsci_profile_end": (c_int, [_P, _P, _P]),
    "deqsci_debug_pair_strip_rows": (c_int, [c_int, c_int, c_int, c_int]),
    "deqsci_debug_pair_weight_map": (ctypes.c_longlong, [_P, ctypes.c_longlong]),
    "deqsci_debug_hidden_layer": (c_int, [_P, c_int, _P, _P, c_int, c_int, c_int, _P]),
}

_lib = None


def lib():
    """Loads libdeqsci.so once.  Raises DeqsciError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DeqsciError(
                "libdeqsci.so is not built (%s). Run `python -m deqsci_b200.build` "
                "(nvcc, sm_100a); there is no CPU or PyTorch fallback for the native path." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status, what=""):
    if status != 0:
        msg = lib().deqsci_last_error()
        raise DeqsciError("%s failed (status %d): %s" % (what or "libdeqsci call", status,
                                                         msg.decode() if msg else "?"))
