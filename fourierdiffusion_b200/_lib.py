"""ctypes binding of libfdiff_b200.so (C ABI: include/fdiff_b200.h).

There is deliberately no fallback: if the shared library is missing or no B200 is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfdiff_b200.so")

FD_MODEL_TRANSFORMER, FD_MODEL_LSTM, FD_MODEL_MLP = 0, 1, 2
FD_SCHED_VP, FD_SCHED_VE = 0, 1
FD_MATH_FP32, FD_MATH_TF32 = 0, 1
FD_ABI_VERSION = 1


class FdConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("device", C.c_int32),
        ("model_kind", C.c_int32),
        ("max_len", C.c_int32),
        ("n_channels", C.c_int32),
        ("d_model", C.c_int32),
        ("n_head", C.c_int32),
        ("num_layers", C.c_int32),
        ("d_ff", C.c_int32),
        ("sched_kind", C.c_int32),
        ("sched_p0", C.c_double),
        ("sched_p1", C.c_double),
        ("fourier_noise_scaling", C.c_int32),
        ("math_mode", C.c_int32),
    ]


# every symbol include/fdiff_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_F = C.c_void_p  # float* passed as raw addresses (tensor.data_ptr())
SYMBOLS = {
    "fd_abi_version": (C.c_int, []),
    "fd_last_error": (C.c_char_p, []),
    "fd_create": (C.c_int, [C.POINTER(FdConfig), C.POINTER(_P)]),
    "fd_destroy": (C.c_int, [_P]),
    "fd_set_weight": (C.c_int, [_P, C.c_char_p, _F, C.c_int64]),
    "fd_finalize_weights": (C.c_int, [_P]),
    "fd_score": (C.c_int, [_P, _F, C.c_float, _F, C.c_int32, _P]),
    "fd_step": (C.c_int, [_P, _F, _F, _F, C.c_double, C.c_float, _F, C.c_int32, _P]),
    "fd_prior": (C.c_int, [_P, _F, _F, C.c_int32, _P]),
    "fd_ffn_block": (C.c_int, [_P, C.c_int32, _F, C.c_int32, _P]),
    "fd_attention_block": (C.c_int, [_P, C.c_int32, _F, C.c_int32, _P]),
    "fd_encoder_stack": (C.c_int, [_P, _F, C.c_int32, _P]),
    "fd_normal": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint32, _F, C.c_int32, _P]),
    "fd_score_t": (C.c_int, [_P, _F, _F, _F, C.c_int32, _P]),
    "fd_perturb": (C.c_int, [_P, _F, _F, _F, _F, _F, C.c_int32, _P]),
    "fd_sde_loss": (C.c_int, [_P, _F, _F, _F, C.c_int32, C.c_int32, _F, _F, C.c_int32, _P]),
    "fd_sample": (C.c_int, [_P, C.c_int32, C.c_int32, _F, C.c_float, C.c_uint64, C.c_uint64, _F, _F, _F, _P]),
    "fd_sample_host": (C.c_int, [_P, C.c_int32, C.c_int32, _F, C.c_float, C.c_uint64, C.c_uint64, _F, _F, _F, _P]),
    "fd_dft": (C.c_int, [_F, _F, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "fd_idft": (C.c_int, [_F, _F, C.c_int32, C.c_int32, C.c_int32, _F, _F, C.c_int32, _P]),
    "fd_spectral_density": (C.c_int, [_F, _F, _F, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "fd_wasserstein": (C.c_int, [_F, _F, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, _P]),
    "fd_feature_stats": (C.c_int, [_F, _F, _F, C.c_int64, C.c_int32, C.c_int32, _P]),
    "fd_standardise": (C.c_int, [_F, _F, _F, _F, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P]),
    "fd_launch_count": (C.c_int64, [_P]),
    "fd_global_launch_count": (C.c_int64, []),
    "fd_active_path": (C.c_int, [_P]),
    "fd_set_option": (C.c_int, [_P, C.c_char_p, C.c_int32]),
    "fd_debug_stack_stats": (C.c_int, [_P, C.c_void_p, C.c_int32]),
    "fd_debug_abort_record": (C.c_int, [C.c_void_p]),
    "fd_stack_task_table": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "fd_profile_enable": (C.c_int, [_P, C.c_int32]),
    "fd_profile_ms": (C.c_double, [_P, C.c_char_p]),
    "fd_profile_launches": (C.c_int64, [_P, C.c_char_p]),
}

_lib = None
_lock = threading.Lock()


class FdError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once) and declare the prototypes.  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise FdError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). fourierdiffusion_b200 has no CPU or PyTorch fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.fd_abi_version() != FD_ABI_VERSION:
            raise FdError(f"ABI mismatch: library {lib.fd_abi_version()} vs binding {FD_ABI_VERSION}")
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().fd_last_error()
        raise FdError(msg.decode("utf-8", "replace") if msg else f"fdiff_b200 call failed with code {rc}")
