"""ctypes binding of libdgprf.so (include/dgprf.h).

PyTorch is used only for device memory and streams: every compute call below hands raw
``data_ptr()`` addresses to the C ABI.  There is no CPU fallback -- if the shared library is
missing, or no CUDA device is present, compute calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DGPRF_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libdgprf.so")   # override: A/B builds

MAX_LAYERS = 8
MAX_SEGMENTS = 64
KIND_RBF, KIND_ARC = 0, 1
LIK_GAUSSIAN, LIK_SOFTMAX = 0, 1
PREC_FP32, PREC_TF32 = 0, 1
MODE_EVAL, MODE_TRAIN, MODE_HYPER = 0, 1, 2

c_float_p = C.POINTER(C.c_float)
LAYER_HOOK = C.CFUNCTYPE(None, C.c_int, C.c_void_p)          # dgprf_layer_hook


class Layer(C.Structure):
    _fields_ = [("kind", C.c_int32), ("d_prev", C.c_int32), ("d_x", C.c_int32), ("M", C.c_int32),
                ("g", C.c_int32), ("has_mean", C.c_int32),
                ("off_W", C.c_int64), ("off_log_amp", C.c_int64), ("off_log_inv_ls", C.c_int64),
                ("off_mean", C.c_int64), ("z", C.c_void_p), ("z_cs", C.c_int64)]


class Model(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("likelihood", C.c_int32), ("d_in", C.c_int32),
                ("d_out", C.c_int32), ("n_chains", C.c_int32), ("precision", C.c_int32),
                ("w_base", C.c_void_p), ("w_cs", C.c_int64), ("h_base", C.c_void_p), ("h_cs", C.c_int64),
                ("off_lik_log_var", C.c_int64), ("layer", Layer * MAX_LAYERS)]


class Segment(C.Structure):
    _fields_ = [("offset", C.c_int64), ("length", C.c_int64), ("mass", C.c_float), ("flags", C.c_int32)]


class DgprfError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes); every symbol include/dgprf.h declares
_VP, _I, _I64, _U64, _F, _SZ = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_size_t
_MP, _SP = C.POINTER(Model), C.POINTER(Segment)
SIGNATURES = {
    "dgprf_last_error": (C.c_char_p, []),
    "dgprf_version": (_I, []),
    "dgprf_workspace_bytes": (_I, [_MP, _I, _I, C.POINTER(_SZ)]),
    "dgprf_forward": (_I, [_MP, _VP, _I64, _I, _I, _VP, _SZ, _VP, _VP]),
    "dgprf_loglik": (_I, [_MP, _VP, _I64, _I, _I, _VP, _SZ, _VP, _VP, _VP, _F, _VP]),
    "dgprf_backward": (_I, [_MP, _VP, _I64, _I, _I, _VP, _SZ, _VP]),
    "dgprf_grad_finalize": (_I, [_MP, _I, _I, _VP, _SZ, _VP, _I64, _VP, _I64, _F, _I, _VP]),
    "dgprf_grad_finalize_layer": (_I, [_MP, _I, _I, _I, _VP, _SZ, _VP, _I64, _F, _VP]),
    "dgprf_set_backward_hook": (_I, [_VP, _VP]),
    "dgprf_peer_allreduce": (_I, [_VP, _VP, _I, _I, _I64, C.c_uint32, C.c_uint32, _VP]),
    "dgprf_peer_allreduce_status": (_I, [C.POINTER(C.c_uint32)]),
    "dgprf_gradients": (_I, [_MP, _VP, _I64, _VP, _I64, _I, _I, _VP, _SZ, _VP, _I64, _VP, _I64, _F, _I, _VP, _F, _I, _VP]),
    "dgprf_sgmcmc_update": (_I, [_VP, _VP, _I64, _I64, _I, _VP, _I64, _I, _I64, _SP, _I,
                                 _F, _F, _F, _F, _I, _U64, _U64, _VP, _VP, _VP]),
    "dgprf_sgmcmc_step": (_I, [_MP, _VP, _I64, _VP, _I64, _I, _I,
                               _VP, _VP, _I64, _SP, _I, _VP, _VP, _I64, _SP, _I,
                               _F, _F, _F, _F, _I, _U64, _U64, _VP, _VP, _VP, _VP,
                               _VP, _SZ, _VP, _VP]),
    "dgprf_sgmcmc_step_graph": (_I, [_MP, _VP, _I64, _VP, _I64, _I, _I,
                                     _VP, _VP, _I64, _SP, _I, _VP, _VP, _I64, _SP, _I,
                                     _F, _F, _F, _F, _I, _U64, _U64, _VP,
                                     _VP, _SZ, _VP, _VP]),
    "dgprf_sgmcmc_step_host": (_I, [_MP, _VP, _VP, _I, _I, _VP, _VP, _I, _I,
                                    _VP, _VP, _I64, _SP, _I, _VP, _VP, _I64, _SP, _I,
                                    _F, _F, _F, _F, _I, _U64, _U64, _VP, _SZ, _VP, _VP, _VP]),
    "dgprf_log_prior": (_I, [_VP, _I64, _I64, _I, _VP, _VP]),
    "dgprf_predictive_reduce": (_I, [_VP, _VP, _I, _I64, _I64, _F, _I, _VP, _VP, _VP, _VP]),
    "dgprf_adam_step": (_I, [_VP, _VP, _VP, _VP, _I64, _F, _F, _F, _F, _I, _VP]),
    "dgprf_welford_update": (_I, [_VP, _VP, _VP, _I64, _I, _VP]),
    "dgprf_mass_estimate": (_I, [_VP, _VP, _SP, _I, _I, _I, _VP, _VP]),
    "dgprf_rf_features": (_I, [_I, _VP, _I, _I, _VP, _VP, _VP, _VP, _I, _VP, _VP]),
    "dgprf_gp_matmul": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    "dgprf_gaussian_log_prob": (_I, [_VP, _VP, _VP, _I, _I, _VP, _VP]),
    "dgprf_softmax_log_prob": (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP]),
    "dgprf_philox_normal": (_I, [_VP, _I64, _U64, _U64, _U64, _I, _VP]),
    "dgprf_profile_start": (_I, []),
    "dgprf_profile_stop": (_I, [_I, C.c_char_p, C.POINTER(C.c_float), C.POINTER(_I)]),
}


def lib() -> C.CDLL:
    """Load libdgprf.so (built in-tree by ``__graft_entry__.build()``); fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DgprfError(f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
                             "(there is no CPU / PyTorch fallback for the hot path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().dgprf_last_error()
        raise DgprfError(f"libdgprf error {rc}: {msg.decode() if msg else '?'}")


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise DgprfError("dgprf needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Raw device address of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device (the raw getter is ~10x cheaper than building a
    torch.cuda.Stream object: this sits on the per-step path of the sampler)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def as_dev(x, device, dtype=torch.float32) -> torch.Tensor:
    """Tensor plumbing: anything array-like -> contiguous fp32 tensor on `device`."""
    if hasattr(x, "tensor") and not torch.is_tensor(x):
        x = x.tensor
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    x = x.detach()
    if x.dtype != dtype or x.device != device:
        x = x.to(device=device, dtype=dtype)
    return x.contiguous()


def profile_start():
    check(lib().dgprf_profile_start())


def profile_stop(max_records=4096):
    """-> list of (kernel name, milliseconds) in launch order."""
    names = C.create_string_buffer(32 * max_records)
    ms = (C.c_float * max_records)()
    n = C.c_int(0)
    check(lib().dgprf_profile_stop(max_records, names, ms, C.byref(n)))
    raw = names.raw
    return [(raw[32 * i:32 * i + 32].split(b"\0", 1)[0].decode(), float(ms[i])) for i in range(n.value)]


def make_segments(entries) -> "C.Array[Segment]":
    arr = (Segment * max(len(entries), 1))()
    for i, (off, length, mass, flags) in enumerate(entries):
        arr[i].offset, arr[i].length, arr[i].mass, arr[i].flags = int(off), int(length), float(mass), int(flags)
    return arr
