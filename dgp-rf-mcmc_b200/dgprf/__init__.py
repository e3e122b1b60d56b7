"""dgprf -- host side of the B200-native DGP-RF SG-MCMC hot path (ctypes over libdgprf.so)."""
from ._ffi import DgprfError, LIB_PATH  # noqa: F401
from .variable import Variable, DevTensor  # noqa: F401

from . import _ffi as _f

_PRECISIONS = {"fp32": _f.PREC_FP32, "tf32": _f.PREC_TF32}
_default_precision = _f.PREC_FP32


def set_default_precision(name: str) -> None:
    """'fp32': SIMT FFMA kernels (parity mode, rtol 1e-4).  'tf32': tcgen05 kind::tf32 tensor-core
    kernels where the layer shape allows (looser stated bound).  Applies to models
    built afterwards; `model.set_precision(name)` switches an existing model."""
    global _default_precision
    _default_precision = _PRECISIONS[name]


def default_precision() -> int:
    return _default_precision
