"""dgprf -- host side of the B200-native DGP-RF SG-MCMC hot path (ctypes over libdgprf.so)."""
from ._ffi import DgprfError, LIB_PATH  # noqa: F401
from .variable import Variable, DevTensor  # noqa: F401
