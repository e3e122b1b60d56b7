"""ARCKernel -- arc-cosine kernel hyper-parameters; only degree 1 exists in the reference
(kernels/arc_cosine.py:13-16 raises NotImplementedError otherwise)."""
from ._base import _StationaryHypers


class ARCKernel(_StationaryHypers):
    kernel_type = "ARC"

    def __init__(self, n_feature=1, amplitude=1., length_scale=None, trainable=True, is_ard=False, degree=1,
                 name=None):
        if degree != 1:
            raise NotImplementedError
        self.degree = degree
        self._init_hypers(n_feature, amplitude, length_scale, trainable, is_ard, name)
