"""RBFKernel -- hyper-parameter holder of k(x,y) = amp^2 exp(-|x-y|^2 / (2 ls^2)).
Mirrors kernels/RBF.py:5-53 of the reference (same constructor, attributes, properties)."""
from ._base import _StationaryHypers


class RBFKernel(_StationaryHypers):
    kernel_type = "RBF"

    def __init__(self, n_feature=1, amplitude=1., length_scale=None, trainable=True, is_ard=False, name=None):
        self._init_hypers(n_feature, amplitude, length_scale, trainable, is_ard, name)
