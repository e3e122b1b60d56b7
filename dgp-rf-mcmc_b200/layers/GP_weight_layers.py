"""GP weight layer F = Phi W, W ~ N(0, 1) at construction (layers/GP_weight_layers.py:4-20)."""
import torch

from dgprf import _ffi
from dgprf.variable import Variable, out


class GPLayer:
    def __init__(self, in_feature, out_feature, name=None):
        self.name = name
        self.in_feature = int(in_feature)
        self.out_feature = int(out_feature)
        self.W = Variable(torch.randn(self.in_feature, self.out_feature), name="GP_layer_W")

    @property
    def trainable_variables(self):
        return [self.W]

    def __call__(self, X, allow_gradient_from_W=True):
        # allow_gradient_from_W only selects stop_gradient in the reference (:13-15); gradients
        # here come from the explicit backward kernels, so both branches compute the same product.
        dev = _ffi.require_cuda()
        X = _ffi.as_dev(X, dev)
        assert X.ndim == 2 and X.shape[1] == self.in_feature
        o = torch.empty(X.shape[0], self.out_feature, device=dev, dtype=torch.float32)
        _ffi.check(_ffi.lib().dgprf_gp_matmul(_ffi.ptr(X), _ffi.ptr(self.W.tensor.contiguous()), X.shape[0],
                                              self.in_feature, self.out_feature, _ffi.ptr(o), _ffi.stream_ptr()))
        return out(o)

    def assign_W(self, W_value):
        self.W.assign(W_value)
