"""Random-feature layers (layers/rf_layers.py:5-94 of the reference).

    Omega = exp(log_inv_length_scale)[:, None] * z + mean          (:36-38, :82-84)
    RBF:  Phi = amp/sqrt(M) * [cos(X Omega), sin(X Omega)]         (:42-44)   n_rf = 2*out_feature
    ARC:  Phi = sqrt(2) amp/sqrt(M) * relu(X Omega)                (:88-90)   n_rf = out_feature

Called on their own the layers run the stand-alone CUDA op ``dgprf_rf_features``; inside a
DGP_RF model the fused [RF -> GP] kernels are used instead and X Omega never reaches HBM.
"""
import torch

from dgprf import _ffi
from dgprf.variable import Variable, out
from kernels import RBFKernel, ARCKernel


class _RandomFeatureLayer:
    _kind = None
    _kernel_cls = None
    _features_per_omega = 1

    def __init__(self, kernel, out_feature, random_fixed=True, set_nonzero_mean=False, name=None):
        assert isinstance(kernel, self._kernel_cls), f"Input kernel is not {self._kernel_cls.kernel_type}!"
        self.name = name
        self.kernel = kernel
        self.in_feature = int(kernel.n_feature)
        self.out_feature = int(out_feature)
        self.n_rf = self._features_per_omega * self.out_feature
        self.random_fixed = random_fixed
        dev = kernel.log_amplitude.device
        # the fixed N(0,1) draw; also kept when random_fixed=False so the flag can be toggled
        self.z = torch.randn(self.in_feature, self.out_feature, device=dev, dtype=torch.float32)
        self.set_nonzero_mean = set_nonzero_mean
        if set_nonzero_mean:
            self.mean = Variable(torch.zeros(self.in_feature, 1), name="mean", device=dev)
        else:
            self.mean = torch.zeros(self.in_feature, 1, device=dev, dtype=torch.float32)

    @property
    def trainable_variables(self):
        vs = list(self.kernel.trainable_variables)
        if isinstance(self.mean, Variable) and self.mean.trainable:
            vs.append(self.mean)
        return vs

    def _log_inv_ls_vector(self):
        t = self.kernel.log_inv_length_scale.tensor
        return t if t.ndim == 1 else t.reshape(1).expand(self.in_feature).contiguous()

    def __call__(self, X):
        """X: [B, in_feature] -> [B, n_rf]."""
        dev = _ffi.require_cuda()
        X = _ffi.as_dev(X, dev)
        assert X.ndim == 2 and X.shape[1] == self.in_feature, "input width does not match the kernel"
        z = self.z if self.random_fixed else torch.randn_like(self.z)   # rf_layers.py:39-41
        mean = self.mean.tensor if isinstance(self.mean, Variable) else None
        Phi = torch.empty(X.shape[0], self.n_rf, device=dev, dtype=torch.float32)
        _ffi.check(_ffi.lib().dgprf_rf_features(
            self._kind, _ffi.ptr(X), X.shape[0], self.in_feature, _ffi.ptr(z.contiguous()),
            _ffi.ptr(self._log_inv_ls_vector()), _ffi.ptr(self.kernel.log_amplitude.tensor),
            _ffi.ptr(mean.reshape(-1).contiguous()) if mean is not None else None,
            self.out_feature, _ffi.ptr(Phi), _ffi.stream_ptr()))
        return out(Phi)

    def set_random_fixed(self, state):
        self.random_fixed = state


class RBFLayer(_RandomFeatureLayer):
    _kind = _ffi.KIND_RBF
    _kernel_cls = RBFKernel
    _features_per_omega = 2


class ARCLayer(_RandomFeatureLayer):
    _kind = _ffi.KIND_ARC
    _kernel_cls = ARCKernel
    _features_per_omega = 1
