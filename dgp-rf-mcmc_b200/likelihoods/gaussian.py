"""Gaussian likelihood with a learnable log-variance (likelihoods/gaussian.py:6-25)."""
import numpy as np
import torch

from dgprf import _ffi
from dgprf.variable import Variable, out


class Gaussian:
    def __init__(self, variance=0.1, trainable=True):
        self.lik_log_var = Variable(np.log(np.float32(variance)), trainable=trainable, name="lik_log_var")

    @property
    def trainable_variables(self):
        return [self.lik_log_var] if self.lik_log_var.trainable else []

    @property
    def variance(self):
        return out(torch.exp(self.lik_log_var.tensor))

    def log_prob(self, F, Y):
        """sum_D log N(Y; F, variance) -> [B]."""
        dev = _ffi.require_cuda()
        F, Y = _ffi.as_dev(F, dev), _ffi.as_dev(Y, dev)
        assert F.shape == Y.shape and F.ndim == 2
        o = torch.empty(F.shape[0], device=dev, dtype=torch.float32)
        _ffi.check(_ffi.lib().dgprf_gaussian_log_prob(_ffi.ptr(F), _ffi.ptr(Y), _ffi.ptr(self.lik_log_var.tensor),
                                                      F.shape[0], F.shape[1], _ffi.ptr(o), _ffi.stream_ptr()))
        return out(o)
