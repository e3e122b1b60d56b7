"""Softmax likelihood: -sparse softmax cross-entropy with labels int(Y[:, 0])
(likelihoods/softmax.py:4-22)."""
import torch

from dgprf import _ffi
from dgprf.variable import out


class Softmax:
    trainable_variables = []

    def log_prob(self, F, Y):
        """F: [B, C] logits, Y: [B, 1] float labels -> [B]."""
        dev = _ffi.require_cuda()
        F, Y = _ffi.as_dev(F, dev), _ffi.as_dev(Y, dev)
        assert F.ndim == 2 and Y.ndim == 2 and Y.shape[0] == F.shape[0]
        y0 = Y[:, 0].contiguous()
        o = torch.empty(F.shape[0], device=dev, dtype=torch.float32)
        _ffi.check(_ffi.lib().dgprf_softmax_log_prob(_ffi.ptr(F), _ffi.ptr(y0), F.shape[0], F.shape[1],
                                                     _ffi.ptr(o), None, _ffi.stream_ptr()))
        return out(o)

    def predict_full(self, F):
        dev = _ffi.require_cuda()
        F = _ffi.as_dev(F, dev)
        p = torch.empty_like(F)
        _ffi.check(_ffi.lib().dgprf_softmax_log_prob(_ffi.ptr(F), None, F.shape[0], F.shape[1], None,
                                                     _ffi.ptr(p), _ffi.stream_ptr()))
        return out(p)
