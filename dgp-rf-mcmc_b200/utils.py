"""Layer chaining, densities and the cyclical step-rate schedule (utils.py of the reference).

    BNN_from_list            utils.py:5-25     sequential [RF, GP] * L
    BNN_from_list_input_cat  utils.py:28-44    concat([F, X]) before every RF layer but the first
    log_gaussian             utils.py:46-47
    cyclical_step_rate       utils.py:49-73    (host-side scalar arithmetic in float32)
"""
import numpy as np
import torch

from dgprf import _ffi
from dgprf.variable import out
from layers import RBFLayer, ARCLayer, GPLayer

_LOG_2PI = float(np.log(2.0 * np.pi))


class BNN_from_list:
    _input_cat = False

    def __init__(self, layer_list, name=None):
        self.name = name
        self.layers = layer_list
        self._engine_owner = None      # set by DGP_RF: enables the fused [RF -> GP] kernels

    def _fused_ok(self):
        return self._engine_owner is not None and all(
            layer.random_fixed for layer in self.layers if not isinstance(layer, GPLayer))

    def __call__(self, X, allow_gradient_from_W=True):
        if self._fused_ok():
            return out(self._engine_owner._engine.forward(X)[0])
        return self._call_layerwise(X, allow_gradient_from_W)

    def _call_layerwise(self, X, allow_gradient_from_W=True):
        F = X
        for layer in self.layers:
            F = layer(F, allow_gradient_from_W=allow_gradient_from_W) if isinstance(layer, GPLayer) else layer(F)
        return F

    def set_random_fixed(self, state):
        for layer in self.layers:
            if isinstance(layer, (RBFLayer, ARCLayer)):
                assert hasattr(layer, 'random_fixed'), "Layers cannot set random_fixed!"
                layer.set_random_fixed(state)

    @property
    def gp_layers(self):
        return (self.layers[2 * l + 1] for l in range(len(self.layers) // 2))

    @property
    def trainable_variables(self):
        vs = []
        for layer in self.layers:
            for v in layer.trainable_variables:
                if not any(v is u for u in vs):
                    vs.append(v)
        return vs


class BNN_from_list_input_cat(BNN_from_list):
    _input_cat = True

    def _call_layerwise(self, X, allow_gradient_from_W=True):
        dev = _ffi.require_cuda()
        X = _ffi.as_dev(X, dev)
        F = X
        last = len(self.layers) - 1
        for l, layer in enumerate(self.layers):
            if isinstance(layer, GPLayer):
                F = layer(F, allow_gradient_from_W=allow_gradient_from_W)
            else:
                if l != 0 and l != last:
                    F = torch.cat([F.as_subclass(torch.Tensor), X], dim=-1)   # tensor plumbing: [F, X]
                F = layer(F)
        return F


def log_gaussian(x, mean=0., var=1.):
    """Elementwise log N(x; mean, var).  Stand-alone helper (tensor plumbing); the model's
    likelihood / prior reductions use the fused CUDA kernels instead."""
    x = x.tensor if hasattr(x, "tensor") and not torch.is_tensor(x) else torch.as_tensor(x)
    var = torch.as_tensor(var, dtype=x.dtype, device=x.device)
    mean = torch.as_tensor(mean, dtype=x.dtype, device=x.device)
    return out(-0.5 * (_LOG_2PI + torch.log(var) + (x - mean) ** 2 / var))


def cyclical_step_rate(step_index, cycle_length, schedule='cosine', min_value=0.001):
    """Step-rate in (min_value, 1] for cyclical SG-MCMC; returns (rate, is_end_of_cycle).
    step_index counts from 1; float32 arithmetic like the reference."""
    step_index, cycle_length = int(step_index), int(cycle_length)
    if step_index <= 0:
        raise ValueError('Step index should be larger than zero!')
    f32 = np.float32
    frac = f32((step_index - 1) % cycle_length) / f32(cycle_length)
    lo, span = f32(min_value), f32(1.0 - min_value)
    if schedule == 'cosine':
        step_rate = lo + span * f32(0.5) * (np.cos(f32(np.pi) * frac, dtype=f32) + f32(1.0))
    elif schedule == 'glide':
        step_rate = lo + span * np.exp(-frac / (f32(1.0) - frac), dtype=f32)
    elif schedule == 'flat':
        step_rate = 1.0
    else:
        raise NotImplementedError
    return step_rate, (step_index % cycle_length) == 0
