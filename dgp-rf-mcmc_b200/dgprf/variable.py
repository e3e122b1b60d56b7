"""Parameter objects of the drop-in surface.

The reference keeps every parameter as a ``tf.Variable`` and hangs sampler state off it as
ad-hoc attributes (``param.moments``, ``param.M``; models/dgp.py:235-240).  ``Variable``
keeps those names but its storage is a *view into the model's flat parameter buffer*, so one
fused kernel can update every tensor of every chain (K5), and ``moments`` is a view into the
flat momentum buffer.
"""
from __future__ import annotations

import numpy as np
import torch


class DevTensor(torch.Tensor):
    """torch.Tensor whose ``.numpy()`` works from the device, as callers of the reference
    expect from TF eager tensors (e.g. ``line.numpy()[:, 0]`` in the demo notebooks)."""

    def numpy(self, *args, **kwargs):  # type: ignore[override]
        return self.detach().as_subclass(torch.Tensor).cpu().numpy(*args, **kwargs)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)


def out(t: torch.Tensor) -> DevTensor:
    return t.as_subclass(DevTensor)


class Variable:
    """Minimal tf.Variable look-alike over a torch tensor."""

    def __init__(self, value, trainable=True, name=None, device=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
                else torch.device("cpu")
        t = torch.as_tensor(np.asarray(value, dtype=np.float32) if not torch.is_tensor(value) else value)
        self._t = t.detach().to(device=device, dtype=torch.float32).clone()
        self.trainable = bool(trainable)
        self.name = name
        # set by the owning model when it adopts the variable into its flat buffers
        self._owner = None
        self._seg = None          # index of the segment in the owner's table
        self._mom = None          # view into the flat momentum buffer
        self._M = None

    # ---- storage -------------------------------------------------------------------------
    @property
    def tensor(self) -> torch.Tensor:
        return self._t

    def _rebind(self, view: torch.Tensor):
        """Move the storage into `view` (same shape), keeping the current value."""
        view.copy_(self._t.reshape(view.shape))
        self._t = view

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def dtype(self):
        return self._t.dtype

    @property
    def device(self):
        return self._t.device

    def numpy(self):
        return self._t.detach().cpu().numpy().copy()

    def value(self):
        return out(self._t.clone())

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __float__(self):
        return float(self._t)

    def __repr__(self):
        return f"<dgprf.Variable {self.name!r} shape={self.shape} trainable={self.trainable}>"

    # ---- tf.Variable mutators ------------------------------------------------------------
    def assign(self, value):
        if isinstance(value, Variable):
            value = value.tensor
        v = torch.as_tensor(value) if not torch.is_tensor(value) else value
        self._t.copy_(v.to(device=self._t.device, dtype=torch.float32).reshape(self._t.shape))
        return self

    def assign_add(self, delta):
        v = torch.as_tensor(delta) if not torch.is_tensor(delta) else delta
        self._t.add_(v.to(device=self._t.device, dtype=torch.float32).reshape(self._t.shape))
        return self

    # ---- sampler state (models/dgp.py:235-240) ------------------------------------------
    @property
    def moments(self):
        if self._mom is None:
            raise AttributeError("moments")
        return out(self._mom)

    @moments.setter
    def moments(self, value):
        if self._owner is None:
            raise AttributeError("variable is not owned by a model; call precond_update first")
        if self._mom is None:
            self._mom = self._owner._moment_view(self)
        v = torch.as_tensor(value) if not torch.is_tensor(value) else value
        self._mom.copy_(v.to(device=self._mom.device, dtype=torch.float32).reshape(self._mom.shape))

    @property
    def M(self):
        if self._M is None:
            raise AttributeError("M")
        return self._M

    @M.setter
    def M(self, value):
        self._M = float(value)
        if self._owner is not None:
            self._owner._set_mass(self, self._M)
