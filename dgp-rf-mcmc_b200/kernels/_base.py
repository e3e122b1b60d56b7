"""Shared constructor logic of the kernel hyper-parameter holders (kernels/RBF.py:6-41 and
kernels/arc_cosine.py:6-44 are identical apart from the degree check)."""
import numpy as np

from dgprf.variable import Variable, out
import torch


class _StationaryHypers:
    kernel_type = None

    def _init_hypers(self, n_feature, amplitude, length_scale, trainable, is_ard, name):
        self.name = name
        self.n_feature = int(n_feature)
        ls = np.sqrt(np.float32(self.n_feature)) if length_scale is None else np.asarray(length_scale, dtype=np.float32)
        ls = np.asarray(ls, dtype=np.float32)
        if ls.ndim >= 2:
            raise ValueError("The length scale of RBF dim error!")
        inv_ls = np.float32(1.0) / ls
        if inv_ls.ndim == 0 and is_ard:
            inv_ls = inv_ls * np.ones(self.n_feature, dtype=np.float32)
            self.is_ard = is_ard
        else:
            if inv_ls.ndim == 1 and inv_ls.size != self.n_feature:
                raise ValueError("The size of length scale and features do not match!")
            self.is_ard = inv_ls.ndim == 1
            if self.is_ard != is_ard:
                print(f"Arg is_ard={is_ard} does not match the length_scale!")
                print(f"Already set is_ard={self.is_ard}")
        self.log_amplitude = Variable(np.log(np.float32(amplitude)), trainable=trainable, name="log_amplitude")
        self.log_inv_length_scale = Variable(np.log(inv_ls.astype(np.float32)), trainable=trainable,
                                             name="log_inv_length_scale")

    @property
    def amplitude(self):
        return out(torch.exp(self.log_amplitude.tensor))

    @property
    def inv_length_scale(self):
        return out(torch.exp(self.log_inv_length_scale.tensor))

    @property
    def length_scale(self):
        return out(1.0 / torch.exp(self.log_inv_length_scale.tensor))

    @property
    def trainable_variables(self):
        return [v for v in (self.log_amplitude, self.log_inv_length_scale) if v.trainable]
