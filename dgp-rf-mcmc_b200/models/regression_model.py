"""RegressionDGP: DGP_RF with a Gaussian likelihood plus the evaluation forwards
(models/regression_model.py:6-50)."""
import torch

from dgprf import _ffi
from dgprf.variable import out
from likelihoods import Gaussian
from models.dgp import DGP_RF


class RegressionDGP(DGP_RF):
    def __init__(self, d_in, d_out, n_hidden_layers=1, n_rf=20, n_gp=2, likelihood=None,
                 kernel_type_list=None, kernel_trainable=True,
                 random_fixed=True, input_cat=False, set_nonzero_mean=False, name=None):
        super().__init__(d_in, d_out, n_hidden_layers=n_hidden_layers, n_rf=n_rf, n_gp=n_gp,
                         likelihood=likelihood, kernel_type_list=kernel_type_list,
                         kernel_trainable=kernel_trainable, random_fixed=random_fixed, input_cat=input_cat,
                         set_nonzero_mean=set_nonzero_mean, name=name)

    @staticmethod
    def _default_likelihood():
        return Gaussian()

    def feed_forward(self, ds):
        """Output mean of the last batch of ds (regression_model.py:16-22)."""
        last = None
        for x_batch, _ in ds:
            last = self.BNN(x_batch)
        return last

    def feed_forward_all_layers(self, X):
        """Outputs of every GP layer, layers applied one by one without input concatenation
        (regression_model.py:24-31)."""
        F = X
        outputs = []
        for l, layer in enumerate(self.BNN.layers):
            F = layer(F)
            if l % 2 == 1:
                outputs.append(F)
        return outputs

    def eval_log_likelihood_and_se(self, ds):
        """Per-point log p(y|f) [N] and squared error (mean over D_out) [N] over all batches of ds."""
        assert isinstance(self.likelihood, Gaussian), "The likelihood of the model is not Gaussian!"
        lps, ses = [], []
        for x_batch, y_batch in ds:
            if self.BNN._fused_ok():
                ll, se, _ = self._engine.evaluate(x_batch, y_batch)
                lps.append(ll[0]); ses.append(se[0])
            else:
                dev = _ffi.require_cuda()
                f = self.BNN(x_batch).as_subclass(torch.Tensor)
                y = _ffi.as_dev(y_batch, dev)
                lps.append(self.likelihood.log_prob(f, y).as_subclass(torch.Tensor))
                ses.append(((y - f) ** 2).mean(-1))
        return out(torch.cat(lps, dim=0)), out(torch.cat(ses, dim=0))

    def collect_W(self):
        """Compat shim for the demo drivers' `DemoRegressionDGP.collect_W()` (used by
        experiments/utils_training_demo.py:57,138 but defined nowhere in the reference)."""
        return {f"W_{i}": w.numpy() for i, w in enumerate(self.W_mcmc)}


DemoRegressionDGP = RegressionDGP
