"""ClassificationDGP: DGP_RF with a softmax likelihood plus accuracy / log-likelihood evaluation
(models/classification_model.py:7-60)."""
import torch

from dgprf import _ffi
from dgprf.variable import out
from likelihoods import Softmax
from models.dgp import DGP_RF


class ClassificationDGP(DGP_RF):
    def __init__(self, d_in, d_out, n_hidden_layers=1, n_rf=30, n_gp=10, likelihood=None,
                 kernel_type_list=None, random_fixed=True, input_cat=False,
                 kernel_trainable=True, set_nonzero_mean=False, name=None):
        super().__init__(d_in, d_out, n_hidden_layers=n_hidden_layers, n_rf=n_rf, n_gp=n_gp,
                         likelihood=likelihood, kernel_type_list=kernel_type_list, input_cat=input_cat,
                         random_fixed=random_fixed, kernel_trainable=kernel_trainable,
                         set_nonzero_mean=set_nonzero_mean, name=name)

    def _correct_flags(self, X_batch, Y_batch):
        if self.BNN._fused_ok():
            _, correct, _ = self._engine.evaluate(X_batch, Y_batch)
            return correct[0]
        dev = _ffi.require_cuda()
        probs = self.likelihood.predict_full(self.BNN(X_batch)).as_subclass(torch.Tensor)
        labels = _ffi.as_dev(Y_batch, dev).reshape(-1)
        return (probs.argmax(-1).to(torch.float32) == labels).to(torch.float32)

    def eval_batch_accuracy(self, X_batch, Y_batch):
        """Accuracy of the current parameter sample on one batch (labels are floats [N, 1])."""
        return out(self._correct_flags(X_batch, Y_batch).mean())

    def eval_all_accuracy(self, ds_test):
        right, seen = None, 0
        for img_batch, label_batch in ds_test:
            c = self._correct_flags(img_batch, label_batch)
            right = c.sum() if right is None else right + c.sum()
            seen += c.numel()
        return out(right / float(seen))

    def eval_test_free_random(self, ds_test):
        """Accuracy with a fresh z drawn on every forward (classification_model.py:43-47)."""
        self.BNN.set_random_fixed(False)
        try:
            return self.eval_all_accuracy(ds_test)
        finally:
            self.BNN.set_random_fixed(True)

    def eval_log_likelihood(self, ds):
        lps = []
        for x_batch, y_batch in ds:
            lps.append(self.log_likelihood(x_batch, y_batch).as_subclass(torch.Tensor))
        return out(torch.cat(lps, dim=0))
