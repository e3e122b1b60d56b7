"""Sampler drivers and the stochastic-EM M-step (experiments/utils_training.py of the reference).

On the hot path:
    predictive_average      utils_training.py:79-85, 160-166   Bayesian model average over stored samples (K7)
    MCEM_Q_maximizer        utils_training.py:339-359          Q = mean_s -U(.; W_s); one Adam step on the kernel
                                                                and likelihood hyper-parameters (K8 + Adam kernel)
Callers of the path (section 8f, kept to the reference's loop semantics):
    regression_train / classification_train   utils_training.py:11-172
    MCEM_sampler / MCEM                       utils_training.py:174-379

The reference loads UCI / MNIST by name inside these functions; data loading is out of scope here, so the
drivers take the datasets as arguments (`ds_train`, `ds_test`: iterables of (X, Y) minibatches, re-iterable).
The stored W samples are distinct tensors (the reference appends the live variables).
"""
import math

import numpy as np
import torch

from dgprf import _ffi
from dgprf.variable import out
from utils import cyclical_step_rate


# ------------------------------------------------------------------------------------------------
# predictive averaging (K7)
# ------------------------------------------------------------------------------------------------
def predictive_average(log_p, aux=None, aux_is_se=True, n_total_samples=None, return_lse=False):
    """log_p, aux: [S, N] (tensors or lists of [N] tensors).  Returns
    (mean_n(logsumexp_s log_p - log S), sqrt(mean(aux)) if aux_is_se else mean(aux))."""
    dev = _ffi.require_cuda()
    lp = torch.stack([_ffi.as_dev(t, dev) for t in log_p], 0) if isinstance(log_p, (list, tuple)) else _ffi.as_dev(log_p, dev)
    ax = None
    if aux is not None:
        ax = torch.stack([_ffi.as_dev(t, dev) for t in aux], 0) if isinstance(aux, (list, tuple)) else _ffi.as_dev(aux, dev)
        if ax.ndim == 1:                      # per-sample accuracies [S]: mean of the per-sample values
            ax = ax[:, None].expand(-1, lp.shape[1]).contiguous()
    S, N = lp.shape
    res = torch.empty(2, device=dev)
    lse = torch.empty(N, device=dev) if return_lse else None
    scratch = torch.empty(2 * ((N + 255) // 256) + 2, device=dev)
    _ffi.check(_ffi.lib().dgprf_predictive_reduce(lp.data_ptr(), _ffi.ptr(ax), S, N, N,
                                                  math.log(S if n_total_samples is None else n_total_samples),
                                                  int(bool(aux_is_se)), _ffi.ptr(lse), res.data_ptr(), scratch.data_ptr(),
                                                  _ffi.stream_ptr()))
    r = res.cpu()
    ret = (float(r[0]), float(r[1]) if aux is not None else None)
    return ret + (lse,) if return_lse else ret


# ------------------------------------------------------------------------------------------------
# stochastic-EM M-step (K8)
# ------------------------------------------------------------------------------------------------
class Adam:
    """keras.optimizers.Adam semantics (the optimizer of the notebooks, train_regression_EM_sin.ipynb
    cell 6) on the model's flat hyper-parameter buffer; the step itself is `dgprf_adam_step`."""

    def __init__(self, learning_rate=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.t = 0
        self.m = self.v = None

    def apply_flat(self, theta_h, grad_h):
        if self.m is None:
            self.m, self.v = torch.zeros_like(theta_h), torch.zeros_like(theta_h)
        self.t += 1
        _ffi.check(_ffi.lib().dgprf_adam_step(theta_h.data_ptr(), grad_h.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                              theta_h.numel(), self.lr, self.b1, self.b2, self.eps, self.t,
                                              _ffi.stream_ptr()))


def _sample_store(model, W_samples):
    """[S, w_len] flat store of S stored W sets (each a list of L tensors / Variables / arrays)."""
    e = model._engine
    store = torch.zeros(len(W_samples), e.layout.w_len, device=e.device, dtype=torch.float32)
    for s, Ws in enumerate(W_samples):
        for l, w in enumerate(Ws):
            off, ln = e.seg_w[f"W_{l}"][0], e.seg_w[f"W_{l}"][1]
            store[s, off:off + ln] = _ffi.as_dev(w, e.device).reshape(-1)
    return store


def em_q_and_grads(model, W_samples, X_batch, Y_batch, data_size, max_chunk=64):
    """Q = mean_s -U(X, Y; W_s) with W detached and no prior term, and d(-Q)/d hyper as a flat buffer laid out
    like the model's hyper buffer.  The stored samples ride the chain dimension of the kernels (shared z and
    hyper-parameters), `max_chunk` samples per launch."""
    e = model._engine
    store = W_samples if torch.is_tensor(W_samples) else _sample_store(model, W_samples)
    S = store.shape[0]
    B = int(np.shape(X_batch)[0])
    q_sum = 0.0
    g_sum = torch.zeros(e.layout.h_len, device=e.device)
    for s0 in range(0, S, max_chunk):
        chunk = store[s0:s0 + max_chunk].contiguous()
        m = e.model(w_base=chunk, w_cs=e.layout.w_len, n_chains=chunk.shape[0], h_cs=0)
        tot, _, gH = e.gradients(X_batch, Y_batch, data_size, hyper=True, prior_w=False, prior_h=False, m=m)
        q_sum = q_sum + tot.sum() / B
        g_sum += gH.sum(0)
    return q_sum / S, g_sum / S


def MCEM_Q_maximizer(model, data_size, optimizer):
    """maximizer(W_samples, X_batch, Y_batch): one optimizer step on [Omega_hyperparams, Likelihood_hyperparams]
    maximising Q (utils_training.py:339-359).  `optimizer` is an `Adam` instance of this module."""
    def maximizer(W_samples, X_batch, Y_batch):
        e = model._engine
        Q, g = em_q_and_grads(model, W_samples, X_batch, Y_batch, data_size)
        mask = torch.zeros(e.layout.h_len, device=e.device)
        for n in e.trainable_h:                                  # non-trainable hypers keep a zero gradient
            off, ln = e.seg_h[n][0], e.seg_h[n][1]
            mask[off:off + ln] = 1.0
        optimizer.apply_flat(e.theta_h[0], (g * mask).contiguous())
        n_s = W_samples.shape[0] if torch.is_tensor(W_samples) else len(W_samples)
        print("*" * 70)
        print(f"Q function is {float(Q)} averaged by {n_s} samples.")
        print("*" * 70, "\n")
        return out(Q)
    return maximizer


# ------------------------------------------------------------------------------------------------
# sampler drivers
# ------------------------------------------------------------------------------------------------
def _count_batches(ds):
    return sum(1 for _ in ds)


def _run_sampler(model, ds_train, ds_test, train_size, lr_0, momentum_decay, full_bayesian, precond_type, K_batches,
                 second_moment_centered, resample_in_cycle_head, total_epochs, start_sampling_epoch, epochs_per_cycle,
                 print_epoch_cycle, Y_std, task, collect_W, verbose):
    if precond_type != 'identity' and K_batches is None and second_moment_centered is None:
        raise ValueError("Args K_batches or second_moment_centered shouldn't be None!")
    iterations_per_epoch = _count_batches(ds_train)
    cycle_length = epochs_per_cycle * iterations_per_epoch
    log_p, aux, W_samples = [], [], []
    log_Y_std = math.log(Y_std)

    def evaluate(ds):
        if task == "reg":
            lp, se = model.eval_log_likelihood_and_se(ds)
            return lp - log_Y_std, se * Y_std ** 2
        return model.eval_log_likelihood(ds), model.eval_all_accuracy(ds)

    for epoch in range(total_epochs):
        model.precond_update(ds_train, train_size, K_batches=K_batches, full_bayesian=full_bayesian,
                             precond_type=precond_type, second_moment_centered=second_moment_centered)
        batch_index = 0
        for x_batch, y_batch in ds_train:
            batch_index += 1
            if epoch < start_sampling_epoch:                    # burn-in: fixed learning rate, zero temperature
                model.sgmcmc_update(x_batch, y_batch, train_size, lr=lr_0, momentum_decay=momentum_decay,
                                    full_bayesian=full_bayesian, resample_moments=False, temperature=0.)
                continue
            step_index = (epoch - start_sampling_epoch) * iterations_per_epoch + batch_index
            step_rate, is_end = cyclical_step_rate(step_index, cycle_length, schedule='cosine', min_value=0.)
            lr = lr_0 * (step_rate ** 2)
            is_new_cycle = resample_in_cycle_head and (step_index % cycle_length == 1)
            if lr > 0:
                model.sgmcmc_update(x_batch, y_batch, train_size, lr=float(lr), momentum_decay=momentum_decay,
                                    full_bayesian=full_bayesian, resample_moments=is_new_cycle, temperature=1.)
            if is_end:                                          # collect a posterior sample
                lp, ax = evaluate(ds_test)
                log_p.append(lp.as_subclass(torch.Tensor))
                aux.append(ax.as_subclass(torch.Tensor))
                if collect_W:
                    W_samples.append([w.tensor.clone() for w in model.W_mcmc])
                if verbose:
                    print('#' * 20, f'Sample No.{len(log_p)} at Epoch {epoch} ', f"lr = {lr}", '#' * 20)
        if verbose and (epoch + 1) % print_epoch_cycle == 0:
            tr, te = evaluate(ds_train), evaluate(ds_test)
            print(f"Epoch: {epoch}")
            print(f"Mean Log Likelihood -- train: {float(tr[0].mean())}, -- test: {float(te[0].mean())} ")
            if task == "reg":
                print(f"Root Mean Squared Error -- train: {float(tr[1].mean().sqrt())}, -- test: {float(te[1].mean().sqrt())} \n")
            else:
                print(f"Accuracy -- train: {float(tr[1])}, -- test: {float(te[1])} \n")
    log_p_t = torch.stack(log_p, 0)
    aux_t = torch.stack(aux, 0)
    metric = predictive_average(log_p_t, aux_t, aux_is_se=(task == "reg"))
    if verbose:
        print(f"Number of sampled models: {log_p_t.shape[0]} ")
        print(f"Test Log Likelihood of all sampled models: {metric[0]}")
        print(f"Test {'Root MSE' if task == 'reg' else 'mean accuracy'} of all sampled models: {metric[1]}")
    return W_samples, out(log_p_t), out(aux_t)


def regression_train(model, ds_train, ds_test, train_size, Y_std=1.0, lr_0=0.01, momentum_decay=0.9, full_bayesian=True,
                     precond_type='identity', K_batches=None, second_moment_centered=None,
                     resample_in_cycle_head=False, total_epochs=5000, start_sampling_epoch=2000, epochs_per_cycle=50,
                     print_epoch_cycle=100, verbose=True):
    """Cyclical SG-MCMC for regression (utils_training.py:11-91): burn-in at T=0 and lr_0, then lr = lr_0 rate^2
    with a cosine cycle, one posterior sample at every cycle end; returns (log_p [S, N_test], mse [S, N_test])."""
    _, log_p, mse = _run_sampler(model, ds_train, ds_test, train_size, lr_0, momentum_decay, full_bayesian, precond_type,
                                 K_batches, second_moment_centered, resample_in_cycle_head, total_epochs,
                                 start_sampling_epoch, epochs_per_cycle, print_epoch_cycle, Y_std, "reg", False, verbose)
    return log_p, mse


def classification_train(model, ds_train, ds_test, train_size, lr_0=0.01, momentum_decay=0.9, full_bayesian=True,
                         precond_type='identity', K_batches=None, second_moment_centered=None,
                         resample_in_cycle_head=False, total_epochs=5000, start_sampling_epoch=2000,
                         epochs_per_cycle=50, print_epoch_cycle=100, verbose=True):
    """utils_training.py:93-172; returns (log_p [S, N_test], acc [S])."""
    _, log_p, acc = _run_sampler(model, ds_train, ds_test, train_size, lr_0, momentum_decay, full_bayesian, precond_type,
                                 K_batches, second_moment_centered, resample_in_cycle_head, total_epochs,
                                 start_sampling_epoch, epochs_per_cycle, print_epoch_cycle, 1.0, "cls", False, verbose)
    return log_p, acc


def MCEM_sampler(model, ds_train, ds_test, train_size, Y_std=1.0, task="reg", lr_0=0.01, momentum_decay=0.9,
                 precond_type='identity', K_batches=None, second_moment_centered=None, resample_in_cycle_head=True,
                 start_sampling_epoch=2000, epochs_per_cycle=50, verbose=False):
    """E-step sampler factory (utils_training.py:174-337): sampler(num_samples) runs burn-in + num_samples cycles
    with the hyper-parameters fixed (full_bayesian=False) and returns (W_samples, log_p, mse | acc)."""
    def sampler(num_samples=100, print_epoch_cycle=100):
        return _run_sampler(model, ds_train, ds_test, train_size, lr_0, momentum_decay, False, precond_type, K_batches,
                            second_moment_centered, resample_in_cycle_head,
                            start_sampling_epoch + num_samples * epochs_per_cycle, start_sampling_epoch, epochs_per_cycle,
                            print_epoch_cycle, Y_std, task, True, verbose)
    return sampler


def MCEM(sampler_EM, maximizer, sampler_fixing_hyper, total_EM_steps, ds_train, num_samples_EM=100,
         num_samples_fixing_hyper=200, print_epoch_cycle_EM=100, print_epoch_cycle_fixing=100):
    """Monte-Carlo EM (utils_training.py:361-379): E = sample W, M = one maximizer step on the next minibatch."""
    em_step = 0
    while em_step < total_EM_steps:
        for x_batch, y_batch in ds_train:
            em_step += 1
            W_samples, _, _ = sampler_EM(num_samples=num_samples_EM, print_epoch_cycle=print_epoch_cycle_EM)
            maximizer(W_samples, x_batch, y_batch)
            if em_step == total_EM_steps:
                break
    _, log_p, mse_or_acc = sampler_fixing_hyper(num_samples=num_samples_fixing_hyper,
                                                print_epoch_cycle=print_epoch_cycle_fixing)
    return log_p, mse_or_acc
