"""Sampler drivers and the stochastic-EM M-step (experiments/utils_training.py of the reference).

On the hot path:
    predictive_average      utils_training.py:79-85, 160-166   Bayesian model average over stored samples (K7)
    MCEM_Q_maximizer        utils_training.py:339-359          Q = mean_s -U(.; W_s); one Adam step on the kernel
                                                                and likelihood hyper-parameters (K8 + Adam kernel)
Callers of the path (section 8f), same names and positional signatures as the reference:
    regression_train / classification_train                    utils_training.py:11-172
    MCEM_sampler_UCI / MCEM_sampler_classification / MCEM       utils_training.py:174-379
    MCEM_windows / MCEM_increasing_windows                     utils_training.py:381-473
plus what a device-resident loop needs: `SampleWindow` (ring buffer of stored samples), `EpochGraph` (one CUDA-graph
launch per epoch), `MCEM_sampler` (the sampler factory over explicit datasets).

The by-name drivers load UCI CSVs / an MNIST .npz through experiments/utils_dataset.py; `data=(ds_train, ds_test,
train_size, Y_std)` hands in-memory sets to the same loops (`ds_*`: re-iterable sources of (X, Y) minibatches).
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
# stored posterior samples: a device ring buffer
# ------------------------------------------------------------------------------------------------
class SampleWindow:
    """The last `capacity` posterior samples, on the device: flat W sets [cap, w_len], their per-test-point log-densities
    [cap, N_test] and squared errors [cap, N_test] (or accuracies [cap]).  `push` overwrites the oldest slot, so the
    sliding window of MCEM_windows (utils_training.py:393-405: concat, then drop row 0) costs one row copy per E-step
    instead of re-concatenating the window; the Bayesian model average (K7) and the M-step read the slots in place --
    both are symmetric in the sample order."""

    def __init__(self, model, capacity):
        self.model, self.capacity = model, int(capacity)
        self.W = torch.zeros(self.capacity, model._engine.layout.w_len, device=model._engine.device)
        self.log_p = self.aux = None
        self.count = 0          # samples pushed so far
        self.head = 0           # next slot to write

    def __len__(self):
        return min(self.count, self.capacity)

    def push(self, W_flat, log_p, aux):
        """W_flat [w_len] (or a list of L tensors), log_p [N], aux [N] | scalar."""
        if not torch.is_tensor(W_flat):
            W_flat = _sample_store(self.model, [W_flat])[0]
        log_p = log_p.as_subclass(torch.Tensor).reshape(-1)
        aux = aux.as_subclass(torch.Tensor).reshape(-1)
        if self.log_p is None:
            self.log_p = torch.zeros(self.capacity, log_p.numel(), device=self.W.device)
            self.aux = torch.zeros(self.capacity, aux.numel(), device=self.W.device)
        self.W[self.head].copy_(W_flat.reshape(-1))
        self.log_p[self.head].copy_(log_p)
        self.aux[self.head].copy_(aux)
        self.head = (self.head + 1) % self.capacity
        self.count += 1

    def stored(self):
        """(W [S, w_len], log_p [S, N], aux [S, N | 1]) views of the S live slots."""
        S = len(self)
        return self.W[:S], self.log_p[:S], self.aux[:S]

    def average(self, aux_is_se):
        _, lp, ax = self.stored()
        return predictive_average(lp, ax if ax.shape[1] > 1 else ax[:, 0], aux_is_se=aux_is_se)

    def pick(self, i):
        """Sample i of the window as a one-row flat store (the M-step input of MCEM_windows, :422-423)."""
        return self.W[i:i + 1]


# ------------------------------------------------------------------------------------------------
# one epoch of sampling steps as ONE CUDA-graph launch
# ------------------------------------------------------------------------------------------------
class EpochGraph:
    """The minibatch steps of one epoch (utils_training.py:45-61) captured as a CUDA graph.

    Learning rate, temperature, resample flag and the minibatch pointers of every step are baked into the captured
    kernel nodes; what changes between replays lives in device memory: the dataset's permuted buffers (re-gathered by
    `DeviceDataset.reshuffle`, same addresses) and the Philox step base (`dgprf_sgmcmc_step_graph`).  A replay is one
    launch for the whole epoch, bit-identical to the same steps issued one by one (tests/test_graph_epochs_gpu.py)."""

    def __init__(self, model, ds, data_size, lrs, momentum_decay, temperature, resample_first, full_bayesian):
        e = model._engine
        self.model, self.n = model, len(lrs)
        assert self.n == len(ds) and all(lr > 0 for lr in lrs)
        if not hasattr(model, "_step_dev"):
            model._step_dev = torch.zeros(1, dtype=torch.int64, device=e.device)
        model._step_dev.fill_(model._step)
        batches = [ds.batch(i) for i in range(self.n)]
        e.workspace(e.model(), batches[0][0].shape[0], _ffi.MODE_HYPER if full_bayesian else _ffi.MODE_TRAIN)  # allocate before capture
        self.graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            for i, (xb, yb) in enumerate(batches):
                e.step(xb, yb, float(data_size), float(lrs[i]), float(momentum_decay), float(temperature),
                       bool(resample_first and i == 0), bool(full_bayesian), model._seed, i + 1, step_base=model._step_dev)

    def replay(self):
        m = self.model
        m._step_dev.fill_(m._step)          # host step counter -> device base: step i of the epoch draws noise (step + i)
        self.graph.replay()
        m._step += self.n


# ------------------------------------------------------------------------------------------------
# sampler drivers
# ------------------------------------------------------------------------------------------------
def _count_batches(ds):
    return len(ds) if hasattr(ds, "__len__") else sum(1 for _ in ds)


def _run_sampler(model, ds_train, ds_test, train_size, lr_0, momentum_decay, full_bayesian, precond_type, K_batches,
                 second_moment_centered, resample_in_cycle_head, total_epochs, start_sampling_epoch, epochs_per_cycle,
                 print_epoch_cycle, Y_std, task, collect_W, verbose, graph=None, on_sample=None):
    """The loop of utils_training.py:41-77 / :121-152 / :206-236.  graph=True replays one captured CUDA graph per epoch
    (needs a DeviceDataset and the identity preconditioner: masses are baked into the graph); graph=None (the default) does so
    whenever that is possible and issues the steps one by one otherwise; graph=False never captures.  Bit-identical either way."""
    if precond_type != 'identity' and K_batches is None and second_moment_centered is None:
        raise ValueError("Args K_batches or second_moment_centered shouldn't be None!")
    iterations_per_epoch = _count_batches(ds_train)
    cycle_length = epochs_per_cycle * iterations_per_epoch
    log_p, aux, W_samples = [], [], []
    log_Y_std = math.log(Y_std)
    use_graph = graph is None or bool(graph)            # None: one graph launch per epoch whenever the set-up allows it
    if use_graph:
        why = None
        if precond_type != 'identity':
            why = "the per-tensor masses of a non-identity preconditioner are baked into the captured kernels"
        elif not hasattr(ds_train, "reshuffle"):
            why = "the minibatches must sit at fixed device addresses (experiments.utils_dataset.DeviceDataset)"
        elif not model.BNN._fused_ok():
            why = "random_fixed=False redraws z on the host before every step"
        elif not (ds_train.drop_remainder or ds_train.N % ds_train.batch_size == 0):
            why = "all minibatches of an epoch must have the same size (drop_remainder=True)"
        if why is not None:
            if graph is None:
                use_graph = False                           # default: fall back to one launch sequence per step
            else:
                raise ValueError(f"graph=True is not possible here: {why}")
    graphs = {}

    def evaluate(ds):
        if task == "reg":
            lp, se = model.eval_log_likelihood_and_se(ds)
            return lp - log_Y_std, se * Y_std ** 2
        return model.eval_log_likelihood(ds), model.eval_all_accuracy(ds)

    def schedule(epoch):
        """(lr, temperature, resample flag, is_end) of every step of `epoch`."""
        rows = []
        for batch_index in range(1, iterations_per_epoch + 1):
            if epoch < start_sampling_epoch:                    # burn-in: fixed learning rate, zero temperature
                rows.append((lr_0, 0., False, False))
                continue
            step_index = (epoch - start_sampling_epoch) * iterations_per_epoch + batch_index
            step_rate, is_end = cyclical_step_rate(step_index, cycle_length, schedule='cosine', min_value=0.)
            lr = lr_0 * (step_rate ** 2)
            is_new_cycle = bool(resample_in_cycle_head) and (step_index % cycle_length == 1)
            rows.append((float(lr), 1., is_new_cycle, is_end))
        return rows

    for epoch in range(total_epochs):
        model.precond_update(ds_train, train_size, K_batches=K_batches, full_bayesian=full_bayesian,
                             precond_type=precond_type, second_moment_centered=second_moment_centered)
        rows = schedule(epoch)
        sampled_lr = None
        if use_graph and all(r[0] > 0 and not r[2] for r in rows[1:]) and not any(r[3] for r in rows[:-1]):
            key = -1 if epoch < start_sampling_epoch else (epoch - start_sampling_epoch) % epochs_per_cycle
            ds_train.reshuffle()
            if key not in graphs:
                graphs[key] = EpochGraph(model, ds_train, train_size, [r[0] for r in rows], momentum_decay, rows[0][1],
                                         rows[0][2], full_bayesian)
            graphs[key].replay()
            if rows[-1][3]:
                sampled_lr = rows[-1][0]
        else:
            for (x_batch, y_batch), (lr, T, resample, is_end) in zip(ds_train, rows):
                if lr > 0:
                    model.sgmcmc_update(x_batch, y_batch, train_size, lr=lr, momentum_decay=momentum_decay,
                                        full_bayesian=full_bayesian, resample_moments=resample, temperature=T)
                if is_end:
                    sampled_lr = lr
        if sampled_lr is not None:                              # a cycle ended with this epoch: collect a posterior sample
            lp, ax = evaluate(ds_test)
            log_p.append(lp.as_subclass(torch.Tensor))
            aux.append(ax.as_subclass(torch.Tensor))
            if collect_W:
                W_samples.append([w.tensor.clone() for w in model.W_mcmc])
            if on_sample is not None:
                on_sample(model, log_p[-1], aux[-1])
            if verbose:
                print('#' * 20, f'Sample No.{len(log_p)} at Epoch {epoch} ', f"lr = {sampled_lr}", '#' * 20)
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


def _uci_data(dataset_name, batch_size, data_dir, verbose=True):
    """Data handling of utils_training.py:19-32: Y_std, remainder-dropped training size, whole-set fallback."""
    from experiments.utils_dataset import download_UCI_data_info, load_UCI_dataset
    _, _, _, _, _, _, Y_std = download_UCI_data_info(dataset_name, data_path=data_dir)
    ds_train, ds_test, train_shape, _ = load_UCI_dataset(dataset_name, batch_size=batch_size, data_dir=data_dir, verbose=verbose)
    train_size = train_shape[0]
    if train_size - train_size % batch_size == 0:               # batch size > train size
        print("Training size is 0 after remainder dropping! Using the whole data as one batch! ")
        ds_train, ds_test, _, _ = load_UCI_dataset(dataset_name, batch_size=batch_size, data_dir=data_dir,
                                                   drop_train_remainder=False, verbose=False)
    if verbose:
        print(f"Training size is {len(ds_train) * min(batch_size, train_size)} after remainder dropping. ")
    return ds_train, ds_test, train_size, float(Y_std[0])


def _mnist_data(dataset_name, batch_size, data_dir):
    from experiments.utils_dataset import load_tf_dataset, normalize_MNIST
    ds_train, ds_test, train_full_size, _ = load_tf_dataset(dataset_name, transform_fn=normalize_MNIST, batch_size=batch_size,
                                                           data_dir=data_dir)
    return ds_train, ds_test, train_full_size, 1.0


def _resolve(data, dataset_name, batch_size, data_dir, task, verbose=True):
    """`data=(ds_train, ds_test, train_size[, Y_std])` (in-memory / synthetic sets) or the reference's by-name loading."""
    if data is not None:
        ds_train, ds_test, train_size = data[0], data[1], data[2]
        return ds_train, ds_test, train_size, (float(data[3]) if len(data) > 3 else 1.0)
    if task == "reg":
        return _uci_data(dataset_name, batch_size, data_dir, verbose)
    return _mnist_data(dataset_name, batch_size, data_dir)


def regression_train(model, dataset_name='boston', batch_size=200, data_dir='./data/',
                     lr_0=0.01, momentum_decay=0.9, full_bayesian=True,
                     precond_type='identity', K_batches=None, second_moment_centered=None,
                     resample_in_cycle_head=False,
                     total_epochs=5000, start_sampling_epoch=2000, epochs_per_cycle=50,
                     print_epoch_cycle=100, *, data=None, verbose=True, graph=None):
    """Cyclical SG-MCMC for regression (utils_training.py:11-91, same positional signature): burn-in at T=0 and lr_0,
    then lr = lr_0 rate^2 with a cosine cycle, one posterior sample at every cycle end; returns
    (log_p [S, N_test], mse [S, N_test]).  Extensions (keyword-only): `data=(ds_train, ds_test, train_size, Y_std)`
    replaces the by-name UCI loading; `graph=True` launches every epoch as one CUDA graph."""
    ds_train, ds_test, train_size, Y_std = _resolve(data, dataset_name, batch_size, data_dir, "reg", verbose)
    _, log_p, mse = _run_sampler(model, ds_train, ds_test, train_size, lr_0, momentum_decay, full_bayesian, precond_type,
                                 K_batches, second_moment_centered, resample_in_cycle_head, total_epochs,
                                 start_sampling_epoch, epochs_per_cycle, print_epoch_cycle, Y_std, "reg", False, verbose, graph)
    return log_p, mse


def classification_train(model, dataset_name='mnist', batch_size=200, data_dir='./tensorflow_datasets/',
                         lr_0=0.01, momentum_decay=0.9, full_bayesian=True,
                         precond_type='identity', K_batches=None, second_moment_centered=None,
                         resample_in_cycle_head=False,
                         total_epochs=5000, start_sampling_epoch=2000, epochs_per_cycle=50,
                         print_epoch_cycle=100, *, data=None, verbose=True, graph=None):
    """utils_training.py:93-172; returns (log_p [S, N_test], acc [S])."""
    ds_train, ds_test, train_size, _ = _resolve(data, dataset_name, batch_size, data_dir, "cls", verbose)
    _, log_p, acc = _run_sampler(model, ds_train, ds_test, train_size, lr_0, momentum_decay, full_bayesian, precond_type,
                                 K_batches, second_moment_centered, resample_in_cycle_head, total_epochs,
                                 start_sampling_epoch, epochs_per_cycle, print_epoch_cycle, 1.0, "cls", False, verbose, graph)
    return log_p, acc


def MCEM_sampler(model, ds_train, ds_test, train_size, Y_std=1.0, task="reg", lr_0=0.01, momentum_decay=0.9,
                 precond_type='identity', K_batches=None, second_moment_centered=None, resample_in_cycle_head=True,
                 start_sampling_epoch=2000, epochs_per_cycle=50, verbose=False, graph=None):
    """E-step sampler factory (utils_training.py:174-337): sampler(num_samples) runs burn-in + num_samples cycles
    with the hyper-parameters fixed (full_bayesian=False) and returns (W_samples, log_p, mse | acc)."""
    def sampler(num_samples=100, print_epoch_cycle=100):
        return _run_sampler(model, ds_train, ds_test, train_size, lr_0, momentum_decay, False, precond_type, K_batches,
                            second_moment_centered, resample_in_cycle_head,
                            start_sampling_epoch + num_samples * epochs_per_cycle, start_sampling_epoch, epochs_per_cycle,
                            print_epoch_cycle, Y_std, task, True, verbose, graph)
    return sampler


def MCEM_sampler_UCI(model, dataset_name='boston', batch_size=200, data_dir='./data/',
                     lr_0=0.01, momentum_decay=0.9,
                     precond_type='identity', K_batches=None, second_moment_centered=None,
                     resample_in_cycle_head=True, start_sampling_epoch=2000, epochs_per_cycle=50, *, data=None, graph=None):
    """utils_training.py:174-256 (same signature): the E-step sampler over a UCI regression set."""
    ds_train, ds_test, train_size, Y_std = _resolve(data, dataset_name, batch_size, data_dir, "reg")
    return MCEM_sampler(model, ds_train, ds_test, train_size, Y_std, "reg", lr_0, momentum_decay, precond_type, K_batches,
                        second_moment_centered, resample_in_cycle_head, start_sampling_epoch, epochs_per_cycle, True, graph)


def MCEM_sampler_classification(model, dataset_name='mnist', batch_size=200, data_dir='./tensorflow_datasets/',
                                lr_0=0.01, momentum_decay=0.9,
                                precond_type='identity', K_batches=None, second_moment_centered=None,
                                resample_in_cycle_head=True, start_sampling_epoch=2000, epochs_per_cycle=50, *, data=None,
                                graph=None):
    """utils_training.py:258-337 (same signature): the E-step sampler over a tfds-style classification set."""
    ds_train, ds_test, train_size, _ = _resolve(data, dataset_name, batch_size, data_dir, "cls")
    return MCEM_sampler(model, ds_train, ds_test, train_size, 1.0, "cls", lr_0, momentum_decay, precond_type, K_batches,
                        second_moment_centered, resample_in_cycle_head, start_sampling_epoch, epochs_per_cycle, True, graph)


def _repeat(ds):
    if hasattr(ds, "repeat"):
        return ds.repeat()
    def gen():
        while True:
            yield from ds
    return gen()


def MCEM(sampler_EM, maximizer, sampler_fixing_hyper, total_EM_steps, ds_train, num_samples_EM=100,
         num_samples_fixing_hyper=200, print_epoch_cycle_EM=100, print_epoch_cycle_fixing=100):
    """Monte-Carlo EM (utils_training.py:361-379): E = sample W, M = one maximizer step on the next minibatch."""
    em_step = 0
    for x_batch, y_batch in _repeat(ds_train):
        em_step += 1
        print("#" * 15, f"EM step {em_step} of total {total_EM_steps} steps. E Step: ", "#" * 15)
        W_samples, _, _ = sampler_EM(num_samples=num_samples_EM, print_epoch_cycle=print_epoch_cycle_EM)
        print("#" * 15, f"EM step {em_step} of total {total_EM_steps} steps, M Step: ", "#" * 15)
        maximizer(W_samples, x_batch, y_batch)
        if em_step == total_EM_steps:
            break
    print("#" * 15, f"After {total_EM_steps} EM steps, fixing hyperparams and sample from posterior.", "#" * 15)
    _, log_p, mse_or_acc = sampler_fixing_hyper(num_samples=num_samples_fixing_hyper,
                                                print_epoch_cycle=print_epoch_cycle_fixing)
    return log_p, mse_or_acc


def _model_of(sampler_EM, maximizer):
    for fn in (maximizer, sampler_EM):
        for cell in (getattr(fn, "__closure__", None) or ()):
            try:
                v = cell.cell_contents
            except ValueError:
                continue
            if hasattr(v, "_engine") and hasattr(v, "W_mcmc"):
                return v
    raise ValueError("cannot find the model behind sampler_EM / maximizer")


def _mcem_windows(sampler_EM, maximizer, sampler_fixing_hyper, total_EM_steps, ds_train, num_samples_fixing_hyper,
                  window_size, print_epoch_cycle_EM, print_epoch_cycle_fixing, rng=None):
    model = _model_of(sampler_EM, maximizer)
    window = SampleWindow(model, window_size)
    rng = np.random if rng is None else rng
    em_step = 0
    for x_batch, y_batch in _repeat(ds_train):
        em_step += 1
        # E step: ONE new posterior sample joins the window; beyond window_size the oldest one leaves
        print("#" * 15, f"EM step {em_step} of total {total_EM_steps} steps. E Step: ", "#" * 15)
        W_samples, log_p, mse_or_acc = sampler_EM(num_samples=1, print_epoch_cycle=print_epoch_cycle_EM)
        window.push(W_samples[-1], log_p[-1], mse_or_acc[-1])
        is_acc = window.aux.shape[1] == 1                        # accuracies are one scalar per sample, squared errors [N]
        predict_log_p, predict_rmse_or_acc = window.average(aux_is_se=not is_acc)
        print("*" * 20, " End of E step ", "*" * 20)
        print(f"Number of all sampled models in window: {len(window)} ")
        print(f"Test Log Likelihood of all models in window: {predict_log_p}")
        print(f"Test {'Mean Acc' if is_acc else 'Root MSE'} of all models in window: {predict_rmse_or_acc}\n")
        # M step on ONE randomly chosen sample of the window
        print("#" * 15, f"EM step {em_step} of total {total_EM_steps} steps, M Step: ", "#" * 15)
        i = int(rng.randint(len(window)))
        maximizer(window.pick(i), x_batch, y_batch)
        if em_step == total_EM_steps:
            break
    print("#" * 15, f"After {total_EM_steps} EM steps, fixing hyperparams and sample from posterior.", "#" * 15)
    return window


def MCEM_windows(sampler_EM, maximizer, sampler_fixing_hyper, total_EM_steps, ds_train,
                 num_samples_fixing_hyper=200, window_size=300,
                 print_epoch_cycle_EM=100, print_epoch_cycle_fixing=100):
    """Moving-window MCEM (utils_training.py:381-429): every EM step draws one new W sample into a sliding window of at
    most `window_size` samples (reported as a Bayesian model average), and the M-step uses one random window sample."""
    _mcem_windows(sampler_EM, maximizer, sampler_fixing_hyper, total_EM_steps, ds_train, num_samples_fixing_hyper, window_size,
                  print_epoch_cycle_EM, print_epoch_cycle_fixing)
    _, log_p, mse_or_acc = sampler_fixing_hyper(num_samples=num_samples_fixing_hyper,
                                                print_epoch_cycle=print_epoch_cycle_fixing)
    return log_p, mse_or_acc


def MCEM_increasing_windows(sampler_EM, maximizer, sampler_fixing_hyper, total_EM_steps, ds_train,
                            num_samples_fixing_hyper=200, window_size=300,
                            print_epoch_cycle_EM=100, print_epoch_cycle_fixing=100):
    """utils_training.py:431-473: the regression-only twin of MCEM_windows (the window grows to window_size, then slides)."""
    return MCEM_windows(sampler_EM, maximizer, sampler_fixing_hyper, total_EM_steps, ds_train, num_samples_fixing_hyper,
                        window_size, print_epoch_cycle_EM, print_epoch_cycle_fixing)
