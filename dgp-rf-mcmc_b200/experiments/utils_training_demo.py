"""Demo twins of the sampler drivers (experiments/utils_training_demo.py of the reference): the toy notebooks' variants
that also return the sampled regression lines of every layer and the stored W samples for plotting.

    regression_train_demo      utils_training_demo.py:10-85
    MCEM_sampler_demo          utils_training_demo.py:87-169
    MCEM_Q_maximizer_demo      utils_training_demo.py:171-191
    MCEM_demo                  utils_training_demo.py:193-213
    MCEM_windows_demo          utils_training_demo.py:215-259

Same positional signatures as the reference.  `model_demo` is a RegressionDGP (the reference's `DemoRegressionDGP`, which its
repository uses but never defines: RegressionDGP + collect_W() -> {'W_i': ndarray [F_i, g_i]}).  The loops are the ones of
experiments/utils_training.py (`_run_sampler`), with a per-sample hook collecting `feed_forward_all_layers(X_test)` and
`collect_W()`.
"""
import numpy as np

from experiments.utils_training import (MCEM_Q_maximizer, SampleWindow, _repeat, _run_sampler)


def _check_sizes(train_size, batch_size):
    if train_size % batch_size != 0:
        raise ValueError(f"In the demo, train size {train_size} should be exactly divided by batch size {batch_size}.")


class _Collector:
    """Per-sample hook: regression lines of all layers at X_test and the W dictionary."""

    def __init__(self, model_demo, X_test):
        self.X_test = X_test
        self.lines = []
        self.W = {'W_' + str(i): [] for i in range(model_demo.n_hidden_layers)}

    def __call__(self, model_demo, log_p, mse):
        self.lines.append(model_demo.feed_forward_all_layers(self.X_test))
        W_sampled = model_demo.collect_W()
        for k in self.W:
            self.W[k].append(W_sampled[k])


def regression_train_demo(model_demo, ds_train, ds_test, train_size, batch_size, X_test,
                          lr_0=0.01, momentum_decay=0.9,
                          resample_in_cycle_head=True,
                          total_epochs=5000, start_sampling_epoch=2000, epochs_per_cycle=50,
                          print_epoch_cycle=100, *, verbose=True, graph=None):
    """W-only cyclical SGHMC with the identity preconditioner; returns (log_p [S, N], mse [S, N], lines, W):
    lines[s][l] = output of GP layer l at X_test for sample s, W['W_l'][s] = ndarray [F_l, g_l]."""
    _check_sizes(train_size, batch_size)
    col = _Collector(model_demo, X_test)
    _, log_p, mse = _run_sampler(model_demo, ds_train, ds_test, train_size, lr_0, momentum_decay, False, 'identity', None, None,
                                 resample_in_cycle_head, total_epochs, start_sampling_epoch, epochs_per_cycle, print_epoch_cycle,
                                 1.0, "reg", False, verbose, graph, on_sample=col)
    return log_p, mse, col.lines, col.W


def MCEM_sampler_demo(model_demo, ds_train, ds_test, train_size, batch_size, X_test,
                      lr_0=0.01, momentum_decay=0.9, resample_in_cycle_head=False,
                      start_sampling_epoch=0, epochs_per_cycle=50, *, verbose=True, graph=None):
    def sampler(num_samples=100, return_lines_Wdict=False, print_epoch_cycle=100):
        _check_sizes(train_size, batch_size)
        col = _Collector(model_demo, X_test) if return_lines_Wdict else None
        W_samples, log_p, mse = _run_sampler(model_demo, ds_train, ds_test, train_size, lr_0, momentum_decay, False, 'identity',
                                             None, None, resample_in_cycle_head,
                                             start_sampling_epoch + num_samples * epochs_per_cycle, start_sampling_epoch,
                                             epochs_per_cycle, print_epoch_cycle, 1.0, "reg", True, verbose, graph, on_sample=col)
        if return_lines_Wdict:
            return W_samples, log_p, mse, col.lines, col.W
        return W_samples, log_p, mse
    return sampler


def MCEM_Q_maximizer_demo(model_demo, data_size, optimizer):
    """The M-step of the demos is the M-step of the UCI drivers (utils_training_demo.py:171-191 == utils_training.py:339-359)."""
    return MCEM_Q_maximizer(model_demo, data_size, optimizer)


def MCEM_demo(sampler_EM, maximizer, sampler_fixing_hyper, total_EM_steps, ds_train,
              num_samples_EM=100, num_samples_fixing_hyper=200,
              print_epoch_cycle_EM=100, print_epoch_cycle_fixing=100):
    em_step = 0
    for x_batch, y_batch in _repeat(ds_train):
        em_step += 1
        print("#" * 15, f"EM step {em_step} of total {total_EM_steps} steps. E Step: ", "#" * 15)
        W_samples, _, _ = sampler_EM(num_samples=num_samples_EM, return_lines_Wdict=False,
                                     print_epoch_cycle=print_epoch_cycle_EM)
        print("#" * 15, f"EM step {em_step} of total {total_EM_steps} steps, M Step: ", "#" * 15)
        maximizer(W_samples, x_batch, y_batch)
        if em_step == total_EM_steps:
            break
    print("#" * 15, f"After {total_EM_steps} EM steps, fixing hyperparams and sample from posterior.", "#" * 15)
    _, log_p, mse, lines, W_dict = sampler_fixing_hyper(num_samples=num_samples_fixing_hyper, return_lines_Wdict=True,
                                                        print_epoch_cycle=print_epoch_cycle_fixing)
    return log_p, mse, lines, W_dict


def MCEM_windows_demo(sampler_EM, maximizer, sampler_fixing_hyper, total_EM_steps, ds_train,
                      num_samples_fixing_hyper=200, window_size=50,
                      print_epoch_cycle_EM=100, print_epoch_cycle_fixing=100):
    from experiments.utils_training import _model_of
    model = _model_of(sampler_EM, maximizer)
    window = SampleWindow(model, window_size)
    em_step = 0
    for x_batch, y_batch in _repeat(ds_train):
        em_step += 1
        print("#" * 15, f"EM step {em_step} of total {total_EM_steps} steps. E Step: ", "#" * 15)
        W_samples, log_p, mse = sampler_EM(num_samples=1, return_lines_Wdict=False, print_epoch_cycle=print_epoch_cycle_EM)
        window.push(W_samples[-1], log_p[-1], mse[-1])
        predict_log_p, predict_rmse = window.average(aux_is_se=True)
        print("*" * 20, " End of E step ", "*" * 20)
        print(f"Number of all sampled models in window: {len(window)} ")
        print(f"Test Log Likelihood of all models in window: {predict_log_p}")
        print(f"Test Root MSE of all models in window: {predict_rmse}\n")
        print("#" * 15, f"EM step {em_step} of total {total_EM_steps} steps, M Step: ", "#" * 15)
        i = np.random.randint(len(window))
        maximizer(window.pick(i), x_batch, y_batch)
        if em_step == total_EM_steps:
            break
    print("#" * 15, f"After {total_EM_steps} EM steps, fixing hyperparams and sample from posterior.", "#" * 15)
    _, log_p, mse, lines, W_dict = sampler_fixing_hyper(num_samples=num_samples_fixing_hyper, return_lines_Wdict=True,
                                                        print_epoch_cycle=print_epoch_cycle_fixing)
    return log_p, mse, lines, W_dict
