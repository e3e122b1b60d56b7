"""Sampler loop off the host: one CUDA-graph launch per epoch, the device ring buffer of stored samples,
the moving-window MCEM drivers and the demo twins (experiments/utils_training.py:41-66, 381-473;
experiments/utils_training_demo.py:10-259)."""
import math

import numpy as np
import pytest
import torch

from experiments.utils_dataset import DeviceDataset
from experiments.utils_training import (Adam, EpochGraph, MCEM_increasing_windows, MCEM_Q_maximizer, MCEM_sampler,
                                        MCEM_windows, SampleWindow, predictive_average, regression_train)
from experiments.utils_training_demo import (MCEM_demo, MCEM_Q_maximizer_demo, MCEM_sampler_demo, MCEM_windows_demo,
                                             regression_train_demo)
from helpers import make_model
from models.regression_model import DemoRegressionDGP, RegressionDGP

pytestmark = pytest.mark.gpu


def _toy(seed=0, N=120, Nt=40, D=3):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(N + Nt, D, generator=g)
    Y = torch.sin(X.sum(-1, keepdim=True)) + 0.1 * torch.randn(N + Nt, 1, generator=g)
    return X[:N], Y[:N], X[N:], Y[N:]


def _model(seed, D=3, cls=RegressionDGP, input_cat=True):
    torch.manual_seed(seed)
    m = cls(D, 1, n_hidden_layers=2, n_rf=40, n_gp=[3, 1], input_cat=input_cat)
    m.seed(seed)
    return m


def test_graph_epoch_equals_the_same_steps_issued_one_by_one():
    X, Y, _, _ = _toy()
    a, b = _model(1), _model(1)
    for m in (a, b):
        m.precond_update(None, 120, precond_type="identity")
    b._engine.mom_w.copy_(a._engine.mom_w)
    dsa = DeviceDataset(X, Y, 30, seed=5); dsb = DeviceDataset(X, Y, 30, seed=5)
    lrs = [0.01, 0.008, 0.006, 0.004]
    dsb.reshuffle()
    g = EpochGraph(b, dsb, 120, lrs, 0.9, 1.0, True, False)            # capture does not execute anything
    assert torch.equal(a._engine.theta_w, b._engine.theta_w)
    for epoch in range(3):
        for i, (xb, yb) in enumerate(dsa):                              # eager: reshuffles at the start of the pass
            a.sgmcmc_update(xb, yb, 120, lr=lrs[i], momentum_decay=0.9, temperature=1.0, resample_moments=(i == 0))
        if epoch > 0:
            dsb.reshuffle()
        g.replay()                                                      # graph: ONE launch for the four steps
    torch.cuda.synchronize()
    assert a._step == b._step == 12
    assert torch.equal(a._engine.theta_w, b._engine.theta_w) and torch.equal(a._engine.mom_w, b._engine.mom_w)
    assert torch.isfinite(a._engine.theta_w).all()


@pytest.mark.parametrize("variant", ["tf32", "full_bayes"])
def test_graph_epoch_of_the_layered_paths(variant):
    """The layered kernels (tensor-core mode; full-Bayesian step) launch with programmatic stream serialization: a captured
    epoch of them must replay to exactly what the same steps give one by one."""
    X, Y, _, _ = _toy(3, N=512, D=9)
    fb = variant == "full_bayes"

    def make():
        torch.manual_seed(4)
        m = RegressionDGP(9, 1, n_hidden_layers=2, n_rf=256, n_gp=[9, 1], input_cat=True)
        m.seed(4)
        if variant == "tf32":
            m.set_precision("tf32")
        m.precond_update(None, 512, precond_type="identity", full_bayesian=fb)
        return m
    a, b = make(), make()
    b._engine.mom_w.copy_(a._engine.mom_w)
    b._engine.mom_h.copy_(a._engine.mom_h)
    dsa = DeviceDataset(X, Y, 256, seed=7); dsb = DeviceDataset(X, Y, 256, seed=7)
    lrs = [0.004, 0.003]
    dsb.reshuffle()
    g = EpochGraph(b, dsb, 512, lrs, 0.9, 1.0, False, fb)
    for epoch in range(3):
        for i, (xb, yb) in enumerate(dsa):
            a.sgmcmc_update(xb, yb, 512, lr=lrs[i], momentum_decay=0.9, temperature=1.0, full_bayesian=fb)
        if epoch > 0:
            dsb.reshuffle()
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(a._engine.theta_w, b._engine.theta_w) and torch.equal(a._engine.mom_w, b._engine.mom_w)
    assert torch.equal(a._engine.theta_h, b._engine.theta_h)
    assert torch.isfinite(a._engine.theta_w).all()


@pytest.mark.parametrize("resample", [False, True])
def test_driver_with_graph_epochs_is_bit_identical_to_the_eager_driver(resample):
    X, Y, Xt, Yt = _toy(2)
    res = []
    for graph in (False, True):
        m = _model(7)
        ds_train = DeviceDataset(X, Y, 40, seed=11)
        ds_test = DeviceDataset(Xt, Yt, 64, shuffle=False, drop_remainder=False)
        log_p, mse = regression_train(m, data=(ds_train, ds_test, 120, 2.0), lr_0=0.02, momentum_decay=0.9, full_bayesian=False,
                                      resample_in_cycle_head=resample, total_epochs=14, start_sampling_epoch=4,
                                      epochs_per_cycle=5, verbose=False, graph=graph)
        res.append((log_p.as_subclass(torch.Tensor).clone(), mse.as_subclass(torch.Tensor).clone(), m._engine.theta_w.clone()))
    assert res[0][0].shape == (2, 40)
    for x, y in zip(res[0], res[1]):
        assert torch.equal(x, y)
    with pytest.raises(ValueError):                                     # masses of a preconditioner would be baked in
        regression_train(_model(7), data=(DeviceDataset(X, Y, 40), [(Xt.cuda(), Yt.cuda())], 120), precond_type='rmsprop',
                         K_batches=2, second_moment_centered=False, total_epochs=1, start_sampling_epoch=1, verbose=False,
                         graph=True)


def test_sample_window_is_a_sliding_window():
    m = _model(3)
    w_len = m._engine.layout.w_len
    win = SampleWindow(m, 4)
    g = torch.Generator().manual_seed(0)
    hist = []
    for s in range(7):
        W = torch.randn(w_len, generator=g).cuda(); lp = torch.randn(9, generator=g).cuda(); se = torch.rand(9, generator=g).cuda()
        hist.append((W, lp, se))
        win.push(W, lp, se)
        live = hist[-4:]                                                # utils_training.py:399-405: concat, then drop row 0
        assert len(win) == len(live)
        Ws, lps, ses = win.stored()
        assert sorted(Ws.sum(1).tolist()) == sorted(torch.stack([h[0] for h in live]).sum(1).tolist())
        a, b = win.average(aux_is_se=True)
        a_ref, b_ref = predictive_average(torch.stack([h[1] for h in live]), torch.stack([h[2] for h in live]))
        assert a == pytest.approx(a_ref, rel=1e-6) and b == pytest.approx(b_ref, rel=1e-6)
    assert win.count == 7 and win.pick(2).shape == (1, w_len)


@pytest.mark.parametrize("driver", [MCEM_windows, MCEM_increasing_windows])
def test_mcem_windows(driver, capsys):
    X, Y, Xt, Yt = _toy(4)
    m = _model(9)
    ds_train = DeviceDataset(X, Y, 40, seed=1)
    ds_test = [(Xt.cuda(), Yt.cuda())]
    h0 = m._engine.theta_h.clone()
    sampler = MCEM_sampler(m, ds_train, ds_test, 120, lr_0=0.01, start_sampling_epoch=1, epochs_per_cycle=2)
    maximizer = MCEM_Q_maximizer(m, 120, Adam(0.01))
    np.random.seed(0)
    log_p, mse = driver(sampler, maximizer, sampler, total_EM_steps=5, ds_train=ds_train, num_samples_fixing_hyper=3,
                        window_size=3)
    out = capsys.readouterr().out
    assert log_p.shape == (3, 40) and mse.shape == (3, 40) and torch.isfinite(log_p).all()
    assert not torch.equal(h0, m._engine.theta_h)                       # five M-steps moved the hyper-parameters
    sizes = [int(l.split(":")[1]) for l in out.splitlines() if l.startswith("Number of all sampled models in window")]
    assert sizes == [1, 2, 3, 3, 3]                                     # grows to window_size, then slides
    assert out.count("averaged by 1 samples") == 5                      # the M-step sees ONE window sample (:422-423)


def test_demo_twins(capsys):
    X, Y, Xt, Yt = _toy(6, N=60, Nt=20, D=1)
    m = _model(2, D=1, cls=DemoRegressionDGP, input_cat=False)      # the demos chain the layers without input concatenation
    ds_train = DeviceDataset(X, Y, 20, seed=2)
    ds_test = [(Xt.cuda(), Yt.cuda())]
    Xline = torch.linspace(-2, 2, 15)[:, None]
    with pytest.raises(ValueError):
        regression_train_demo(m, ds_train, ds_test, 60, 25, Xline)
    log_p, mse, lines, W = regression_train_demo(m, ds_train, ds_test, 60, 20, Xline, total_epochs=8, start_sampling_epoch=2,
                                                 epochs_per_cycle=3, print_epoch_cycle=100, verbose=False, graph=True)
    assert log_p.shape == (2, 20) and len(lines) == 2 and len(lines[0]) == 2 and lines[0][1].shape == (15, 1)
    assert set(W) == {"W_0", "W_1"} and len(W["W_0"]) == 2 and W["W_0"][0].shape == (80, 3)
    assert not np.array_equal(W["W_1"][0], W["W_1"][1])                 # distinct stored samples
    sampler = MCEM_sampler_demo(m, ds_train, ds_test, 60, 20, Xline, start_sampling_epoch=1, epochs_per_cycle=2, verbose=False)
    r3 = sampler(num_samples=2)
    assert len(r3) == 3 and len(r3[0]) == 2 and r3[1].shape == (2, 20)
    maximizer = MCEM_Q_maximizer_demo(m, 60, Adam(0.01))
    out1 = MCEM_demo(sampler, maximizer, sampler, 2, ds_train, num_samples_EM=2, num_samples_fixing_hyper=2)
    assert len(out1) == 4 and out1[0].shape == (2, 20) and len(out1[2]) == 2 and set(out1[3]) == {"W_0", "W_1"}
    np.random.seed(1)
    out2 = MCEM_windows_demo(sampler, maximizer, sampler, 3, ds_train, num_samples_fixing_hyper=2, window_size=2)
    assert len(out2) == 4 and out2[1].shape == (2, 20)
    assert "Number of all sampled models in window: 2" in capsys.readouterr().out
