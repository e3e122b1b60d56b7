"""GPU tests of the host-level pieces of the path: predictive averaging (a15), the stochastic-EM M-step (a16)
and the sampler driver loop, against the oracle."""
import math

import pytest
import torch

import dgprf_oracle as O
from experiments.utils_training import (Adam, MCEM, MCEM_Q_maximizer, MCEM_sampler, em_q_and_grads, predictive_average,
                                        regression_train)
from helpers import make_model, oracle_params, rel_err
from utils import cyclical_step_rate

pytestmark = pytest.mark.gpu


def test_predictive_average_function():
    g = torch.Generator().manual_seed(0)
    lp = 2.0 * torch.randn(9, 300, generator=g) - 1.0
    se = torch.rand(9, 300, generator=g)
    a_ref, b_ref = O.predictive_average(lp.double(), se.double())
    a, b = predictive_average(lp, se)
    assert a == pytest.approx(float(a_ref), rel=1e-5) and b == pytest.approx(float(b_ref), rel=1e-5)
    acc = torch.rand(9, generator=g)
    a2, m2 = predictive_average([r for r in lp], acc, aux_is_se=False)
    assert a2 == pytest.approx(float(a_ref), rel=1e-5) and m2 == pytest.approx(float(acc.mean()), rel=1e-5)


@pytest.mark.parametrize("name", ["protein_small", "mnist_small", "ragged_mixed"])
def test_mcem_q_maximizer_matches_oracle(name, capsys):
    model, X, Y, c = make_model(name)
    p = oracle_params(model)
    g = torch.Generator().manual_seed(4)
    S = 5
    W_samples = [[torch.randn(w.shape, generator=g) for w in p.W] for _ in range(S)]
    q_ref, g_ref = O.em_q_and_grads(p, [[w.double() for w in Ws] for Ws in W_samples], X.double(), Y.double(), c["N"])
    Q, gflat = em_q_and_grads(model, W_samples, X, Y, c["N"], max_chunk=2)       # exercises the chunking
    assert float(Q) == pytest.approx(float(q_ref), rel=1e-4)
    e = model._engine
    got = e.named_from_flat(gflat[None], "h")
    for n, ref in g_ref.items():
        assert rel_err(got[n], ref) < 1e-4, n
    # one M-step == keras-Adam on every trainable hyper-parameter, W untouched
    before_W = e.theta_w.clone()
    named = dict(p.hyper_named())
    opt = Adam(learning_rate=0.01)
    maximizer = MCEM_Q_maximizer(model, c["N"], opt)
    maximizer(W_samples, X, Y)
    assert "Q function is" in capsys.readouterr().out
    for n, ref in g_ref.items():
        th, _, _ = O.adam_step(named[n], ref.reshape(named[n].shape), torch.zeros_like(named[n]), torch.zeros_like(named[n]), 1)
        assert rel_err(e.view(n), th) < 1e-4, n
    assert torch.equal(before_W, e.theta_w)


def test_sampler_driver_follows_the_reference_loop():
    """regression_train on the sin-demo shape: shapes / sample count of the cyclical loop, and the burn-in part
    (T = 0, hence deterministic) replayed step for step against an oracle loop."""
    model, X, Y, c = make_model("sin_demo")
    p = oracle_params(model)
    N, Bn = 60, 20
    g = torch.Generator().manual_seed(2)
    Xall, Yall = torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g)
    ds_train = [(Xall[i:i + Bn], Yall[i:i + Bn]) for i in range(0, N, Bn)]
    ds_test = [(Xall[:30], Yall[:30])]
    model.precond_update(None, N, precond_type="identity")
    e = model._engine
    mom = {n: e.view(n, "mom").double().cpu().clone() for n in e.names(False)}
    # oracle: 2 burn-in epochs (T=0) then one cycle of 2 epochs sampled at T=1 is stochastic -> compare burn-in only
    q = p
    for epoch in range(2):
        for xb, yb in ds_train:
            _, _, q, mom = O.sgmcmc_step(q, mom, xb.double(), yb.double(), N, lr=0.01, momentum_decay=0.9,
                                         temperature=0.0, full_bayesian=False, eps=None)
    log_p, mse = regression_train(model, data=(ds_train, ds_test, N), lr_0=0.01, momentum_decay=0.9, full_bayesian=False,
                                  total_epochs=2 + 2, start_sampling_epoch=2, epochs_per_cycle=2, verbose=False,
                                  resample_in_cycle_head=True)
    assert log_p.shape == (1, 30) and mse.shape == (1, 30)            # one cycle -> one posterior sample
    assert torch.isfinite(log_p).all() and torch.isfinite(mse).all()
    # replay: a fresh model with the same state must match the oracle after the burn-in part alone
    model2, _, _, _ = make_model("sin_demo")
    model2.precond_update(None, N, precond_type="identity")
    p2 = oracle_params(model2)
    mom2 = {n: model2._engine.view(n, "mom").double().cpu().clone() for n in model2._engine.names(False)}
    q2 = p2
    for epoch in range(2):
        for xb, yb in ds_train:
            _, _, q2, mom2 = O.sgmcmc_step(q2, mom2, xb.double(), yb.double(), N, lr=0.01, momentum_decay=0.9,
                                           temperature=0.0, full_bayesian=False, eps=None)
            model2.sgmcmc_update(xb, yb, N, lr=0.01, momentum_decay=0.9, temperature=0.)
    for l in range(2):
        assert rel_err(model2._engine.view(f"W_{l}"), q2.W[l]) < 5e-4, l       # 6 chained fp32 steps


def test_mcem_loop_runs_and_moves_hypers():
    model, X, Y, c = make_model("sin_demo")
    N, Bn = 60, 20
    g = torch.Generator().manual_seed(5)
    Xall = torch.rand(N, 1, generator=g) * 6 - 3
    Yall = torch.sin(Xall) + 0.1 * torch.randn(N, 1, generator=g)
    ds_train = [(Xall[i:i + Bn], Yall[i:i + Bn]) for i in range(0, N, Bn)]
    ds_test = [(Xall[:30], Yall[:30])]
    h0 = model._engine.theta_h.clone()
    sampler = MCEM_sampler(model, ds_train, ds_test, N, lr_0=0.01, start_sampling_epoch=2, epochs_per_cycle=2)
    maximizer = MCEM_Q_maximizer(model, N, Adam(0.01))
    log_p, mse = MCEM(sampler, maximizer, sampler, total_EM_steps=2, ds_train=ds_train, num_samples_EM=2,
                      num_samples_fixing_hyper=3)
    assert log_p.shape == (3, 30) and torch.isfinite(log_p).all()
    assert not torch.equal(h0, model._engine.theta_h)                 # the M-steps changed the hyper-parameters
    lp, rmse = predictive_average(log_p, mse)
    assert math.isfinite(lp) and rmse > 0


def test_host_minibatch_paths_match_device_path():
    """sgmcmc_update on pageable host tensors (staging copies), pinned host tensors (zero-copy) and device tensors
    gives the same parameters bit for bit, and the minibatch log-likelihood read back matches the oracle."""
    import dgprf_oracle as O
    res = []
    for kind in ("device", "pageable", "pinned"):
        model, X, Y, c = make_model("protein_small")
        model.seed(5)
        model.precond_update(None, c["N"], precond_type="identity")
        for n in model._engine.names(False):                      # identical momentum in the three runs
            model._vars[n].moments = torch.full(model._vars[n].shape, 0.25)
        u = torch.zeros(1)
        if kind == "device":
            model.sgmcmc_update(X.cuda(), Y.cuda(), c["N"], lr=0.01, momentum_decay=0.9)
        elif kind == "pageable":
            model.sgmcmc_update(X.clone(), Y.clone(), c["N"], lr=0.01, momentum_decay=0.9, u_host=u)
        else:
            u = u.pin_memory()
            model.sgmcmc_update(X.pin_memory(), Y.pin_memory(), c["N"], lr=0.01, momentum_decay=0.9, u_host=u)
        torch.cuda.synchronize()
        res.append((model._engine.theta_w.clone(), float(u[0])))
        if kind == "device":
            ll_ref = float(O.log_likelihood(oracle_params(make_model("protein_small")[0]), X.double(), Y.double()).sum())
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][0], res[2][0])
    assert res[1][1] == pytest.approx(ll_ref, rel=1e-4) and res[2][1] == pytest.approx(ll_ref, rel=1e-4)
