"""GPU parity: the CUDA path (through the drop-in classes -> C ABI) against the fp64 CPU oracle
on identical inputs and identical injected noise.  Tolerance for DGPRF_PREC_FP32, stated by
BASELINE.json's north_star: rtol 1e-4 (here: max-abs error / max-abs reference <= 1e-4)."""
import glob
import math
import os

import numpy as np
import pytest
import torch

import dgprf_oracle as O
from helpers import CONFIGS, make_model, oracle_params, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-4
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", list(CONFIGS))
def test_forward_and_loglik(name):
    model, X, Y, c = make_model(name)
    p = oracle_params(model)
    Xd, Yd = X.double(), Y.double()
    F_ref = O.bnn_forward(p, Xd)
    F = model.BNN(X)
    assert rel_err(F, F_ref) < RTOL
    ll_ref = O.log_likelihood(p, Xd, Yd)
    assert rel_err(model.log_likelihood(X, Y), ll_ref) < RTOL
    u_ref = O.U(p, Xd, Yd, c["N"])
    assert float(model.U(X, Y, c["N"])) == pytest.approx(float(u_ref), rel=RTOL)
    u_fb = O.U(p, Xd, Yd, c["N"], full_bayesian=True)
    assert float(model.U(X, Y, c["N"], full_bayesian=True)) == pytest.approx(float(u_fb), rel=RTOL)
    u_em = O.U(p, Xd, Yd, c["N"], allow_gradient_from_W=False)
    assert float(model.U(X, Y, c["N"], allow_gradient_from_W=False)) == pytest.approx(float(u_em), rel=RTOL)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_layerwise_ops_match_oracle(name):
    """The classes called one at a time (RBFLayer/ARCLayer/GPLayer/likelihood.log_prob)."""
    model, X, Y, c = make_model(name)
    p = oracle_params(model)
    rf, gp = model.BNN.layers[0], model.BNN.layers[1]
    Phi_ref = O.rf_layer(p.kinds[0], X.double(), p.z[0], p.log_inv_ls[0], p.log_amp[0], p.mean[0])
    Phi = rf(X)
    assert Phi.shape == (X.shape[0], rf.n_rf)
    assert rel_err(Phi, Phi_ref) < RTOL
    assert rel_err(gp(Phi), Phi_ref @ p.W[0]) < RTOL
    F_ref = O.bnn_forward(p, X.double())
    assert rel_err(model.BNN._call_layerwise(X), F_ref) < RTOL
    if c["task"] == "reg":
        assert rel_err(model.likelihood.log_prob(F_ref.float(), Y), O.gaussian_log_prob(F_ref, Y.double(), p.lik_log_var)) < RTOL
    else:
        assert rel_err(model.likelihood.log_prob(F_ref.float(), Y), O.softmax_log_prob(F_ref, Y.double())) < RTOL
        assert rel_err(model.likelihood.predict_full(F_ref.float()), O.softmax_predict_full(F_ref)) < RTOL


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("full_bayes", [False, True])
def test_gradients(name, full_bayes):
    model, X, Y, c = make_model(name)
    p = oracle_params(model)
    u_ref, g_ref = O.grads_autograd(p, X.double(), Y.double(), c["N"], full_bayes)
    u, g = model.grad_U(X, Y, c["N"], full_bayesian=full_bayes)
    assert float(u) == pytest.approx(float(u_ref), rel=RTOL)
    for n, ref in g_ref.items():
        assert rel_err(g[n], ref) < RTOL, n


@pytest.mark.parametrize("name", list(CONFIGS))
def test_w_gradients_fused_and_layered_kernels(name):
    """dgprf_gradients runs the row-fused step kernel (without its update) where the model is eligible and the layered
    forward / seed / backward kernels otherwise or on request: both against the oracle, and the routing itself."""
    from dgprf import _ffi
    model, X, Y, c = make_model(name)
    p = oracle_params(model)
    u_ref, g_ref = O.grads_autograd(p, X.double(), Y.double(), c["N"], False)
    e = model._engine
    ll_ref = float(O.log_likelihood(p, X.double(), Y.double()).sum())
    for fused in (True, False):
        _ffi.profile_start()
        tot, gW, _ = e.gradients(X, Y, c["N"], hyper=False, prior_w=True, prior_h=False, fused=fused)
        kernels = [nm for nm, _ in _ffi.profile_stop()]
        if fused and max(c["n_gp"]) <= 32:                          # every small config with n_gp <= 32 is K10's
            assert kernels[0] == "k10_step_cluster", kernels
        if not fused:
            assert not any(k.startswith(("k10_", "k9_")) for k in kernels), kernels
        assert float(tot[0]) == pytest.approx(ll_ref, rel=RTOL)
        g = e.named_from_flat(gW, "w")
        for n, ref in g_ref.items():
            assert rel_err(g[n].reshape(ref.shape), ref) < RTOL, (n, fused)


@pytest.mark.parametrize("name", ["protein_small", "mnist_small", "ragged_mixed"])
def test_em_hyper_gradients(name):
    """M-step gradients: W detached, prior term zero (experiments/utils_training.py:345-354)."""
    model, X, Y, c = make_model(name)
    p = oracle_params(model)
    q, gq = O.em_q_and_grads(p, [p.W], X.double(), Y.double(), c["N"])
    u, g = model.grad_U(X, Y, c["N"], full_bayesian=False, allow_gradient_from_W=False, hyper=True)
    assert float(u) == pytest.approx(-float(q), rel=RTOL)
    for n, ref in gq.items():
        assert rel_err(g[n], ref) < RTOL, n


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("mode", ["sghmc", "sgld", "burnin", "resample", "full_bayes"])
def test_one_sampling_step_with_injected_noise(name, mode):
    model, X, Y, c = make_model(name)
    fb = mode == "full_bayes"
    beta = 0.0 if mode == "sgld" else 0.9
    T = 0.0 if mode == "burnin" else 1.0
    model.precond_update(None, c["N"], precond_type="identity", full_bayesian=fb)
    e = model._engine
    p = oracle_params(model)
    names = e.names(fb)
    g = torch.Generator().manual_seed(7)
    mom = {n: e.view(n, "mom").double().cpu().clone() for n in names}
    eps = {n: torch.randn(e.view(n).shape, generator=g, dtype=torch.float64) for n in names}
    res = {n: torch.randn(e.view(n).shape, generator=g, dtype=torch.float64) for n in names} if mode == "resample" else None
    _, _, p_new, m_new = O.sgmcmc_step(p, mom, X.double(), Y.double(), c["N"], lr=0.02, momentum_decay=beta,
                                       temperature=T, full_bayesian=fb, eps=eps, resample=res)
    model.sgmcmc_update(X, Y, c["N"], lr=0.02, momentum_decay=beta, temperature=T, full_bayesian=fb,
                        eps=eps, resample=res)
    new = dict(p_new.w_named() + p_new.hyper_named())
    for n in names:
        assert rel_err(e.view(n), new[n]) < RTOL, ("theta", n)
        assert rel_err(e.view(n, "mom"), m_new[n]) < RTOL, ("moments", n)
    # nothing outside the watched variables moves, and alignment padding stays exactly zero
    if not fb:
        old = dict(p.hyper_named())
        for n in e.seg_h:
            assert rel_err(e.view(n), old[n]) == 0.0, n
    live = torch.zeros_like(e.theta_w[0], dtype=torch.bool)
    for off, ln, _, _ in e.seg_w.values():
        live[off:off + ln] = True
    assert float(e.theta_w[0][~live].abs().sum()) == 0.0 and float(e.mom_w[0][~live].abs().sum()) == 0.0


def test_golden_step_fixtures():
    """Committed fp64 oracle fixtures (tests/golden/step_*.npz, written by make_golden.py)."""
    from make_golden import CASES, case_inputs
    from models.classification_model import ClassificationDGP
    from models.regression_model import RegressionDGP
    for f in sorted(glob.glob(os.path.join(GOLDEN, "step_*.npz"))):
        ref = np.load(f)
        name = str(ref["case"])
        d_in, d_out, L, n_rf, n_gp, kinds, cat, lik, mean, B, N, fb, beta = CASES[name]
        p, X, Y, N, fb, beta, mom, eps = case_inputs(name)
        cls = RegressionDGP if lik == "gaussian" else ClassificationDGP
        model = cls(d_in, d_out, n_hidden_layers=L, n_rf=n_rf, n_gp=n_gp, kernel_type_list=kinds, input_cat=cat,
                    set_nonzero_mean=mean)
        e = model._engine
        for l in range(L):
            e.z[l][0].copy_(p.z[l])
            model._vars[f"W_{l}"].assign(p.W[l])
            model._vars[f"log_amp_{l}"].assign(p.log_amp[l])
            model._vars[f"log_inv_ls_{l}"].assign(p.log_inv_ls[l])
            if mean:
                model._vars[f"mean_{l}"].assign(p.mean[l])
        model.precond_update(None, N, precond_type="identity", full_bayesian=fb)
        for n, m in mom.items():
            model._vars[n].moments = m
        assert rel_err(model.BNN(X.float()), torch.tensor(ref[f"F_{L - 1}"])) < RTOL
        assert float(model.U(X.float(), Y.float(), N, full_bayesian=fb)) == pytest.approx(float(ref["U"]), rel=RTOL)
        model.sgmcmc_update(X.float(), Y.float(), N, lr=0.01, momentum_decay=beta, temperature=1.0,
                            full_bayesian=fb, eps=eps)
        for n in mom:
            assert rel_err(e.view(n), torch.tensor(ref[f"theta_{n}"])) < RTOL, (name, n)
            assert rel_err(e.view(n, "mom"), torch.tensor(ref[f"mom_{n}"])) < RTOL, (name, n)


def test_evaluation_methods():
    model, X, Y, c = make_model("protein_small")
    p = oracle_params(model)
    ds = [(X[:64], Y[:64]), (X[64:], Y[64:])]                      # ragged last batch
    lp, se = model.eval_log_likelihood_and_se(ds)
    lp_ref, se_ref = O.eval_log_likelihood_and_se(p, X.double(), Y.double())
    assert lp.shape == (X.shape[0],) and rel_err(lp, lp_ref) < RTOL and rel_err(se, se_ref) < RTOL
    assert rel_err(model.feed_forward(ds), O.bnn_forward(p, X[64:].double())) < RTOL

    model, X, Y, c = make_model("mnist_small")
    p = oracle_params(model)
    ds = [(X[:100], Y[:100]), (X[100:], Y[100:])]
    assert float(model.eval_all_accuracy(ds)) == pytest.approx(float(O.eval_accuracy(p, X.double(), Y.double())), abs=1e-6)
    assert float(model.eval_batch_accuracy(X, Y)) == pytest.approx(float(O.eval_accuracy(p, X.double(), Y.double())), abs=1e-6)
    assert rel_err(model.eval_log_likelihood(ds), O.eval_log_likelihood(p, X.double(), Y.double())) < RTOL
    acc = float(model.eval_test_free_random(ds))                   # fresh z per forward: just a valid accuracy
    assert 0.0 <= acc <= 1.0 and all(l.random_fixed for l in model.BNN.layers[::2])

    model, X, Y, c = make_model("sin_demo")
    p = oracle_params(model)
    outs = model.feed_forward_all_layers(X)
    Fs, _ = O.bnn_forward(p, X.double(), return_all=True)
    assert len(outs) == 2 and all(rel_err(a, b) < RTOL for a, b in zip(outs, Fs))
    assert set(model.collect_W()) == {"W_0", "W_1"}


@pytest.mark.parametrize("centered", [False, True])
@pytest.mark.parametrize("full_bayes", [False, True])
def test_precond_rmsprop(centered, full_bayes):
    model, X, Y, c = make_model("protein_small")
    model.precond_update(None, c["N"], precond_type="identity", full_bayesian=full_bayes)
    e = model._engine
    p = oracle_params(model)
    names = e.names(full_bayes)
    mom = {n: e.view(n, "mom").double().cpu().clone() for n in names}
    K = 4
    g = torch.Generator().manual_seed(3)
    ds = [(torch.randn(50, 9, generator=g), torch.randn(50, 1, generator=g)) for _ in range(K + 2)]
    grads = [O.grads_autograd(p, x.double(), y.double(), c["N"], full_bayes)[1] for x, y in ds[:K]]
    mass_ref, mom_ref = O.precond_rmsprop(grads, mom, {n: 1.0 for n in names}, centered)
    model.precond_update(ds, c["N"], K_batches=K, full_bayesian=full_bayes, precond_type="rmsprop",
                         second_moment_centered=centered)
    for n in names:
        assert model._vars[n].M == pytest.approx(mass_ref[n], rel=1e-3), n
        assert rel_err(e.view(n, "mom"), mom_ref[n]) < 1e-3, n
    with pytest.raises(AssertionError):
        model.precond_update(ds[:2], c["N"], K_batches=K, precond_type="rmsprop")   # dgp.py:274
    # a step with non-unit masses follows the oracle
    eps = {n: torch.randn(e.view(n).shape, generator=g, dtype=torch.float64) for n in names}
    mom = {n: e.view(n, "mom").double().cpu().clone() for n in names}
    mass = {n: model._vars[n].M for n in names}
    p = oracle_params(model)
    _, _, p_new, m_new = O.sgmcmc_step(p, mom, X.double(), Y.double(), c["N"], lr=0.01, momentum_decay=0.9,
                                       temperature=1.0, full_bayesian=full_bayes, mass=mass, eps=eps)
    model.sgmcmc_update(X, Y, c["N"], lr=0.01, momentum_decay=0.9, full_bayesian=full_bayes, eps=eps)
    new = dict(p_new.w_named() + p_new.hyper_named())
    for n in names:
        assert rel_err(e.view(n), new[n]) < RTOL, n
