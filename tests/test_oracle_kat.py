"""CPU tests that pin the oracle: known-answer scalars recovered from the reference's executed
notebooks, and internal consistency of the hand-derived backward with autograd."""
import json
import math
import os

import numpy as np
import pytest
import torch

import dgprf_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_cyclical_lr_known_answers():
    """lr printed at the end of a cycle by the reference:
    train_regression_demo_sin.ipynb cells 13/7, train_regression_EM_sin.ipynb cell 7."""
    kat = json.load(open(os.path.join(GOLDEN, "notebook_kat.json")))
    for row in kat["cyclical_lr_at_cycle_end"]:
        lr, is_end = O.cyclical_lr(row["lr_0"], row["cycle_length"], row["cycle_length"])
        assert is_end
        assert float(lr) == pytest.approx(row["lr"], rel=2e-6)


def test_cyclical_step_rate_shape_and_errors():
    with pytest.raises(ValueError):
        O.cyclical_step_rate(0, 10)
    with pytest.raises(NotImplementedError):
        O.cyclical_step_rate(1, 10, "nope")
    r1, e1 = O.cyclical_step_rate(1, 50, "cosine", 0.0)
    assert float(r1) == 1.0 and not e1
    r, e = O.cyclical_step_rate(51, 50, "cosine", 0.0)
    assert float(r) == 1.0 and not e                       # new cycle restarts at rate 1
    assert O.cyclical_step_rate(50, 50, "flat")[0] == 1.0
    rates = [float(O.cyclical_step_rate(s, 50, "glide", 0.001)[0]) for s in range(1, 51)]
    assert all(a >= b for a, b in zip(rates, rates[1:]))   # monotone decay within a cycle


def test_default_hyper_init_known_answers():
    """log_amplitude=0.0, log_inv_length_scale=[0.] for d_in=1 (train_regression_EM_sin.ipynb cells 5, 20);
    log(1/sqrt(d)) in general (kernels/RBF.py:16-17,21,40)."""
    p = O.init_params(1, 1, 1, 10, 1)
    assert float(p.log_amp[0]) == 0.0 and p.log_inv_ls[0].tolist() == [0.0]
    p = O.init_params(9, 1, 3, 8, [9, 9, 1], input_cat=True)
    assert [t.numel() for t in p.log_inv_ls] == [9, 18, 18]
    assert float(p.log_inv_ls[1][0]) == pytest.approx(-0.5 * math.log(18.0))
    assert float(p.lik_log_var) == pytest.approx(math.log(0.1))


@pytest.mark.parametrize("kinds,lik,cat,mean", [(["RBF"] * 3, "gaussian", True, False),
                                                (["ARC"] * 3, "softmax", True, True),
                                                (["RBF", "ARC"], "gaussian", False, True)])
@pytest.mark.parametrize("full_bayes", [False, True])
def test_analytic_backward_matches_autograd(kinds, lik, cat, mean, full_bayes):
    L = len(kinds)
    d_out = 2 if lik == "gaussian" else 5
    p = O.init_params(4, d_out, L, 16, [3] * (L - 1) + [d_out], kinds, cat, lik, set_nonzero_mean=mean, seed=3)
    if mean:
        p.mean = [0.3 * torch.randn_like(m) for m in p.mean]
    g = torch.Generator().manual_seed(5)
    X = torch.randn(20, 4, generator=g, dtype=torch.float64)
    Y = torch.randn(20, d_out, generator=g, dtype=torch.float64) if lik == "gaussian" \
        else torch.randint(0, d_out, (20, 1), generator=g).double()
    u, ga = O.grads_autograd(p, X, Y, 100, full_bayes)
    u2, gb = O.grads_analytic(p, X, Y, 100, full_bayes, hyper=full_bayes)
    assert float(u) == pytest.approx(float(u2), rel=1e-13)
    for n in ga:
        assert torch.allclose(ga[n], gb[n].reshape(ga[n].shape), rtol=1e-10, atol=1e-13), n


def test_sgld_is_sghmc_with_zero_decay():
    """beta = 0: theta <- theta - lr/M g + sqrt(2 T lr/(N M)) eps."""
    torch.manual_seed(0)
    th, m, g, eps = (torch.randn(7, 3, dtype=torch.float64) for _ in range(4))
    N, lr, T, M = 500.0, 0.02, 0.7, 2.5
    t_new, m_new = O.sgmcmc_update(th, m, g, N, lr, 0.0, T, M, eps)
    ref = th - lr / M * g + math.sqrt(2 * T * lr / (N * M)) * eps
    assert torch.allclose(t_new, ref, rtol=1e-12, atol=1e-14)


def test_predictive_average_matches_numpy():
    rng = np.random.default_rng(0)
    lp = rng.normal(size=(7, 33)); se = rng.random(size=(7, 33))
    a, b = O.predictive_average(torch.tensor(lp), torch.tensor(se))
    mx = lp.max(0)
    ref = (mx + np.log(np.exp(lp - mx).sum(0)) - np.log(7)).mean()
    assert float(a) == pytest.approx(ref, rel=1e-12)
    assert float(b) == pytest.approx(math.sqrt(se.mean()), rel=1e-12)


def test_precond_rmsprop_scales_minimum_to_one():
    torch.manual_seed(1)
    names = ["W_0", "W_1"]
    grads = [{n: torch.randn(5, 2, dtype=torch.float64) * (1 + 3 * i) for i, n in enumerate(names)} for _ in range(6)]
    mom = {n: torch.randn(5, 2, dtype=torch.float64) for n in names}
    for centered in (False, True):
        mass, new_m = O.precond_rmsprop(grads, mom, {n: 1.0 for n in names}, centered)
        assert min(mass.values()) == pytest.approx(1.0)
        assert mass["W_1"] > mass["W_0"]
        for n in names:
            assert torch.allclose(new_m[n], math.sqrt(mass[n]) * mom[n])


def test_golden_step_vectors_reproduce():
    """tests/golden/step_*.npz were written by tests/golden/make_golden.py from this oracle (fp64);
    they freeze the oracle so an accidental edit of it shows up here."""
    import glob
    files = sorted(glob.glob(os.path.join(GOLDEN, "step_*.npz")))
    assert files, "golden step vectors missing"
    from make_golden import run_case
    for f in files:
        ref = np.load(f)
        got = run_case(str(ref["case"]))
        for k in ref.files:
            if k == "case":
                continue
            assert np.allclose(got[k], ref[k], rtol=1e-9, atol=1e-12), (f, k)


# ---- the reference's own per-cycle printouts (tests/golden/make_notebook_traces.py) -----------------------------------
def _traces():
    return json.load(open(os.path.join(GOLDEN, "notebook_sin_demo_traces.json")))


def sin_demo_data(seed, n_train=60, n_test=100, std_noise=0.02):
    """The notebook's data recipe (train_regression_demo_sin.ipynb cell 1), seeded here (the notebook's draw is not)."""
    rng = np.random.RandomState(seed)
    X = np.concatenate([rng.uniform(-2., -1., n_train // 2), rng.uniform(1., 2., n_train - n_train // 2)])
    Y = np.sin(np.pi * X) + rng.randn(n_train) * std_noise
    y_mean, y_std = np.average(Y), np.std(Y)
    Xt = np.linspace(-5., 5., n_test)
    Yt = (np.sin(np.pi * Xt) - y_mean) / y_std
    f = lambda a: torch.tensor(np.float32(a)).reshape(-1, 1)
    return f(X), f((Y - y_mean) / y_std), f(Xt), f(Yt)


def band(values, burn):
    v = np.asarray(values[burn:], dtype=np.float64)
    return float(np.percentile(v, 5)), float(np.median(v)), float(np.percentile(v, 95))


def test_notebook_ll_rmse_pairs_pin_the_gaussian_likelihood():
    """Every (mean log-likelihood, RMSE) pair the reference printed -- 1040 cycles x (train, test) -- obeys
    LL = -0.5 log(2 pi var) - RMSE^2 / (2 var) with the notebook's variance 0.01: the oracle's Gaussian log-density
    (likelihoods/gaussian.py:20-25, variance kept as log-variance) reproduces each printed LL from the printed RMSE."""
    n = 0
    for run in _traces()["runs"]:
        llv = torch.tensor(math.log(run["config"]["lik_variance"]), dtype=torch.float64)
        for split in ("train", "test"):
            ll = torch.tensor(run[split + "_ll"], dtype=torch.float64)
            rmse = torch.tensor(run[split + "_rmse"], dtype=torch.float64)
            got = O.gaussian_log_prob(torch.zeros(len(rmse), 1, dtype=torch.float64), rmse[:, None], llv)
            # printed in fp32 by TensorFlow: 7 significant digits of LL and of RMSE (RMSE^2 / 0.02 amplifies the latter)
            tol = 2e-6 * (ll.abs() + 1.0) + 4e-6 * rmse ** 2 / run["config"]["lik_variance"]
            assert bool(((got - ll).abs() <= tol).all()), float(((got - ll).abs() / tol).max())
            n += len(ll)
    assert n == 2080


def _oracle_demo_chain(cfg, X, Y, n_cycles, seed):
    """regression_train_demo (experiments/utils_training_demo.py:10-85) restated on the oracle: W only, identity
    preconditioner, cosine cycles from step 1, momentum resampled at every cycle head; returns the train RMSE and the mean
    train log-likelihood of the sample at the end of every cycle (what the notebook prints)."""
    N, B = X.shape[0], 20
    p = O.init_params(1, 1, cfg["n_hidden_layers"], cfg["n_rf"], cfg["n_gp"], None, False, "gaussian",
                      lik_var=cfg["lik_variance"], seed=seed, dtype=torch.float64)
    g = torch.Generator().manual_seed(1000 + seed)
    mom = {n: torch.randn(t.shape, generator=g, dtype=torch.float64) for n, t in p.w_named()}
    it_per_epoch = N // B
    cycle = cfg["epochs_per_cycle"] * it_per_epoch
    rmse, ll = [], []
    Xd, Yd = X.double(), Y.double()
    step_index = 0
    for epoch in range(n_cycles * cfg["epochs_per_cycle"]):
        perm = torch.randperm(N, generator=g)
        for b in range(it_per_epoch):
            idx = perm[b * B:(b + 1) * B]
            step_index += 1
            lr, is_end = O.cyclical_lr(cfg["lr_0"], step_index, cycle)
            eps = {n: torch.randn(t.shape, generator=g, dtype=torch.float64) for n, t in p.w_named()}
            res = None
            if cfg["resample_in_cycle_head"] and step_index % cycle == 1:
                res = {n: torch.randn(t.shape, generator=g, dtype=torch.float64) for n, t in p.w_named()}
            _, _, p, mom = O.sgmcmc_step(p, mom, Xd[idx], Yd[idx], N, lr=float(lr), momentum_decay=cfg["momentum_decay"],
                                         temperature=1.0, eps=eps, resample=res, analytic=True)
            if is_end:
                lp, se = O.eval_log_likelihood_and_se(p, Xd, Yd)
                rmse.append(float(se.mean().sqrt())); ll.append(float(lp.mean()))
    return rmse, ll


def test_oracle_sampler_lands_in_the_band_of_the_reference_run():
    """Statistical pin of the sampling path against a real run of the reference (its data set is unseeded, so this is a band,
    not a vector): the 2-layer sin demo of train_regression_demo_sin.ipynb cell 13 printed, over 1000 cycles, a train RMSE
    with 5 % / median / 95 % = 0.051 / 0.067 / 0.091 and a mean train log-likelihood of 0.97 / 1.16 / 1.25 (after the first
    200 cycles).  The oracle, run through the same driver loop with the same settings on a data set from the same recipe,
    must put its median inside that 5-95 % band.  A sampler with a wrong step size, noise scale, temperature or prior term
    leaves it: e.g. T = 0 gives a median RMSE of ~0.03, twice the noise ~0.11."""
    run = [r for r in _traces()["runs"] if r["cell"] == 13][0]
    lo_r, _, hi_r = band(run["train_rmse"], 200)
    lo_l, _, hi_l = band(run["train_ll"], 200)
    X, Y, _, _ = sin_demo_data(seed=5)
    rmse, ll = _oracle_demo_chain(run["config"], X, Y, n_cycles=24, seed=1)
    _, med_r, _ = band(rmse, 8)
    _, med_l, _ = band(ll, 8)
    print(f"oracle median train RMSE {med_r:.4f} (reference band {lo_r:.4f} .. {hi_r:.4f}), "
          f"median train LL {med_l:.4f} (reference band {lo_l:.4f} .. {hi_l:.4f})")
    assert lo_r <= med_r <= hi_r
    assert lo_l <= med_l <= hi_l
