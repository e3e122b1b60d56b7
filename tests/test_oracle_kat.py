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
