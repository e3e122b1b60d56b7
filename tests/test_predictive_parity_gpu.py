"""Posterior-predictive parity (BASELINE.json north_star: "posterior-predictive RMSE/NLL/accuracy must match within a
stated Monte-Carlo tolerance"; SURVEY section 8c: within 3 standard errors over chains / seeds).

Eight seeded GPU chains run through the drop-in drivers (regression_train / classification_train ->
model.sgmcmc_update -> C ABI, in-kernel Philox noise) and eight oracle chains through a restatement of the same loop
(experiments/utils_training.py:41-85, 121-166) in fp64 with torch-generator noise.  Chains differ in z, W, momentum and
noise, so the comparison is statistical: the chain-mean of each ensemble metric must agree within
3 * sqrt(s_gpu^2 / n + s_oracle^2 / n)  (the stated Monte-Carlo tolerance; s = standard deviation over chains).

Also the SGLD sanity check of the reference's experiments/SGLD-demo.ipynb (cells 2-6): SGLD must find BOTH modes of
the Gaussian mixture 0.5 N(-2, 2) + 0.5 N(2, 0.2) -- here through the update kernel with its in-kernel noise."""
import math

import pytest
import torch

import dgprf_oracle as O
from dgprf import _ffi
from experiments.utils_training import classification_train, predictive_average, regression_train
from models.classification_model import ClassificationDGP
from models.regression_model import RegressionDGP

pytestmark = pytest.mark.gpu
N_CHAINS = 8
SCHED = dict(total_epochs=36, start_sampling_epoch=12, epochs_per_cycle=4)      # 6 posterior samples per chain


def _data(task, seed=0):
    g = torch.Generator().manual_seed(seed)
    N, Nt, D = 160, 120, 3
    X = torch.randn(N + Nt, D, generator=g)
    if task == "reg":
        w = torch.randn(D, 1, generator=g)
        Y = torch.sin(X @ w) + 0.3 * torch.randn(N + Nt, 1, generator=g)
    else:
        w = torch.randn(D, 3, generator=g)
        Y = (X @ w + 0.5 * torch.randn(N + Nt, 3, generator=g)).argmax(-1, keepdim=True).float()
    return X[:N], Y[:N], X[N:], Y[N:]


def _oracle_chain(task, seed, X, Y, Xt, Yt, arch, B, lr_0, beta):
    """The sampler loop of experiments/utils_training.py:41-85 (regression) / :121-166 (classification), W only,
    identity preconditioner, burn-in at T = 0 then cosine cycles, one sample per cycle end."""
    N = X.shape[0]
    p = O.init_params(arch["d_in"], arch["d_out"], arch["L"], arch["n_rf"], arch["n_gp"], arch["kinds"], arch["input_cat"],
                      "gaussian" if task == "reg" else "softmax", seed=seed, dtype=torch.float64)
    g = torch.Generator().manual_seed(10_000 + seed)
    mom = {n: torch.randn(t.shape, generator=g, dtype=torch.float64) for n, t in p.w_named()}
    it_per_epoch = N // B
    cycle = SCHED["epochs_per_cycle"] * it_per_epoch
    log_p, aux = [], []
    for epoch in range(SCHED["total_epochs"]):
        perm = torch.randperm(N, generator=g)                       # tf.data shuffle, drop_remainder
        for b in range(it_per_epoch):
            idx = perm[b * B:(b + 1) * B]
            xb, yb = X[idx].double(), Y[idx].double()
            if epoch < SCHED["start_sampling_epoch"]:
                lr, T, is_end = lr_0, 0.0, False
            else:
                step_index = (epoch - SCHED["start_sampling_epoch"]) * it_per_epoch + b + 1
                lr, is_end = O.cyclical_lr(lr_0, step_index, cycle)
                T = 1.0
            eps = {n: torch.randn(t.shape, generator=g, dtype=torch.float64) for n, t in p.w_named()}
            _, _, p, mom = O.sgmcmc_step(p, mom, xb, yb, N, lr=float(lr), momentum_decay=beta, temperature=T, eps=eps,
                                         analytic=True)
            if is_end:
                if task == "reg":
                    lp, se = O.eval_log_likelihood_and_se(p, Xt.double(), Yt.double())
                else:
                    lp, se = O.eval_log_likelihood(p, Xt.double(), Yt.double()), O.eval_accuracy(p, Xt.double(), Yt.double())
                log_p.append(lp); aux.append(se)
    lp = (torch.logsumexp(torch.stack(log_p), 0) - math.log(len(log_p))).mean()
    m = torch.stack(aux).mean()
    return float(lp), float(m.sqrt() if task == "reg" else m)


def _gpu_chain(task, seed, X, Y, Xt, Yt, arch, B, lr_0, beta):
    torch.manual_seed(seed)
    cls = RegressionDGP if task == "reg" else ClassificationDGP
    model = cls(arch["d_in"], arch["d_out"], n_hidden_layers=arch["L"], n_rf=arch["n_rf"], n_gp=arch["n_gp"],
                kernel_type_list=arch["kinds"], input_cat=arch["input_cat"])
    model.seed(seed)
    N = X.shape[0]
    Xd, Yd = X.cuda(), Y.cuda()
    g = torch.Generator().manual_seed(20_000 + seed)

    class Shuffled:                                                 # a re-iterable dataset: fresh shuffle per epoch, drop_remainder
        def __iter__(self):
            perm = torch.randperm(N, generator=g).cuda()
            for b in range(N // B):
                idx = perm[b * B:(b + 1) * B]
                yield Xd[idx], Yd[idx]
    ds_test = [(Xt.cuda(), Yt.cuda())]
    train = regression_train if task == "reg" else classification_train
    log_p, aux = train(model, data=(Shuffled(), ds_test, N), lr_0=lr_0, momentum_decay=beta, full_bayesian=False,
                       precond_type='identity', resample_in_cycle_head=False, verbose=False, **SCHED)
    assert log_p.shape[0] == (SCHED["total_epochs"] - SCHED["start_sampling_epoch"]) // SCHED["epochs_per_cycle"]
    return predictive_average(log_p, aux, aux_is_se=(task == "reg"))


def _compare(name, a, b):
    ta, tb = torch.tensor(a, dtype=torch.float64), torch.tensor(b, dtype=torch.float64)
    se = math.sqrt(float(ta.var(unbiased=True)) / len(a) + float(tb.var(unbiased=True)) / len(b))
    diff = abs(float(ta.mean() - tb.mean()))
    print(f"{name}: gpu {float(ta.mean()):.4f} +- {float(ta.std()):.4f}   oracle {float(tb.mean()):.4f} +- {float(tb.std()):.4f}   "
          f"|diff| {diff:.4f}   3 SE {3 * se:.4f}")
    assert diff <= 3.0 * se, (name, diff, 3.0 * se)


@pytest.mark.parametrize("task", ["reg", "cls"])
def test_predictive_metrics_match_the_oracle_within_3_standard_errors(task):
    X, Y, Xt, Yt = _data(task)
    if task == "reg":            # 2-layer RBF regression, SGHMC
        arch = dict(d_in=3, d_out=1, L=2, n_rf=24, n_gp=[3, 1], kinds=["RBF", "RBF"], input_cat=True)
        B, lr_0, beta = 40, 0.02, 0.9
    else:                        # 2-layer arc-cosine softmax classification, SGLD (beta = 0), cfg3's sampler
        arch = dict(d_in=3, d_out=3, L=2, n_rf=24, n_gp=[4, 3], kinds=["ARC", "ARC"], input_cat=True)
        B, lr_0, beta = 40, 0.02, 0.0
    gpu = [_gpu_chain(task, 100 + c, X, Y, Xt, Yt, arch, B, lr_0, beta) for c in range(N_CHAINS)]
    orc = [_oracle_chain(task, 200 + c, X, Y, Xt, Yt, arch, B, lr_0, beta) for c in range(N_CHAINS)]
    _compare(f"{task} test log-likelihood", [r[0] for r in gpu], [r[0] for r in orc])
    _compare(f"{task} " + ("RMSE" if task == "reg" else "accuracy"), [r[1] for r in gpu], [r[1] for r in orc])


def test_sgld_finds_both_modes_of_the_mixture():
    """experiments/SGLD-demo.ipynb cells 2-4: x <- x + lr grad log p(x) + sqrt(2 lr) eps with a decaying lr,
    target 0.5 N(-2, 2) + 0.5 N(2, 0.2), start x = 5.  Here 4096 independent walkers ride one parameter buffer; every
    step is ONE dgprf_sgmcmc_update call with momentum_decay = 0 (SGLD), data_size = 1 and in-kernel Philox noise."""
    n = 4096
    dev = torch.device("cuda")
    x = torch.full((1, n), 5.0, device=dev)
    mom = torch.zeros(1, n, device=dev)
    segs = _ffi.make_segments([(0, n, 1.0, 0)])                       # no extra N(0,1) prior term
    m1, v1, m2, v2 = -2.0, 2.0, 2.0, 0.2

    def grad_U(t):                                                    # U = -log p
        a = 0.5 * torch.exp(-0.5 * (t - m1) ** 2 / v1) / math.sqrt(2 * math.pi * v1)
        b = 0.5 * torch.exp(-0.5 * (t - m2) ** 2 / v2) / math.sqrt(2 * math.pi * v2)
        return (a * (t - m1) / v1 + b * (t - m2) / v2) / (a + b)

    L = _ffi.lib()
    T0, T1 = 2500, 1500
    for t in range(T0 + T1):
        # a constant step first so that the walkers (all started in the right-hand mode's basin) mix between the modes,
        # then the notebook's polynomial decay so that the discretisation bias of the narrow mode (var 0.2) vanishes
        lr = 0.05 if t < T0 else 0.05 * (1 + t - T0) ** (-0.55)
        g = grad_U(x).contiguous()
        _ffi.check(L.dgprf_sgmcmc_update(x.data_ptr(), mom.data_ptr(), n, n, 1, g.data_ptr(), n, 1, 0, segs, 1,
                                         lr, 1.0, 0.0, 1.0, 0, 1234, t, None, None, _ffi.stream_ptr()))
    s = x[0].double().cpu()
    # exact moments of the target on either side of the cut (quadrature on a dense grid): the cut truncates the wide mode
    t = torch.linspace(-14.0, 9.0, 400001, dtype=torch.float64)
    dens = 0.5 * torch.exp(-0.5 * (t - m1) ** 2 / v1) / math.sqrt(2 * math.pi * v1) + \
        0.5 * torch.exp(-0.5 * (t - m2) ** 2 / v2) / math.sqrt(2 * math.pi * v2)
    cut = 0.6
    def side(mask):
        w = dens * mask
        mass = float(w.sum())
        mean = float((w * t).sum()) / mass
        return mass / float(dens.sum()), mean, float((w * (t - mean) ** 2).sum()) / mass
    (p_lo, mu_lo, var_lo), (p_hi, mu_hi, var_hi) = side((t <= cut).double()), side((t > cut).double())
    right = float((s > cut).double().mean())
    lo, hi = s[s <= cut], s[s > cut]
    print(f"SGLD mixture: right-mode weight {right:.3f} (exact {p_hi:.3f}); left mean/var {float(lo.mean()):.3f}/{float(lo.var()):.3f} "
          f"(exact {mu_lo:.3f}/{var_lo:.3f}); right mean/var {float(hi.mean()):.3f}/{float(hi.var()):.3f} (exact {mu_hi:.3f}/{var_hi:.3f})")
    assert abs(right - p_hi) < 0.08, right                            # both modes found, with the right weights
    assert abs(float(lo.mean()) - mu_lo) < 0.12 and abs(float(hi.mean()) - mu_hi) < 0.04
    assert abs(float(lo.var()) / var_lo - 1.0) < 0.12 and abs(float(hi.var()) / var_hi - 1.0) < 0.12
