"""CPU oracle for the DGP-RF-MCMC sampling hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a torch-CPU *restatement* of the reference algorithm
(shixinxing/DGP-RF-MCMC, pure-Python TensorFlow-2 eager code).  It exists so the
CUDA path can be checked; it is never imported by the product package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.

PARITY UNPINNED (tensor level).  TensorFlow is not installed in the build
container and cannot be installed offline, and the reference ships no tests, no
seeds and no tensor-level golden vectors.  The oracle is therefore pinned only by
  * the scalar known-answer values recoverable from executed notebook output
    (cyclical learning rates, default hyper-parameter initialisation; see
    ``tests/test_oracle_kat.py``), and
  * internal consistency: the hand-derived backward (``grads_analytic``) is
    checked against ``torch.autograd`` (the stand-in for ``tf.GradientTape``)
    in fp64.

Every function cites the reference file:line it follows (paths relative to the
reference root).  dtype is a parameter: fp64 is the truth used for tolerances,
fp32 is the "same arithmetic" comparator and the timed CPU baseline.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------- #
# utils.py
# --------------------------------------------------------------------------- #
def log_gaussian(x, mean=0.0, var=1.0):
    """utils.py:46-47  -0.5*(log 2pi + log var + (x-mean)^2/var)."""
    var_t = var if torch.is_tensor(var) else torch.tensor(var, dtype=x.dtype)
    return -0.5 * (LOG_2PI + torch.log(var_t) + (x - mean) ** 2 / var_t)


def cyclical_step_rate(step_index: int, cycle_length: int, schedule: str = "cosine",
                       min_value: float = 0.001):
    """utils.py:49-73.  The reference computes ``frac`` and the cosine in float32
    (``tf.cast(..., tf.float32)``; the Python float ``np.pi`` is converted to the
    tensor dtype), so this restatement does the same with numpy float32 scalars.
    Returns (step_rate: np.float32 | float, is_end: bool)."""
    if step_index <= 0:
        raise ValueError("Step index should be larger than zero!")
    f32 = np.float32
    frac = f32((step_index - 1) % cycle_length) / f32(cycle_length)
    if schedule == "cosine":
        rate = f32(min_value) + f32(1.0 - min_value) * f32(0.5) * (np.cos(f32(np.pi) * frac, dtype=f32) + f32(1.0))
    elif schedule == "glide":
        rate = f32(min_value) + f32(1.0 - min_value) * np.exp(-frac / (f32(1.0) - frac), dtype=f32)
    elif schedule == "flat":
        rate = 1.0
    else:
        raise NotImplementedError
    is_end = (step_index % cycle_length) == 0
    return rate, is_end


def cyclical_lr(lr_0: float, step_index: int, cycle_length: int):
    """experiments/utils_training.py:53-54  lr = lr_0 * rate**2 (cosine, min 0)."""
    rate, is_end = cyclical_step_rate(step_index, cycle_length, "cosine", 0.0)
    return np.float32(lr_0) * (rate ** 2), is_end


# --------------------------------------------------------------------------- #
# Parameter container (explicit tensors instead of tf.Variables on tf.Modules)
# --------------------------------------------------------------------------- #
@dataclass
class DGPParams:
    """All state of one DGP_RF instance (models/dgp.py:9-52, 74-115).

    kinds[l]        'RBF' | 'ARC'                          (dgp.py:45-50)
    z[l]            [d_l, M_l] fixed N(0,1) draw            (rf_layers.py:22,69)
    log_inv_ls[l]   [d_l]  (ARD always: dgp.py:82-83,86-87) (RBF.py:40)
    log_amp[l]      []                                      (RBF.py:39)
    mean[l]         [d_l, 1] zeros unless set_nonzero_mean  (rf_layers.py:24-27)
    W[l]            [F_l, g_l], F_l = 2M (RBF) | M (ARC)    (GP_weight_layers.py:9; dgp.py:103,107)
    lik_log_var     [] or None (Softmax)                    (gaussian.py:12)
    """
    kinds: List[str]
    z: List[torch.Tensor]
    log_inv_ls: List[torch.Tensor]
    log_amp: List[torch.Tensor]
    mean: List[torch.Tensor]
    W: List[torch.Tensor]
    lik_log_var: Optional[torch.Tensor]
    input_cat: bool = False
    mean_trainable: bool = False
    likelihood: str = "gaussian"   # 'gaussian' | 'softmax'

    @property
    def L(self):
        return len(self.kinds)

    def to(self, dtype):
        c = lambda t: None if t is None else t.detach().clone().to(dtype)
        return DGPParams(list(self.kinds), [c(t) for t in self.z], [c(t) for t in self.log_inv_ls],
                         [c(t) for t in self.log_amp], [c(t) for t in self.mean], [c(t) for t in self.W],
                         c(self.lik_log_var), self.input_cat, self.mean_trainable, self.likelihood)

    def hyper_named(self):
        """Named hyper tensors in a fixed order (per layer log_amp, log_inv_ls,
        [mean]; then lik_log_var).  Within-layer order as printed by the
        reference (train_regression_EM_sin.ipynb cell 20)."""
        out = []
        for l in range(self.L):
            out.append((f"log_amp_{l}", self.log_amp[l]))
            out.append((f"log_inv_ls_{l}", self.log_inv_ls[l]))
            if self.mean_trainable:
                out.append((f"mean_{l}", self.mean[l]))
        if self.lik_log_var is not None:
            out.append(("lik_log_var", self.lik_log_var))
        return out

    def w_named(self):
        return [(f"W_{l}", self.W[l]) for l in range(self.L)]


def layer_dims(d_in: int, n_rf: Sequence[int], n_gp: Sequence[int], input_cat: bool):
    """models/dgp.py:76-79  RF-layer input widths d_l."""
    d = [d_in]
    for g in list(n_gp)[:-1]:
        d.append(int(g) + (d_in if input_cat else 0))
    return d


def init_params(d_in, d_out, n_hidden_layers=1, n_rf=20, n_gp=2, kinds=None, input_cat=False,
                likelihood="gaussian", lik_var=0.1, set_nonzero_mean=False, seed=0,
                dtype=torch.float64) -> DGPParams:
    """Default initialisation of the reference: z, W ~ N(0,1) (rf_layers.py:22,
    GP_weight_layers.py:9), log_amp = log 1 = 0 (RBF.py:39), log_inv_ls =
    log(1/sqrt(d_l)) (RBF.py:16-17,21,40), lik_log_var = log 0.1 (gaussian.py:7,12)."""
    L = n_hidden_layers
    n_rf = [n_rf] * L if np.isscalar(n_rf) else list(n_rf)
    n_gp = [n_gp] * L if np.isscalar(n_gp) else list(n_gp)
    assert len(n_rf) == L and len(n_gp) == L
    kinds = ["RBF"] * L if kinds is None else list(kinds)
    d = layer_dims(d_in, n_rf, n_gp, input_cat)
    g = torch.Generator().manual_seed(seed)
    z, ls, la, mu, W = [], [], [], [], []
    for l in range(L):
        M = int(n_rf[l])
        F = 2 * M if kinds[l] == "RBF" else M
        z.append(torch.randn(d[l], M, generator=g, dtype=torch.float64).to(dtype))
        ls.append(torch.full((d[l],), math.log(1.0 / math.sqrt(d[l])), dtype=dtype))
        la.append(torch.zeros((), dtype=dtype))
        mu.append(torch.zeros(d[l], 1, dtype=dtype))
        W.append(torch.randn(F, int(n_gp[l]), generator=g, dtype=torch.float64).to(dtype))
    llv = torch.tensor(math.log(lik_var), dtype=dtype) if likelihood == "gaussian" else None
    return DGPParams(kinds, z, ls, la, mu, W, llv, input_cat, set_nonzero_mean, likelihood)


# --------------------------------------------------------------------------- #
# layers/
# --------------------------------------------------------------------------- #
def rf_layer(kind, X, z, log_inv_ls, log_amp, mean):
    """layers/rf_layers.py:29-45 (RBF) and :75-91 (ARC), ARD branch.
    Omega = exp(log_inv_ls)[:,None]*z + mean;  P = X @ Omega;
    RBF: amp/sqrt(M) * [cos P, sin P]   ARC: sqrt(2)*amp/sqrt(M) * relu(P)."""
    M = z.shape[1]
    Omega = torch.exp(log_inv_ls)[:, None] * z + mean
    P = X @ Omega
    amp = torch.exp(log_amp)
    if kind == "RBF":
        return amp / math.sqrt(M) * torch.cat([torch.cos(P), torch.sin(P)], dim=-1)
    elif kind == "ARC":
        return math.sqrt(2.0) * amp / math.sqrt(M) * torch.relu(P)
    raise NotImplementedError


def bnn_forward(p: DGPParams, X, allow_gradient_from_W=True, return_all=False):
    """utils.py:10-16 (plain chain) / :32-44 (input concat: ``concat([F, X])``
    before every RF layer but the first); GP layer = Phi @ W
    (GP_weight_layers.py:11-15, stop_gradient when not allow_gradient_from_W)."""
    F = X
    Fs, Phis = [], []
    for l in range(p.L):
        inp = F if (l == 0 or not p.input_cat) else torch.cat([F, X], dim=-1)
        Phi = rf_layer(p.kinds[l], inp, p.z[l], p.log_inv_ls[l], p.log_amp[l], p.mean[l])
        W = p.W[l] if allow_gradient_from_W else p.W[l].detach()
        F = Phi @ W
        Phis.append(Phi)
        Fs.append(F)
    if return_all:
        return Fs, Phis
    return F


# --------------------------------------------------------------------------- #
# likelihoods/
# --------------------------------------------------------------------------- #
def gaussian_log_prob(F, Y, lik_log_var):
    """likelihoods/gaussian.py:18-25  sum_D log N(y; f, exp(lik_log_var)) -> [B]."""
    return log_gaussian(Y, mean=F, var=torch.exp(lik_log_var)).sum(dim=-1)


def softmax_log_prob(F, Y):
    """likelihoods/softmax.py:8-15  -sparse_softmax_xent(int(Y[:,0]), F) -> [B]."""
    labels = Y[:, 0].to(torch.int64)
    return -torch.nn.functional.cross_entropy(F, labels, reduction="none")


def softmax_predict_full(F):
    """likelihoods/softmax.py:17-22."""
    return torch.softmax(F, dim=-1)


def log_likelihood(p: DGPParams, X, Y, allow_gradient_from_W=True):
    """models/dgp.py:118-127 -> [B]."""
    F = bnn_forward(p, X, allow_gradient_from_W)
    if p.likelihood == "gaussian":
        return gaussian_log_prob(F, Y, p.lik_log_var)
    return softmax_log_prob(F, Y)


# --------------------------------------------------------------------------- #
# models/dgp.py : potential, gradients, update
# --------------------------------------------------------------------------- #
def prior_W(p: DGPParams):
    """models/dgp.py:129-136."""
    return sum(log_gaussian(w).sum() for w in p.W)


def trainables(p: DGPParams, full_bayesian: bool):
    """W-only: W_0..W_{L-1} (dgp.py:68,195).  Full-Bayes: every trainable
    variable (dgp.py:201); returned as (name, tensor) pairs -- parity tests key
    noise by NAME because tf.Module flatten order cannot be verified here."""
    return p.w_named() + (p.hyper_named() if full_bayesian else [])


def U(p: DGPParams, X, Y, data_size, full_bayesian=False, allow_gradient_from_W=True):
    """models/dgp.py:161-182  U = -(log_prior/N + sum_i ll_i / B)."""
    B = float(X.shape[0])
    N = float(data_size)
    if not full_bayesian:
        log_prior = prior_W(p) / N if allow_gradient_from_W else 0.0
        ll = log_likelihood(p, X, Y, allow_gradient_from_W).sum() / B
    else:
        assert allow_gradient_from_W
        log_prior = 0.0
        for _, t in trainables(p, True):
            log_prior = log_prior + log_gaussian(t).sum() / N
        ll = log_likelihood(p, X, Y).sum() / B
    return -(log_prior + ll)


def grads_autograd(p: DGPParams, X, Y, data_size, full_bayesian=False, allow_gradient_from_W=True,
                   wrt: Optional[List[str]] = None):
    """tf.GradientTape of models/dgp.py:194-204 restated with torch.autograd.
    Returns (U value, {name: grad})."""
    q = p.to(X.dtype)
    named = dict(q.w_named() + q.hyper_named())
    if wrt is None:
        wrt = [n for n, _ in trainables(q, full_bayesian)]
    for n in wrt:
        named[n].requires_grad_(True)
    u = U(q, X, Y, data_size, full_bayesian, allow_gradient_from_W)
    gs = torch.autograd.grad(u, [named[n] for n in wrt], allow_unused=True)
    return u.detach(), {n: (torch.zeros_like(named[n]) if g is None else g.detach()) for n, g in zip(wrt, gs)}


def grads_analytic(p: DGPParams, X, Y, data_size, full_bayesian=False, hyper=False,
                   allow_gradient_from_W=True):
    """Hand-derived reverse pass of U -- the formulas the CUDA kernels implement
    (SURVEY.md section 3.2).  ``hyper`` adds the kernel / likelihood
    hyper-parameter gradients; ``full_bayesian`` additionally adds their prior
    terms theta/N.  With allow_gradient_from_W=False (EM M-step,
    experiments/utils_training.py:345-352) the prior term is zero."""
    with torch.no_grad():
        B = X.shape[0]
        N = float(data_size)
        Fs, Phis = bnn_forward(p, X, return_all=True)
        L = p.L
        ins = [X] + [torch.cat([Fs[l - 1], X], dim=-1) if p.input_cat else Fs[l - 1] for l in range(1, L)]
        F = Fs[-1]
        g = {}
        if p.likelihood == "gaussian":
            var = torch.exp(p.lik_log_var)
            r = Y - F
            ll = (-0.5 * (LOG_2PI + p.lik_log_var + r * r / var)).sum(-1)
            dF = -(r / var) / B
            if hyper:
                g["lik_log_var"] = (0.5 * (1.0 - r * r / var)).sum() / B
                if full_bayesian:
                    g["lik_log_var"] = g["lik_log_var"] + p.lik_log_var / N
        else:
            labels = Y[:, 0].to(torch.int64)
            lse = torch.logsumexp(F, dim=-1)
            ll = F.gather(1, labels[:, None])[:, 0] - lse
            dF = torch.softmax(F, -1)
            dF[torch.arange(B), labels] -= 1.0
            dF = dF / B
        prior = 0.0
        if allow_gradient_from_W:
            prior = sum(log_gaussian(w).sum() for w in p.W) / N
            if full_bayesian:
                for n, t in p.hyper_named():
                    prior = prior + log_gaussian(t).sum() / N
        u = -(prior + ll.sum() / B)
        for l in range(L - 1, -1, -1):
            Phi, W, z = Phis[l], p.W[l], p.z[l]
            M = z.shape[1]
            gW = Phi.T @ dF
            if allow_gradient_from_W:
                gW = gW + W / N
            g[f"W_{l}"] = gW
            dPhi = dF @ W.T
            if hyper:
                g[f"log_amp_{l}"] = (dF * Fs[l]).sum() + (p.log_amp[l] / N if full_bayesian else 0.0)
            if p.kinds[l] == "RBF":
                dP = Phi[:, :M] * dPhi[:, M:] - Phi[:, M:] * dPhi[:, :M]
            else:
                dP = dPhi * (math.sqrt(2.0) * torch.exp(p.log_amp[l]) / math.sqrt(M)) * (Phi > 0).to(Phi.dtype)
            s = torch.exp(p.log_inv_ls[l])
            T = dP @ z.T                       # [B, d_l]
            R = dP.sum(-1, keepdim=True)       # [B, 1]
            if hyper:
                g[f"log_inv_ls_{l}"] = s * (ins[l] * T).sum(0) + (p.log_inv_ls[l] / N if full_bayesian else 0.0)
                if p.mean_trainable:
                    g[f"mean_{l}"] = (ins[l] * R).sum(0)[:, None] + (p.mean[l] / N if full_bayesian else 0.0)
            if l > 0:
                gprev = p.W[l - 1].shape[1]
                dF = (s[None, :] * T + p.mean[l][:, 0][None, :] * R)[:, :gprev]
        return u, g


def sgmcmc_update(theta, moments, grad, data_size, lr, momentum_decay, temperature, mass=1.0,
                  eps=None, resample=None):
    """models/dgp.py:206-216 for ONE parameter tensor.
        h = sqrt(lr/N); [m <- resample]; m <- beta*m - h*N*g + sqrt(2(1-beta) T M)*eps;
        theta <- theta + h/M * m
    ``eps`` / ``resample`` are the injected N(0,1) draws (dgp.py:210,212).
    Returns (theta_new, m_new)."""
    N = float(data_size)
    h = math.sqrt(lr / N)
    m = moments if resample is None else resample
    m_new = momentum_decay * m - h * N * grad
    if eps is not None:
        m_new = m_new + math.sqrt(2.0 * (1.0 - momentum_decay) * temperature * mass) * eps
    theta_new = theta + h * (1.0 / mass) * m_new
    return theta_new, m_new


def sgmcmc_step(p: DGPParams, moments: dict, X, Y, data_size, lr=0.01, momentum_decay=0.95,
                temperature=1.0, full_bayesian=False, mass: Optional[dict] = None,
                eps: Optional[dict] = None, resample: Optional[dict] = None, analytic=False):
    """One models/dgp.py:184-216 iteration.  Returns (U, grads, new DGPParams, new
    moments dict).  Noise dicts are keyed by parameter name."""
    if analytic:
        u, g = grads_analytic(p, X, Y, data_size, full_bayesian, hyper=full_bayesian)
    else:
        u, g = grads_autograd(p, X, Y, data_size, full_bayesian)
    q = p.to(X.dtype)
    named = dict(q.w_named() + q.hyper_named())
    new_m = {}
    for n, _ in trainables(q, full_bayesian):
        Mn = 1.0 if mass is None else float(mass[n])
        e = None if eps is None else eps[n].reshape(named[n].shape)
        r = None if resample is None else resample[n].reshape(named[n].shape)
        t_new, m_new = sgmcmc_update(named[n], moments[n].reshape(named[n].shape), g[n].reshape(named[n].shape),
                                     data_size, lr, momentum_decay, temperature, Mn, e, r)
        named[n].copy_(t_new)
        new_m[n] = m_new
    return u, g, q, new_m


def precond_rmsprop(grads_per_batch: List[dict], moments: dict, mass_old: dict,
                    second_moment_centered=False):
    """models/dgp.py:243-297.  ``grads_per_batch`` = K dicts of gradients (the K
    tape passes, :252-257).  Welford per element (:259-271), per-tensor scalar
    mass (:276-288), normalised by the minimum (:294-295), momentum rescaled
    through m_c = m/sqrt(M_old) (:247,296).  Returns (mass, moments)."""
    K = len(grads_per_batch)
    names = list(moments.keys())
    m_c = {n: moments[n] / math.sqrt(float(mass_old[n])) for n in names}
    mean = {n: torch.zeros_like(moments[n]) for n in names}
    m2 = {n: torch.zeros_like(moments[n]) for n in names}
    for k, gk in enumerate(grads_per_batch, start=1):
        for n in names:
            gr = gk[n].reshape(mean[n].shape)
            delta = gr - mean[n]
            mean[n] = mean[n] + delta / k
            m2[n] = m2[n] + delta * (gr - mean[n])
    est = {}
    for n in names:
        if second_moment_centered:
            sq = (m2[n] / float(K - 1)).mean()
        else:
            sq = (mean[n] ** 2 + m2[n] / float(K)).mean()
        est[n] = math.sqrt(float(sq) + 1.0e-7)
    mn = min(est.values())
    mass = {n: est[n] / mn for n in names}
    new_m = {n: math.sqrt(mass[n]) * m_c[n] for n in names}
    return mass, new_m


# --------------------------------------------------------------------------- #
# evaluation / predictive averaging
# --------------------------------------------------------------------------- #
def eval_log_likelihood_and_se(p: DGPParams, X, Y):
    """models/regression_model.py:33-50 -> (log_p [N], se [N]); se is the MEAN
    over D_out (:46)."""
    with torch.no_grad():
        F = bnn_forward(p, X)
        return gaussian_log_prob(F, Y, p.lik_log_var), ((Y - F) ** 2).mean(-1)


def eval_log_likelihood(p: DGPParams, X, Y):
    """models/classification_model.py:49-60."""
    with torch.no_grad():
        return softmax_log_prob(bnn_forward(p, X), Y)


def eval_accuracy(p: DGPParams, X, Y):
    """models/classification_model.py:17-41: argmax(softmax(F)) == label."""
    with torch.no_grad():
        pred = softmax_predict_full(bnn_forward(p, X)).argmax(-1).to(Y.dtype)
        return (pred == Y.reshape(-1)).to(X.dtype).mean()


def predictive_average(log_p, se=None):
    """experiments/utils_training.py:79-85: mean_n(logsumexp_s log_p - log S),
    sqrt(mean(se)) over all samples and points."""
    S = log_p.shape[0]
    lp = (torch.logsumexp(log_p, dim=0) - math.log(S)).mean()
    rmse = None if se is None else torch.sqrt(se.mean())
    return lp, rmse


# --------------------------------------------------------------------------- #
# stochastic-EM M-step
# --------------------------------------------------------------------------- #
def em_q_and_grads(p: DGPParams, W_samples: List[List[torch.Tensor]], X, Y, data_size):
    """experiments/utils_training.py:339-359: Q = mean_s -U(X,Y; W_s) with W
    detached and the prior term zero; returns (Q, {hyper name: d(-Q)/d hyper})."""
    q = p.to(X.dtype)
    hn = q.hyper_named()
    for _, t in hn:
        t.requires_grad_(True)
    Q = 0.0
    for Ws in W_samples:
        q.W = [w.to(X.dtype) for w in Ws]
        Q = Q + (-U(q, X, Y, data_size, full_bayesian=False, allow_gradient_from_W=False))
    Q = Q / float(len(W_samples))
    gs = torch.autograd.grad(-Q, [t for _, t in hn], allow_unused=True)
    return Q.detach(), {n: (torch.zeros_like(t) if g is None else g.detach()) for (n, t), g in zip(hn, gs)}


def adam_step(theta, g, m, v, t, lr=0.01, b1=0.9, b2=0.999, eps=1e-7):
    """keras.optimizers.Adam defaults (epsilon 1e-7), as used by the notebooks
    (train_regression_EM_sin.ipynb cell 6):  lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    theta -= lr_t * m/(sqrt(v)+eps)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return theta - lr_t * m / (torch.sqrt(v) + eps), m, v


# --------------------------------------------------------------------------- #
# timed CPU baseline: op-for-op eager fp32 step (bench.py cpu_baseline / --impl reference)
# --------------------------------------------------------------------------- #
def eager_step_fp32(p: DGPParams, moments: dict, X, Y, data_size, lr, momentum_decay, temperature,
                    full_bayesian=False, gen: Optional[torch.Generator] = None):
    """Unfused eager execution like the reference's TF path: autograd tape over
    U, then a Python loop over parameter tensors with two randn draws each
    (models/dgp.py:207-216).  Mutates p / moments in place; returns U."""
    named = dict(p.w_named() + p.hyper_named())
    names = [n for n, _ in trainables(p, full_bayesian)]
    ts = [named[n].requires_grad_(True) for n in names]
    u = U(p, X, Y, data_size, full_bayesian)
    gs = torch.autograd.grad(u, ts)
    N = float(data_size)
    h = math.sqrt(lr / N)
    with torch.no_grad():
        for n, t, g in zip(names, ts, gs):
            m_new = momentum_decay * moments[n] - h * N * g
            eps = torch.randn(t.shape, generator=gen, dtype=t.dtype)
            m_new = m_new + math.sqrt(2.0 * (1.0 - momentum_decay) * temperature * 1.0) * eps
            moments[n] = m_new
            t.add_(h * m_new)
    for t in ts:
        t.requires_grad_(False)
    return u.detach()
