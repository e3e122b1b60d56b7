"""Generates tests/golden/step_*.npz: one full sampling step of the CPU oracle (fp64) on seeded
inputs for scaled-down versions of BASELINE.json's configs.

Provenance: these vectors come from oracle/dgprf_oracle.py, NOT from the reference itself --
TensorFlow is not installable in the build container, so the reference cannot be run
("parity unpinned" at tensor level; the notebook scalars in notebook_kat.json are the only
values that originate from the reference).  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import dgprf_oracle as O  # noqa: E402

CASES = {
    # name: (d_in, d_out, L, n_rf, n_gp, kinds, input_cat, lik, mean, B, N, full_bayes, beta)
    "sin_demo": (1, 1, 2, 100, [1, 1], None, False, "gaussian", False, 20, 60, False, 0.95),
    "protein_small": (9, 1, 3, 64, [9, 9, 1], None, True, "gaussian", False, 100, 45730, False, 0.9),
    "mnist_small_sgld": (40, 10, 3, 48, [30, 30, 10], ["ARC"] * 3, True, "softmax", False, 64, 60000, False, 0.0),
    "full_bayes_mean": (3, 2, 2, [20, 24], [4, 2], ["RBF", "ARC"], True, "gaussian", True, 33, 500, True, 0.9),
}


def case_inputs(name):
    d_in, d_out, L, n_rf, n_gp, kinds, cat, lik, mean, B, N, fb, beta = CASES[name]
    p = O.init_params(d_in, d_out, L, n_rf, n_gp, kinds, cat, lik, set_nonzero_mean=mean, seed=11)
    g = torch.Generator().manual_seed(12)
    if mean:
        p.mean = [0.3 * torch.randn(m.shape, generator=g, dtype=torch.float64) for m in p.mean]
    X = torch.randn(B, d_in, generator=g, dtype=torch.float64)
    Y = torch.randn(B, d_out, generator=g, dtype=torch.float64) if lik == "gaussian" \
        else torch.randint(0, d_out, (B, 1), generator=g).double()
    names = [n for n, _ in O.trainables(p, fb)]
    named = dict(p.w_named() + p.hyper_named())
    mom = {n: torch.randn(named[n].shape, generator=g, dtype=torch.float64) for n in names}
    eps = {n: torch.randn(named[n].shape, generator=g, dtype=torch.float64) for n in names}
    return p, X, Y, N, fb, beta, mom, eps


def run_case(name):
    p, X, Y, N, fb, beta, mom, eps = case_inputs(name)
    Fs, _ = O.bnn_forward(p, X, return_all=True)
    u, g, q, m_new = O.sgmcmc_step(p, mom, X, Y, N, lr=0.01, momentum_decay=beta, temperature=1.0,
                                   full_bayesian=fb, eps=eps)
    o = {"U": np.asarray(float(u))}
    for l, F in enumerate(Fs):
        o[f"F_{l}"] = F.detach().numpy()
    for n in g:
        o[f"grad_{n}"] = g[n].numpy()
        o[f"mom_{n}"] = m_new[n].numpy()
    for n, t in q.w_named() + (q.hyper_named() if fb else []):
        o[f"theta_{n}"] = t.detach().numpy()
    return o


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, f"step_{name}.npz"), case=name, **run_case(name))
        print("wrote", name)
