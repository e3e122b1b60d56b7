"""Shared test plumbing: build a drop-in model and the oracle state that mirrors it."""
import math

import torch

import dgprf_oracle as O
from models.classification_model import ClassificationDGP
from models.regression_model import RegressionDGP

# name -> (ctor kwargs, task)  -- scaled-down versions of BASELINE.json's configs + edge shapes
CONFIGS = {
    # cfg1: the reference's CPU demo (train_regression_demo_sin.ipynb): 2-layer RBF, n_rf=100, n_gp=[1,1]
    "sin_demo": dict(task="reg", d_in=1, d_out=1, L=2, n_rf=100, n_gp=[1, 1], kinds=None, input_cat=False, B=20, N=60),
    # cfg2 shape (protein): 3-layer RBF, input_cat, n_gp=[9,9,1]; M reduced for the CPU oracle
    "protein_small": dict(task="reg", d_in=9, d_out=1, L=3, n_rf=96, n_gp=[9, 9, 1], kinds=None, input_cat=True, B=150, N=45730),
    # cfg3 shape (MNIST): 3-layer ARC softmax, input_cat, n_gp=[30,30,10]; d_in reduced
    "mnist_small": dict(task="cls", d_in=50, d_out=10, L=3, n_rf=72, n_gp=[30, 30, 10], kinds=["ARC"] * 3, input_cat=True, B=130, N=60000),
    # ragged everything: M not a multiple of 4, B not a multiple of the tile, mixed kernels, trainable mean
    "ragged_mixed": dict(task="reg", d_in=3, d_out=2, L=3, n_rf=[30, 70, 65], n_gp=[5, 2, 2], kinds=["RBF", "ARC", "RBF"], input_cat=True, B=77, N=500, mean=True),
    "single_layer": dict(task="cls", d_in=7, d_out=3, L=1, n_rf=20, n_gp=[3], kinds=["RBF"], input_cat=False, B=5, N=50),
    "wide_gp": dict(task="reg", d_in=4, d_out=40, L=2, n_rf=[33, 129], n_gp=[64, 40], kinds=["ARC", "RBF"], input_cat=False, B=64, N=1000),
}


def make_model(name, seed=0, device=None):
    c = CONFIGS[name]
    torch.manual_seed(seed)
    cls = RegressionDGP if c["task"] == "reg" else ClassificationDGP
    model = cls(c["d_in"], c["d_out"], n_hidden_layers=c["L"], n_rf=c["n_rf"], n_gp=c["n_gp"],
                kernel_type_list=c["kinds"], input_cat=c["input_cat"], set_nonzero_mean=c.get("mean", False))
    e = model._engine
    if c.get("mean", False):       # non-trivial means / hypers so every term of the backward is exercised
        for l in range(c["L"]):
            model._vars[f"mean_{l}"].assign(0.3 * torch.randn(e.spec.layers[l].d, 1))
    for l in range(c["L"]):
        model._vars[f"log_amp_{l}"].assign(torch.tensor(0.1 * (l + 1)))
        model._vars[f"log_inv_ls_{l}"].assign(model._vars[f"log_inv_ls_{l}"].tensor.cpu() + 0.2 * torch.randn(e.spec.layers[l].d))
    g = torch.Generator().manual_seed(seed + 1)
    X = torch.randn(c["B"], c["d_in"], generator=g)
    if c["task"] == "reg":
        Y = torch.randn(c["B"], c["d_out"], generator=g)
    else:
        Y = torch.randint(0, c["d_out"], (c["B"], 1), generator=g).float()
    return model, X, Y, c


def oracle_params(model, dtype=torch.float64) -> O.DGPParams:
    e = model._engine
    L = len(e.spec.layers)
    cp = lambda t: t.detach().to("cpu", dtype).clone()
    mean = [cp(e.view(f"mean_{l}")) if e.spec.layers[l].has_mean else torch.zeros(e.spec.layers[l].d, 1, dtype=dtype)
            for l in range(L)]
    return O.DGPParams([s.kind for s in e.spec.layers], [cp(z[0]) for z in e.z],
                       [cp(e.view(f"log_inv_ls_{l}")) for l in range(L)],
                       [cp(e.view(f"log_amp_{l}")) for l in range(L)], mean,
                       [cp(e.view(f"W_{l}")) for l in range(L)],
                       cp(e.view("lik_log_var")) if e.spec.likelihood == "gaussian" else None,
                       model.input_cat, any(s.has_mean for s in e.spec.layers), e.spec.likelihood)


def rel_err(got, ref):
    got = got.detach().double().cpu().reshape(-1)
    ref = ref.detach().double().cpu().reshape(-1)
    scale = max(float(ref.abs().max()), 1e-30)
    return float((got - ref).abs().max()) / scale


def assert_close(got, ref, rtol=1e-4, atol_scale=1e-6, what=""):
    """Element-wise form of the contract (BASELINE.json north_star rtol 1e-4; SURVEY 8c: atol = 1e-6 * max|ref|):
    |got - ref| <= atol_scale * max|ref| + rtol * |ref| for EVERY element, so errors on small entries count."""
    got = got.detach().double().cpu().reshape(-1)
    ref = ref.detach().double().cpu().reshape(-1)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    atol = atol_scale * max(float(ref.abs().max()), 1e-30)
    excess = (got - ref).abs() - (atol + rtol * ref.abs())
    worst = int(excess.argmax())
    assert float(excess[worst]) <= 0.0, (
        f"{what}: element {worst}: got {float(got[worst]):.9g}, ref {float(ref[worst]):.9g}, "
        f"|diff| {abs(float(got[worst] - ref[worst])):.3g} > atol {atol:.3g} + rtol*|ref| {rtol * abs(float(ref[worst])):.3g}")
