"""Full-Bayesian sampling step (models/dgp.py:184-216 with full_bayesian=True: W AND kernel / likelihood hyper-parameters are
sampled) at the configs[1] shape: the layered kernels in hyper mode + two update launches.  fp32 and tf32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200")); sys.path.insert(0, ROOT)
import torch
from dgprf import _ffi
from models.regression_model import RegressionDGP
B, N = 1000, 45730
for prec in ("fp32", "tf32"):
    torch.manual_seed(0)
    model = RegressionDGP(9, 1, n_hidden_layers=3, n_rf=512, n_gp=[9, 9, 1], input_cat=True)
    model.set_precision(prec)
    model.precond_update(None, N, precond_type="identity", full_bayesian=True)
    X = torch.randn(B, 9, device="cuda"); Y = torch.randn(B, 1, device="cuda")
    kw = dict(lr=1e-3, momentum_decay=0.9, full_bayesian=True)
    for _ in range(20): model.sgmcmc_update(X, Y, N, **kw)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    n = 300
    for _ in range(n): model.sgmcmc_update(X, Y, N, **kw)
    b.record(); torch.cuda.synchronize()
    _ffi.profile_start()
    model.sgmcmc_update(X, Y, N, **kw)
    ker = [(nm, round(t * 1e3, 1)) for nm, t in _ffi.profile_stop()]
    print(f"full-Bayes step, configs[1] shape, {prec}: {a.elapsed_time(b) / n * 1e3:.1f} us/step; kernels {ker}", flush=True)
