import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dgp-rf-mcmc_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import dgprf_oracle as O
from helpers import oracle_params, rel_err
from models.regression_model import RegressionDGP
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 200
M = int(sys.argv[2]) if len(sys.argv) > 2 else 128
model = RegressionDGP(9, 1, n_hidden_layers=3, n_rf=M, n_gp=[9, 9, 1], input_cat=True)
X = torch.randn(B, 9); Y = torch.randn(B, 1)
p = oracle_params(model)
u_ref, g_ref = O.grads_autograd(p, X.double(), Y.double(), 5000, False)
u32, g32 = model.grad_U(X, Y, 5000)
model.set_precision("tf32")
u, g = model.grad_U(X, Y, 5000)
for n in g_ref:
    r = g_ref[n]; a = g[n].double().cpu(); b = g32[n].double().cpu()
    print(n, "tf32 err", rel_err(a, r), "fp32 err", rel_err(b, r), "norms", float(r.abs().max()), float(a.abs().max()))
    if rel_err(a, r) > 1e-2:
        d = (a - r).abs()
        bad_rows = (d.max(1).values > 1e-2 * r.abs().max()).nonzero().flatten()
        print("   bad rows:", bad_rows[:20].tolist(), "...", len(bad_rows), "of", r.shape[0])
        print("   ref[0,:4]", r[0, :4].tolist(), " got", a[0, :4].tolist())
