"""A few sampling steps of one BASELINE config (for ncu captures): python scripts/step_once.py cfg2 [chains] [precision]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
import torch
from dgprf.chains import ChainEnsemble
CFGS = {
    "cfg1": dict(d_in=1, d_out=1, L=2, n_rf=100, n_gp=[1, 1], cat=False, B=20, N=60),
    "cfg2": dict(d_in=9, d_out=1, L=3, n_rf=512, n_gp=[9, 9, 1], cat=True, B=1000, N=45730),
    "cfg4": dict(d_in=90, d_out=1, L=3, n_rf=512, n_gp=[30, 30, 1], cat=True, B=1000, N=515345),
}
c = CFGS[sys.argv[1]]
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
ens = ChainEnsemble(c["d_in"], c["d_out"], c["L"], c["n_rf"], c["n_gp"], input_cat=c["cat"], likelihood="gaussian", n_chains=C, seed=1, precision=prec)
X = torch.randn(c["B"], c["d_in"], device="cuda"); Y = torch.randn(c["B"], 1, device="cuda")
for _ in range(6):
    ens.sgmcmc_update(X, Y, c["N"], lr=1e-3, momentum_decay=0.9)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(ens.engine.theta_w).all()))
