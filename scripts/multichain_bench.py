"""Chain-iterations/s of C independent chains batched per launch (protein shape), per precision."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200")); sys.path.insert(0, ROOT)
import torch
import bench
from dgprf.chains import ChainEnsemble
CFG = bench.CFG
X, Y = bench.synthetic_protein(0, torch.device("cuda"))
B, N = CFG["batch"], CFG["N"]
for prec in ("fp32", "tf32"):
    for CH in (1, 2, 4, 8, 16, 32):
        ens = ChainEnsemble(CFG["D"], 1, CFG["L"], CFG["n_rf"], CFG["n_gp"], input_cat=True, n_chains=CH, seed=7, precision=prec)
        for i in range(5):
            ens.sgmcmc_update(X[:B], Y[:B], N, lr=0.01, momentum_decay=0.9)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        n = 200
        for i in range(n):
            ens.sgmcmc_update(X[:B], Y[:B], N, lr=0.01, momentum_decay=0.9)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        print(f"{prec} chains={CH:3d}  {ms*1e3:8.1f} us/step  {CH*1e3/ms:10.0f} chain-it/s")
        del ens
