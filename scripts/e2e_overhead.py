"""Where the end-to-end step time goes: CPU enqueue cost of model.sgmcmc_update on host minibatches vs GPU time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200")); sys.path.insert(0, ROOT)
import torch
import bench
from models.regression_model import RegressionDGP
CFG = bench.CFG
X, Y = bench.synthetic_protein(0, torch.device("cuda"))
Xh, Yh = X.cpu().pin_memory(), Y.cpu().pin_memory()
model = RegressionDGP(CFG["D"], 1, n_hidden_layers=3, n_rf=512, n_gp=[9, 9, 1], input_cat=True)
model.precond_update(None, CFG["N"], precond_type="identity")
B, N = 1000, CFG["N"]
u = torch.zeros(1).pin_memory()
batches = [(Xh[i * B:(i + 1) * B], Yh[i * B:(i + 1) * B]) for i in range(45)]
for i in range(50): model.sgmcmc_update(*batches[i % 45], N, lr=0.01, momentum_decay=0.9, u_host=u)
torch.cuda.synchronize()
n = 3000
t0 = time.perf_counter()
for i in range(n): model.sgmcmc_update(*batches[i % 45], N, lr=0.01, momentum_decay=0.9, u_host=u)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"pre-sliced batches: cpu enqueue {1e6*(t1-t0)/n:.1f} us/step, total {1e6*(t2-t0)/n:.1f} us/step -> {n/(t2-t0):.0f} it/s")
t0 = time.perf_counter()
for i in range(n):
    lo = (i % 45) * B
    model.sgmcmc_update(Xh[lo:lo + B], Yh[lo:lo + B], N, lr=0.01, momentum_decay=0.9, u_host=u)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"slicing per step:   cpu enqueue {1e6*(t1-t0)/n:.1f} us/step, total {1e6*(t2-t0)/n:.1f} us/step -> {n/(t2-t0):.0f} it/s")
# short bursts: the launch queue never fills, so the enqueue time is pure host cost (Python + ctypes + libdgprf + driver launch)
for burst in (100, 200, 400):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(burst): model.sgmcmc_update(*batches[i % 45], N, lr=0.01, momentum_decay=0.9, u_host=u)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"burst of {burst}: cpu enqueue {1e6*(t1-t0)/burst:.1f} us/step, total {1e6*(t2-t0)/burst:.1f} us/step")
Xd, Yd = X, Y
dev_batches = [(Xd[i * B:(i + 1) * B], Yd[i * B:(i + 1) * B]) for i in range(45)]
for i in range(50): model.sgmcmc_update(*dev_batches[i % 45], N, lr=0.01, momentum_decay=0.9)
torch.cuda.synchronize()
for burst in (200, 3000):
    t0 = time.perf_counter()
    for i in range(burst): model.sgmcmc_update(*dev_batches[i % 45], N, lr=0.01, momentum_decay=0.9)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"DEVICE batches, burst of {burst}: cpu enqueue {1e6*(t1-t0)/burst:.1f} us/step, total {1e6*(t2-t0)/burst:.1f} us/step")
