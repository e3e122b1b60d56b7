import os, sys, time
sys.path.insert(0, "/root/repo/dgp-rf-mcmc_b200"); sys.path.insert(0, "/root/repo")
import torch
from models.regression_model import RegressionDGP
model = RegressionDGP(9, 1, n_hidden_layers=3, n_rf=512, n_gp=[9, 9, 1], input_cat=True)
X = torch.randn(1000, 9, device="cuda"); Y = torch.randn(1000, 1, device="cuda")
e = model._engine
for fused in (True, False):
    for _ in range(20): e.gradients(X, Y, 45730, hyper=False, prior_w=True, prior_h=False, fused=fused)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(300): e.gradients(X, Y, 45730, hyper=False, prior_w=True, prior_h=False, fused=fused)
    b.record(); torch.cuda.synchronize()
    print(f"W-only gradient pass, configs[1] shape, fused={fused}: {a.elapsed_time(b)/300*1e3:.1f} us")
ds = [(X, Y)] * 32
for _ in range(3): model.precond_update(ds, 45730, K_batches=32, precond_type="rmsprop", second_moment_centered=False)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): model.precond_update(ds, 45730, K_batches=32, precond_type="rmsprop", second_moment_centered=False)
torch.cuda.synchronize(); print(f"precond_update rmsprop K=32: {(time.perf_counter()-t0)/10*1e3:.2f} ms")
