"""Host cost of one end-to-end step, layer by layer: the raw C call, engine.step_host, model.sgmcmc_update, with slicing."""
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
kw = dict(lr=0.01, momentum_decay=0.9)
for i in range(50): model.sgmcmc_update(*batches[i % 45], N, u_host=u, **kw)
torch.cuda.synchronize()
e = model._engine


def run(name, fn, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name:42s} n={n:5d}: cpu enqueue {1e6*(t1-t0)/n:6.1f} us/step, total {1e6*(t2-t0)/n:6.1f} us/step -> {n/(t2-t0):7.0f} it/s", flush=True)


def f_model(i): model.sgmcmc_update(*batches[i % 45], N, u_host=u, **kw)
def f_model_slice(i):
    lo = (i % 45) * B
    model.sgmcmc_update(Xh[lo:lo + B], Yh[lo:lo + B], N, u_host=u, **kw)
def f_engine(i):
    xb, yb = batches[i % 45]
    e.step_host(xb, yb, float(N), 0.01, 0.9, 1.0, False, False, 1234, 100000 + i, u_host=u)
def f_dev(i): model.sgmcmc_update(X[:B], Y[:B], N, **kw)

for n in (200, 3000):
    run("model.sgmcmc_update(host, pre-sliced)", f_model, n)
    run("model.sgmcmc_update(host, slice per step)", f_model_slice, n)
    run("engine.step_host", f_engine, n)
    run("model.sgmcmc_update(device batch)", f_dev, n)
if os.environ.get("PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile()
    torch.cuda.synchronize()
    pr.enable()
    for i in range(200): f_model_slice(i)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
