"""True per-kernel GPU durations of one sampling step: the launches are queued behind a GPU
backlog so the CUDA events around each kernel do not include CPU launch latency."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from dgprf import _ffi  # noqa: E402
from models.regression_model import RegressionDGP  # noqa: E402

CFG = bench.CFG
dev = torch.device("cuda", 0)
X, Y = bench.synthetic_protein(0, dev)
model = RegressionDGP(CFG["D"], 1, n_hidden_layers=CFG["L"], n_rf=CFG["n_rf"], n_gp=CFG["n_gp"], input_cat=True)
model.set_precision(os.environ.get("DGPRF_PRECISION", "fp32"))
model.precond_update(None, CFG["N"], precond_type="identity")
B, N = CFG["batch"], CFG["N"]
for i in range(20):
    model.sgmcmc_update(X[:B], Y[:B], N)
torch.cuda.synchronize()
backlog = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for _ in range(60):
    backlog.fill_(1)                      # ~60 x 0.15 ms of GPU work to hide the CPU enqueue time
_ffi.profile_start()
for i in range(STEPS):
    lo = (i % 45) * B
    model.sgmcmc_update(X[lo:lo + B], Y[lo:lo + B], N)
recs = _ffi.profile_stop()
per = {}
for idx, (nm, ms) in enumerate(recs):
    per.setdefault((idx % 8, nm), []).append(ms * 1e3)
tot = 0.0
for (slot, nm), v in sorted(per.items()):
    print(f"slot {slot} {nm:20s} median {statistics.median(v):7.2f} us  min {min(v):7.2f}")
    tot += statistics.median(v)
print(f"sum of kernel medians: {tot:.1f} us/step")
# whole-step GPU time behind a backlog (no per-kernel events)
for _ in range(60):
    backlog.fill_(2)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(STEPS):
    lo = (i % 45) * B
    model.sgmcmc_update(X[lo:lo + B], Y[lo:lo + B], N)
b.record()
torch.cuda.synchronize()
print(f"GPU time per step behind a backlog: {a.elapsed_time(b) * 1e3 / STEPS:.1f} us")
import time
t0 = time.perf_counter()
for i in range(2000):
    lo = (i % 45) * B
    model.sgmcmc_update(X[lo:lo + B], Y[lo:lo + B], N)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"CPU enqueue time per step (python + 8 launches): {(t1 - t0) / 2000 * 1e6:.1f} us")
