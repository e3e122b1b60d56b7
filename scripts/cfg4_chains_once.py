"""configs[3] shape (YearPrediction: d = 90, n_gp = [30, 30, 1], M = 512, B = 1000), 8 chains batched per launch, tf32 mode:
chain-iterations/s and the per-kernel breakdown of one step (A/B runs of the layered path: DGPRF_NO_FUSED_SLAB_SUMS, DGPRF_NO_PDL)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200")); sys.path.insert(0, ROOT)
import torch
from dgprf import _ffi
from dgprf.engine import Engine, ModelSpec

CH = int(os.environ.get("CHAINS", 8))
spec = ModelSpec.build(90, 1, [512] * 3, [30, 30, 1], ["RBF"] * 3, True, False, "gaussian")
torch.manual_seed(0)
e = Engine(spec, CH, precision=_ffi.PREC_TF32)
e.theta_w.normal_()
X = torch.randn(1000, 90, device="cuda"); Y = torch.randn(1000, 1, device="cuda")
for i in range(10):
    e.step(X, Y, 515345.0, 0.01, 0.9, 1.0, False, False, 1, i)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
n = 300
for i in range(n):
    e.step(X, Y, 515345.0, 0.01, 0.9, 1.0, False, False, 1, 10 + i)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
_ffi.profile_start()
e.step(X, Y, 515345.0, 0.01, 0.9, 1.0, False, False, 1, 999)
recs = _ffi.profile_stop()
print(f"chains={CH} tf32: {ms * 1e3:.1f} us/step  {CH * 1e3 / ms:.0f} chain-it/s | " + " ".join(f"{nm}={t * 1e3:.0f}" for nm, t in recs))
