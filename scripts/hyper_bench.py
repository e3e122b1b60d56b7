"""Hyper-gradient pass (forward + backward with d log_inv_ls / d log_amp / d mean, the stochastic-EM M-step and full-Bayes
gradient) at configs[4] scale, per precision, with the per-kernel breakdown from the library's event hook."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
from dgprf import _ffi
from dgprf.engine import Engine, ModelSpec
spec = ModelSpec.build(90, 1, [4096] * 5, [30, 30, 30, 30, 1], ["RBF"] * 5, True, False, "gaussian")
for prec in (sys.argv[1:] or ["fp32", "tf32"]):
    e = Engine(spec, 1, precision={"fp32": _ffi.PREC_FP32, "tf32": _ffi.PREC_TF32}[prec])
    e.theta_w.normal_(); e.theta_h[:, e.layout.off_lik_log_var] = -2.0
    X = torch.randn(65536, 90, device="cuda"); Y = torch.randn(65536, 1, device="cuda")
    for _ in range(2): e.gradients(X, Y, 1e5, hyper=True, prior_w=True, prior_h=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(5): e.gradients(X, Y, 1e5, hyper=True, prior_w=True, prior_h=True)
    b.record(); torch.cuda.synchronize()
    print(prec, "cfg5 hyper-gradient pass:", round(a.elapsed_time(b) / 5, 2), "ms")
    _ffi.profile_start()
    e.gradients(X, Y, 1e5, hyper=True, prior_w=True, prior_h=True)
    print("   ", " ".join(f"{nm}={t * 1e3:.0f}us" for nm, t in _ffi.profile_stop()))
    del e; torch.cuda.empty_cache()
