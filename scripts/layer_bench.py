"""One fused [RF -> GP] layer at a given shape: forward (saving Phi) and backward, SIMT fp32 vs tcgen05 tf32.
Reports time, algorithmic TFLOP/s and the Phi-store GB/s (SURVEY section 8(d) work formulas)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
import torch  # noqa: E402
from dgprf import _ffi  # noqa: E402
from dgprf.engine import Engine, ModelSpec  # noqa: E402

B, d, M, g = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (65536, 120, 4096, 30)))
kind = sys.argv[5] if len(sys.argv) > 5 else "RBF"
spec = ModelSpec.build(d, g, [M], [g], [kind], False, False, "gaussian")
X = torch.randn(B, d, device="cuda")
Y = torch.randn(B, g, device="cuda")
F = 2 * M if kind == "RBF" else M
fwd_flops = 2.0 * B * (d * M + F * g)
bwd_flops = 4.0 * B * F * g
out = {"shape": dict(B=B, d=d, M=M, g=g, kind=kind)}
for prec in ("fp32", "tf32"):
    e = Engine(spec, 1, precision={"fp32": _ffi.PREC_FP32, "tf32": _ffi.PREC_TF32}[prec])
    e.theta_w.normal_()
    e.theta_h[:, e.layout.off_lik_log_var] = -2.0
    for mode, tag in ((_ffi.MODE_EVAL, "fwd_eval"), (_ffi.MODE_TRAIN, "fwd_train")):
        for _ in range(3):
            e.forward(X, mode=mode, want_F=False)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        n = 10
        for _ in range(n):
            e.forward(X, mode=mode, want_F=False)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        out[f"{prec}_{tag}"] = {"ms": ms, "tflops": fwd_flops / ms / 1e9,
                                "phi_store_gbs": (4.0 * B * F / ms / 1e6) if mode == _ffi.MODE_TRAIN else None}
    # full gradient pass (fwd + lik + bwd + finalize)
    for _ in range(2):
        e.gradients(X, Y, 1e5, hyper=False, prior_w=True, prior_h=False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        e.gradients(X, Y, 1e5, hyper=False, prior_w=True, prior_h=False)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    out[f"{prec}_fwd_bwd"] = {"ms": ms, "tflops": (fwd_flops + bwd_flops) / ms / 1e9}
    # per-kernel times of one gradient pass (library event hook)
    _ffi.profile_start()
    e.gradients(X, Y, 1e5, hyper=False, prior_w=True, prior_h=False)
    out[f"{prec}_kernels_ms"] = {nm: round(ms_, 4) for nm, ms_ in _ffi.profile_stop()}
    del e
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
