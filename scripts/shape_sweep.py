"""Step time of one W-only sampling step, fp32 vs tf32, over a grid of minibatch sizes and feature counts
(3-layer RBF, 32 inputs, n_gp = [16, 16, 1], input concatenation): a check that the kernel-selection heuristics have no cliffs."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
from dgprf import _ffi
from dgprf.chains import ChainEnsemble
out = {}
for M in (256, 1024, 4096):
    for B in (512, 2048, 8192, 32768):
        row = {}
        for prec in ("fp32", "tf32"):
            torch.manual_seed(0)
            ens = ChainEnsemble(32, 1, 3, M, [16, 16, 1], input_cat=True, n_chains=1, seed=1, precision=prec)
            X = torch.randn(B, 32, device="cuda"); Y = torch.randn(B, 1, device="cuda")
            for _ in range(5): ens.sgmcmc_update(X, Y, 1e5, lr=1e-4, momentum_decay=0.9)
            n = 100 if B * M <= (1 << 23) else 20
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            for _ in range(n): ens.sgmcmc_update(X, Y, 1e5, lr=1e-4, momentum_decay=0.9)
            b.record(); torch.cuda.synchronize()
            row[prec] = a.elapsed_time(b) / n
            _ffi.profile_start(); ens.sgmcmc_update(X, Y, 1e5, lr=1e-4, momentum_decay=0.9)
            row[prec + "_kernels"] = sorted({nm for nm, _ in _ffi.profile_stop()})
            assert torch.isfinite(ens.engine.theta_w).all()
            del ens; torch.cuda.empty_cache()
        out[f"M={M} B={B}"] = row
        print(f"M={M:5d} B={B:6d}  fp32 {row['fp32']:8.3f} ms   tf32 {row['tf32']:8.3f} ms   x{row['fp32'] / row['tf32']:.2f}   {','.join(k for k in row['tf32_kernels'] if k.startswith('k1') or k.startswith('k2') or k.startswith('k9'))}", flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "shape_sweep.json"), "w"), indent=1)
