"""Iterations/s of one sampling step at the shapes of all five BASELINE.json configs (synthetic data), per precision.
cfg5 is run on ONE GPU with the whole 65536-row minibatch (its data-parallel split over GPUs divides the rows)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
import torch
from dgprf.chains import ChainEnsemble

CFGS = {
    "cfg1 sin demo (2-layer RBF, M=100, B=20)": dict(d_in=1, d_out=1, L=2, n_rf=100, n_gp=[1, 1], kinds=None, cat=False, lik="gaussian", B=20, N=60, C=1, beta=0.95),
    "cfg2 protein (3-layer RBF, M=512, B=1000)": dict(d_in=9, d_out=1, L=3, n_rf=512, n_gp=[9, 9, 1], kinds=None, cat=True, lik="gaussian", B=1000, N=45730, C=1, beta=0.9),
    "cfg3 MNIST (3-layer ARC softmax, M=512, B=2048, SGLD)": dict(d_in=784, d_out=10, L=3, n_rf=512, n_gp=[30, 30, 10], kinds=["ARC"] * 3, cat=True, lik="softmax", B=2048, N=60000, C=1, beta=0.0),
    "cfg4 YearPrediction (3-layer RBF, M=512, B=1000, 8 chains/GPU)": dict(d_in=90, d_out=1, L=3, n_rf=512, n_gp=[30, 30, 1], kinds=None, cat=True, lik="gaussian", B=1000, N=515345, C=8, beta=0.9),
    "cfg5 (5-layer RBF, M=4096, B=65536, 1 GPU)": dict(d_in=90, d_out=1, L=5, n_rf=4096, n_gp=[30, 30, 30, 30, 1], kinds=None, cat=True, lik="gaussian", B=65536, N=515345, C=1, beta=0.9),
}


def flops(c):
    d = [c["d_in"]] + [g + (c["d_in"] if c["cat"] else 0) for g in c["n_gp"][:-1]]
    tot = 0.0
    for l in range(c["L"]):
        M = c["n_rf"]; g = c["n_gp"][l]
        F = M if (c["kinds"] and c["kinds"][l] == "ARC") else 2 * M
        tot += 2.0 * c["B"] * (d[l] * M + F * g) + 4.0 * c["B"] * F * g + (2.0 * c["B"] * c["n_gp"][l - 1] * M if l > 0 else 0.0)
    return tot


only = sys.argv[1] if len(sys.argv) > 1 else None
out = {}
for name, c in CFGS.items():
    if only and only not in name:
        continue
    for prec in ("fp32", "tf32"):
        torch.manual_seed(0)
        ens = ChainEnsemble(c["d_in"], c["d_out"], c["L"], c["n_rf"], c["n_gp"], kernel_type_list=c["kinds"], input_cat=c["cat"],
                            likelihood=c["lik"], n_chains=c["C"], seed=1, precision=prec)
        X = torch.randn(c["B"], c["d_in"], device="cuda")
        Y = torch.randn(c["B"], 1, device="cuda") if c["lik"] == "gaussian" else torch.randint(0, c["d_out"], (c["B"], 1), device="cuda").float()
        kw = dict(lr=1e-3, momentum_decay=c["beta"])
        n_warm, n = (3, 10) if c["B"] > 10000 else (20, 300)
        for _ in range(n_warm):
            ens.sgmcmc_update(X, Y, c["N"], **kw)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(n):
            ens.sgmcmc_update(X, Y, c["N"], **kw)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        assert torch.isfinite(ens.engine.theta_w).all()
        if os.environ.get("DGPRF_BREAKDOWN"):          # per-kernel GPU times of one step (C-ABI profile hook)
            from dgprf import _ffi
            _ffi.profile_start()
            ens.sgmcmc_update(X, Y, c["N"], **kw)
            for nm, t in _ffi.profile_stop():
                print(f"    {nm:24s} {t * 1e3:9.1f} us")
        out[f"{name} [{prec}]"] = {"ms_per_step": round(ms, 4), "chain_it_per_s": round(c["C"] * 1e3 / ms, 1),
                                   "algorithmic_TFLOPs": round(c["C"] * flops(c) / ms / 1e9, 2), "GFLOP_per_chain_step": round(flops(c) / 1e9, 3)}
        print(f"{name} [{prec}]: {ms:.4f} ms/step  {c['C'] * 1e3 / ms:.1f} chain-it/s  {c['C'] * flops(c) / ms / 1e9:.2f} TFLOP/s", flush=True)
        del ens
        torch.cuda.empty_cache()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config_bench.json"), "w"), indent=1)
