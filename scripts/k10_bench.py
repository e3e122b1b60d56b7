"""A/B of the two row-fused step kernels at BASELINE shapes: K9 (one CTA per 8 rows, FFMA) vs K10 (cluster-split, 3xTF32
mma.sync), per-step CUDA-event time of `sgmcmc_update`.  Environment switches are read by the library per call, so one
process can time every variant: DGPRF_NO_K10, DGPRF_K10_MT, DGPRF_K10_CL."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
import torch
from dgprf.chains import ChainEnsemble
from dgprf import _ffi

CFGS = {
    "cfg1": dict(d_in=1, d_out=1, L=2, n_rf=100, n_gp=[1, 1], cat=False, B=20, N=60, beta=0.95),
    "cfg2": dict(d_in=9, d_out=1, L=3, n_rf=512, n_gp=[9, 9, 1], cat=True, B=1000, N=45730, beta=0.9),
    "cfg4": dict(d_in=90, d_out=1, L=3, n_rf=512, n_gp=[30, 30, 1], cat=True, B=1000, N=515345, beta=0.9),
}
VARIANTS = [("k9", {"DGPRF_NO_K10": "1"}), ("k10 auto", {}), ("k10 MT2 CL4", {"DGPRF_K10_MT": "2", "DGPRF_K10_CL": "4"}),
            ("k10 MT1 CL2", {"DGPRF_K10_MT": "1", "DGPRF_K10_CL": "2"}), ("k10 MT2 CL2", {"DGPRF_K10_MT": "2", "DGPRF_K10_CL": "2"}),
            ("k10 MT1 CL4", {"DGPRF_K10_MT": "1", "DGPRF_K10_CL": "4"}), ("k10 MT1 CL1", {"DGPRF_K10_MT": "1", "DGPRF_K10_CL": "1"}),
            ("k10 MT2 CL8", {"DGPRF_K10_MT": "2", "DGPRF_K10_CL": "8"})]
ENVS = ["DGPRF_NO_K10", "DGPRF_K10_MT", "DGPRF_K10_CL"]
out = {}
chains = [int(c) for c in os.environ.get("K10_CHAINS", "1,8").split(",")]
only_cfg = os.environ.get("K10_ONLY")
only_var = os.environ.get("K10_VARIANT")
for name, c in CFGS.items():
    if only_cfg and name != only_cfg:
        continue
    for C in chains:
        for vname, env in VARIANTS:
            if only_var and vname != only_var:
                continue
            for k in ENVS:
                os.environ.pop(k, None)
            os.environ.update(env)
            torch.manual_seed(0)
            try:
                ens = ChainEnsemble(c["d_in"], c["d_out"], c["L"], c["n_rf"], c["n_gp"], input_cat=c["cat"], likelihood="gaussian",
                                    n_chains=C, seed=1, precision="fp32")
                X = torch.randn(c["B"], c["d_in"], device="cuda"); Y = torch.randn(c["B"], 1, device="cuda")
                kw = dict(lr=1e-3, momentum_decay=c["beta"])
                for _ in range(20):
                    ens.sgmcmc_update(X, Y, c["N"], **kw)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 300
                torch.cuda.synchronize(); a.record()
                for _ in range(n):
                    ens.sgmcmc_update(X, Y, c["N"], **kw)
                b.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b) / n
                _ffi.profile_start()
                ens.sgmcmc_update(X, Y, c["N"], **kw)
                ker = [(nm, round(t * 1e3, 1)) for nm, t in _ffi.profile_stop()]
                ok = bool(torch.isfinite(ens.engine.theta_w).all())
                print(f"{name} C={C} {vname:12s}: {ms * 1e3:8.1f} us/step  {C * 1e3 / ms:10.0f} chain-it/s  finite={ok}  {ker}", flush=True)
                out[f"{name} C={C} {vname}"] = {"us_per_step": round(ms * 1e3, 2), "chain_it_per_s": round(C * 1e3 / ms), "kernels": ker}
                del ens
            except Exception as e:
                print(f"{name} C={C} {vname}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
            torch.cuda.empty_cache()
if not only_var:
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "k10_bench.json"), "w"), indent=1)
