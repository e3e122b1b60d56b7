#!/usr/bin/env python
"""Benchmark of the SG-MCMC sampling hot path (BASELINE.json: "SGHMC iters/sec ... 3-layer RF-DGP").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1] -- 3-layer RBF RF-DGP regression on the synthetic
UCI-protein shape (N=45730, D=9), M=512 random features, n_gp=[9,9,1], input concatenation, batch 1000,
SGHMC (beta=0.9, T=1, lr=0.01), one chain per GPU.  A step is one `sgmcmc_update` on one minibatch:
fused RF->GP forward x3, likelihood seed, backward x3, fused Philox update.

  value   iterations/s with the dataset resident in HBM; every timed step is bracketed by its own CUDA
          events on the launch stream and L2 is flushed (256 MiB write) between timed steps.
  e2e     iterations/s through the public drop-in call `model.sgmcmc_update(X_batch, Y_batch, ...)` with
          the minibatch in pinned HOST memory: H2D copy of the batch and D2H read of the minibatch
          log-likelihood inside the timed region, back to back (no flush), barrier+sync on both sides.
  N > 1   one process per GPU, independent chains (chain id = rank), no data-path collective: weak scaling.
  --impl reference   the CPU restatement of the reference (oracle/, fp32 torch eager, all host threads):
          TensorFlow cannot be installed offline, so the reference itself cannot run here.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "dgp-rf-mcmc_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

CFG = dict(workload="configs[1]: 3-layer RBF RF-DGP regression, synthetic UCI-protein shape",
           N=45730, D=9, L=3, n_rf=512, n_gp=[9, 9, 1], input_cat=True, batch=1000,
           sampler="SGHMC", lr=0.01, momentum_decay=0.9, temperature=1.0, chains_per_gpu=1)
METRIC = "sghmc_iterations_per_second"
UNIT = "it/s"


def synthetic_protein(rank, device):
    """X ~ N(0,1) [N,D]; Y = sin(X w) + 0.1 eps, standardised."""
    g = torch.Generator().manual_seed(1234 + rank)
    X = torch.randn(CFG["N"], CFG["D"], generator=g)
    w = torch.randn(CFG["D"], 1, generator=g)
    Y = torch.sin(X @ w) + 0.1 * torch.randn(CFG["N"], 1, generator=g)
    Y = (Y - Y.mean()) / Y.std()
    return X.to(device), Y.to(device)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's fp32 eager step (op-for-op restatement of the reference's TF eager path)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, budget_s):
    import dgprf_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1234)
    X = torch.randn(CFG["N"], CFG["D"], generator=g)
    Y = torch.randn(CFG["N"], 1, generator=g)
    p = O.init_params(CFG["D"], 1, CFG["L"], CFG["n_rf"], CFG["n_gp"], None, CFG["input_cat"], "gaussian",
                      seed=1, dtype=torch.float32)
    mom = {n: torch.randn(t.shape, generator=g) for n, t in p.w_named()}
    B = CFG["batch"]
    nb = CFG["N"] // B

    def one(i):
        lo = (i % nb) * B
        O.eager_step_fp32(p, mom, X[lo:lo + B], Y[lo:lo + B], CFG["N"], CFG["lr"], CFG["momentum_decay"],
                          CFG["temperature"], gen=g)
    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    done = 0
    while done < steps and (time.perf_counter() - t0) < budget_s:
        one(warmup + done)
        done += 1
    dt = time.perf_counter() - t0
    return done / dt, done, dt, threads


def run_reference(args, rank):
    if rank != 0:
        return
    its, done, dt, threads = cpu_reference_run(args.steps, max(args.warmup, 3), budget_s=60.0)
    sample = f"{done} consecutive minibatch steps of the same workload in {dt:.1f} s (cap 60 s)"
    line = {"impl": "reference", "metric": METRIC, "value": its, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 / its, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CFG,
            "cpu_baseline": {"value": its, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "oracle/dgprf_oracle.py eager fp32 restatement; TensorFlow is not installable offline"},
            "e2e": {"value": its, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def algorithmic_flops(spec, B):
    """fwd = sum_l 2B(d_l M_l + F_l g_l); bwd_W = sum_l 4B F_l g_l + sum_{l>=1} 2B g_{l-1} M_l."""
    fwd = [2.0 * B * (s.d * s.M + s.F * s.g) for s in spec.layers]
    bwd = [4.0 * B * s.F * s.g + (2.0 * B * s.d_prev * s.M if l > 0 else 0.0) for l, s in enumerate(spec.layers)]
    return fwd, bwd


# ncu --set full captures of the dominant kernel on THIS workload (profiles/r02_ncu_summary.md): DRAM bytes per launch
NCU_TRAFFIC = {"k10_": (290816.0, "profiles/r02_ncu_k10_final_raw.csv"), "k9_": (408576.0, "profiles/r01_ncu_full_summary.md section E")}

CFG4 = dict(workload="configs[3]: 3-layer RBF RF-DGP on the synthetic YearPrediction shape, 8 chains per GPU batched per launch, "
                     "sharded predictive averaging", N=515345, D=90, L=3, n_rf=512, n_gp=[30, 30, 1], input_cat=True,
            batch=1000, chains_per_gpu=8, N_test=51535, sampler="SGHMC", lr=0.01, momentum_decay=0.9)
CFG5 = dict(workload="configs[4]: 5-layer RBF RF-DGP, M=4096, global minibatch 65536 split by rows over the GPUs (data parallel), "
                     "one gradient all-reduce per step", N=515345, D=90, L=5, n_rf=4096, n_gp=[30, 30, 30, 30, 1],
            input_cat=True, global_batch=65536, precision="tf32", lr=1e-4, momentum_decay=0.9)


def measure_tf32_peak(dev):
    """cuBLAS TF32 matmul 8192^3 (2 N^3 flops), best of 10 -- measured the way MEASURED_PEAKS.json measured bf16."""
    n = 8192
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(); torch.matmul(a, b, out=c); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b, c
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def run_ours(args, rank, world):
    local = int(os.environ.get("LOCAL_RANK", 0))
    # ---- CPU baseline first (N = 1 only): nothing else of this job is running, no rank waits in a collective ----
    cpu_base = None
    if world == 1:
        cpu_its, cpu_done, cpu_dt, threads = cpu_reference_run(10 ** 9, 10, budget_s=15.0)
        cpu_base = {"value": cpu_its, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": f"{cpu_done} minibatch steps of the same workload in {cpu_dt:.1f} s"}
        torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))

    from dgprf import _ffi, dist as D
    from dgprf.chains import ChainEnsemble
    from dgprf.engine import Engine, ModelSpec
    from experiments.utils_training import predictive_average
    from models.regression_model import RegressionDGP

    dist = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist          # NCCL_DEBUG is the caller's: its log lines are not the JSON line
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(100 + rank)
    clocks = ClockSampler(local) if rank == 0 else None       # sampled over the whole GPU part of the run
    X, Y = synthetic_protein(rank, dev)
    model = RegressionDGP(CFG["D"], 1, n_hidden_layers=CFG["L"], n_rf=CFG["n_rf"], n_gp=CFG["n_gp"],
                          input_cat=CFG["input_cat"])
    model.set_precision(args.precision)
    model.seed(2024 + rank)
    model.precond_update(None, CFG["N"], precond_type="identity")
    e = model._engine
    B, N = CFG["batch"], CFG["N"]
    nb = N // B
    kw = dict(lr=CFG["lr"], momentum_decay=CFG["momentum_decay"], temperature=CFG["temperature"])
    stream = torch.cuda.current_stream()

    def step(i):
        lo = (i % nb) * B
        model.sgmcmc_update(X[lo:lo + B], Y[lo:lo + B], N, **kw)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def timed(fn, n, warm):
        """n calls of fn(i) between two events on the launch stream, barrier + synchronize on both sides -> seconds."""
        for i in range(warm):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(n):
            fn(warm + i)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1) * 1e-3

    W, K = max(args.warmup, 3), args.steps
    for i in range(W):
        step(i)
    torch.cuda.synchronize()

    # ---- (1) value: resident inputs, per-step events, L2 flushed between timed steps -------------
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    t_wall0 = time.perf_counter()
    # keep the GPU busy (~1 ms) while the host enqueues the first timed steps: every step is bracketed by its own events, so a
    # late launch (host jitter, worst rank of N) would otherwise be counted as device time of that step
    torch.cuda._sleep(2_000_000)
    for i in range(K):
        flush.fill_(i & 0xFF)
        ev[i][0].record(stream)
        step(W + i)
        ev[i][1].record(stream)
    barrier()
    wall_flushed = time.perf_counter() - t_wall0
    t_dev = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    del flush

    # ---- (1b) warm, back-to-back loop (the sampler's real regime: state stays L2 resident) -------
    t_warm = timed(lambda i: step(W + K + i), K, 0)

    # ---- (2) e2e: host (pinned) minibatches through the public call, H2D + D2H in the timed region
    Xh, Yh = X.cpu().pin_memory(), Y.cpu().pin_memory()
    u_host = torch.zeros(1).pin_memory()

    def e2e_step(i):
        # the public drop-in call on a HOST minibatch: H2D copies of X_b, Y_b, the step, and the D2H write of the
        # minibatch log-likelihood are all issued inside the timed region (dgprf_sgmcmc_step_host, pipelined staging)
        lo = (i % nb) * B
        model.sgmcmc_update(Xh[lo:lo + B], Yh[lo:lo + B], N, u_host=u_host, **kw)

    t_e2e = timed(e2e_step, K, W)
    assert math.isfinite(float(u_host[0])), "sampler diverged"
    t_dev, t_warm, t_e2e = max_over_ranks([t_dev, t_warm, t_e2e])

    # ---- (3) configs[3]: 8 chains per GPU batched per launch + sharded predictive combine (all ranks) ----------
    c4 = CFG4
    CH = c4["chains_per_gpu"]
    g4 = torch.Generator().manual_seed(4321)                   # the test set is the same on every rank
    Xt = torch.randn(c4["N_test"], c4["D"], generator=g4).to(dev)
    Yt = torch.randn(c4["N_test"], 1, generator=g4).to(dev)
    Xb = torch.randn(8 * c4["batch"], c4["D"], device=dev); Yb = torch.randn(8 * c4["batch"], 1, device=dev)
    cfg4_out = {"config": c4, "n_gpus": world, "chains_total": CH * world, "scaling": "weak"}
    for prec in ("fp32", "tf32"):
        ens = ChainEnsemble(c4["D"], 1, c4["L"], c4["n_rf"], c4["n_gp"], input_cat=True, n_chains=CH, chain_base=CH * rank,
                            seed=7, precision=prec)
        B4 = c4["batch"]
        n4 = max(20, min(K, 200))
        t4 = timed(lambda i: ens.sgmcmc_update(Xb[(i % 8) * B4:(i % 8 + 1) * B4], Yb[(i % 8) * B4:(i % 8 + 1) * B4], c4["N"],
                                               lr=c4["lr"], momentum_decay=c4["momentum_decay"]), n4, 5)
        (t4,) = max_over_ranks([t4])
        cfg4_out[prec] = {"value": CH * world * n4 / t4, "unit": "chain-iterations/s", "ms_per_step_all_chains": 1e3 * t4 / n4,
                          "steps": n4}
        if prec == "tf32":
            # predictive averaging: every rank evaluates ITS chains (= its sample set) on the full test set and reduces
            # them to per-point log-sum-exps; one all-gather of [G, N_test] + logsumexp of logsumexps (dgprf/dist.py)
            ll, se = ens.evaluate(Xt, Yt)
            ll = ll.as_subclass(torch.Tensor).contiguous(); se = se.as_subclass(torch.Tensor).contiguous()

            def combine(_i):
                _, _, lse = predictive_average(ll, se, aux_is_se=True, n_total_samples=CH * world, return_lse=True)
                return D.combine_predictive(lse, CH, aux_sum_local=float(se.sum()), aux_is_se=True)
            t_comb = timed(combine, 10, 2) / 10
            lp_sharded, rmse_sharded = combine(0)
            (t_comb,) = max_over_ranks([t_comb])
            # check against ONE reduce over the union of all ranks' samples
            if dist is not None:
                all_ll = [torch.empty_like(ll) for _ in range(world)]; all_se = [torch.empty_like(se) for _ in range(world)]
                dist.all_gather(all_ll, ll); dist.all_gather(all_se, se)
                ll_u, se_u = torch.cat(all_ll, 0), torch.cat(all_se, 0)
            else:
                ll_u, se_u = ll, se
            lp_one, rmse_one = predictive_average(ll_u, se_u, aux_is_se=True)
            cfg4_out["predictive_combine"] = {
                "samples_total": CH * world, "n_test": c4["N_test"], "ms": 1e3 * t_comb,
                "bytes_gathered_per_rank": 4 * c4["N_test"] * world,
                "test_log_lik": lp_sharded, "test_rmse": rmse_sharded,
                "single_reduce_log_lik": lp_one, "single_reduce_rmse": rmse_one,
                "matches_single_reduce": bool(abs(lp_sharded - lp_one) <= 1e-5 * max(1.0, abs(lp_one)) and
                                              abs(rmse_sharded - rmse_one) <= 1e-5 * max(1.0, abs(rmse_one))),
                "note": "per-rank K7 reduce (column log-sum-exps) + NCCL all-gather + logsumexp of logsumexps; timed with the K7 launch"}
            del ll, se, ll_u, se_u
        del ens
        torch.cuda.empty_cache()
    del Xt, Yt, Xb, Yb

    # ---- (4) configs[4]: data-parallel step, global minibatch split by rows, one gradient all-reduce (all ranks) ----
    c5 = CFG5
    spec5 = ModelSpec.build(c5["D"], 1, [c5["n_rf"]] * c5["L"], c5["n_gp"], ["RBF"] * c5["L"], True, False, "gaussian")
    torch.manual_seed(0)                                      # identical initial replicas on every rank
    e5 = Engine(spec5, 1, device=dev, precision=_ffi.PREC_TF32)
    e5.theta_w.normal_()
    e5.mom_w.normal_()
    e5.theta_h[:, e5.layout.off_lik_log_var] = -2.0
    g5 = torch.Generator(device=dev).manual_seed(1)
    X5 = torch.randn(c5["global_batch"], c5["D"], device=dev, generator=g5)
    Y5 = torch.randn(c5["global_batch"], 1, device=dev, generator=g5)
    X5l, Y5l = D.row_shard(X5, Y5, rank, world)
    X5l, Y5l = X5l.contiguous(), Y5l.contiguous()
    del X5, Y5
    kw5 = dict(global_rows=c5["global_batch"], data_size=c5["N"], lr=c5["lr"], momentum_decay=c5["momentum_decay"], seed=3)
    n5 = max(5, min(K, 20))
    t5 = timed(lambda i: D.data_parallel_step(e5, X5l, Y5l, step=i, **kw5), n5, 3)
    flat5 = torch.zeros(e5.layout.w_len + 1, device=dev)
    t_ar = timed(lambda i: (dist.all_reduce(flat5) if dist is not None else None), 20, 3) / 20 if dist is not None else 0.0
    # the step's own reduction: the library's two-shot all-reduce over NVLink peer memory (csrc/k11_peer_allreduce.cu) when
    # symmetric memory could be set up on every rank, else the NCCL all-reduce timed above
    peer5 = getattr(e5, "_dp_peer", (None, None))[1]
    t_peer = timed(lambda i: peer5(), 20, 3) / 20 if peer5 is not None else 0.0
    peer_status = peer5.status() if peer5 is not None else None
    t5, t_ar, t_peer = max_over_ranks([t5, t_ar, t_peer])
    chk = e5.theta_w.double().sum().reshape(1)
    same = True
    if dist is not None:
        lo5, hi5 = chk.clone(), chk.clone()
        dist.all_reduce(lo5, op=dist.ReduceOp.MIN); dist.all_reduce(hi5, op=dist.ReduceOp.MAX)
        same = bool(lo5.item() == hi5.item())
    fwd5, bwd5 = algorithmic_flops(spec5, c5["global_batch"])
    cfg5_out = {"config": c5, "n_gpus": world, "rows_per_gpu": int(X5l.shape[0]), "scaling": "strong",
                "ms_per_step": 1e3 * t5 / n5, "value": n5 / t5, "unit": "it/s", "steps": n5,
                "allreduce_us": 1e6 * t_ar, "allreduce_bytes": 4 * (e5.layout.w_len + 1),
                "peer_allreduce_us": 1e6 * t_peer if peer5 is not None else None, "peer_status": peer_status,
                "reduction": ("n/a (one GPU)" if dist is None else
                              "peer: two-shot all-reduce over NVLink peer memory by the library's own kernels (k11_signal_wait / "
                              "k11_reduce_push: every rank sums its 1/world slice of all ranks' gradients in rank order and pushes "
                              "the result to every rank); allreduce_us is NCCL's all-reduce of the same buffer, for comparison"
                              if peer5 is not None else
                              "nccl: one all-reduce of the flat [gW | sum ll] buffer (symmetric memory unavailable on this box)"),
                "algorithmic_tflops": (sum(fwd5) + sum(bwd5)) / (t5 / n5) / 1e12,
                "replicas_bit_identical": same, "finite": bool(torch.isfinite(e5.theta_w).all())}
    del e5, X5l, Y5l, flat5, peer5
    torch.cuda.empty_cache()

    # ---- the collective part ends here: the other ranks leave, rank 0 finishes its single-GPU measurements ----
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    pk, pk_src = peaks()
    tf32_peak = measure_tf32_peak(dev)
    # ---- per-kernel durations, live (CUDA events on the launch stream, inside libdgprf) --------
    _ffi.profile_start()
    PROF_STEPS = 50
    for i in range(PROF_STEPS):
        step(i)
    recs = _ffi.profile_stop()
    launches_per_step = max(1, len(recs) // PROF_STEPS)          # launches the library's event hook saw per step
    per = {}
    for idx, (nm, ms) in enumerate(recs):
        per.setdefault((nm, idx % launches_per_step), []).append(ms)
    kernels = [{"kernel": nm, "slot": slot, "avg_us": 1e3 * statistics.mean(v)} for (nm, slot), v in sorted(per.items(), key=lambda kv: kv[0][1])]
    step_us = sum(k["avg_us"] for k in kernels)
    fwd_f, bwd_f = algorithmic_flops(e.spec, B)
    L = CFG["L"]
    n_fwd = n_bwd = 0
    for k in kernels:                                            # in launch order: forward layers up, backward layers down
        nm = k["kernel"]
        if nm.startswith("k9_") or nm.startswith("k10_"):
            k["flops"] = sum(fwd_f) + sum(bwd_f)
            k["what"] = ("cluster-split row-fused step (3xTF32 mma.sync): " if nm.startswith("k10_") else "row-fused step (FFMA): ") + \
                "forward + likelihood seed + backward (all layers) + grid barrier + update"
        elif nm.startswith("k1_fwd"):
            k["flops"] = fwd_f[min(n_fwd, L - 1)]; k["what"] = f"fwd layer {n_fwd}"; n_fwd += 1
        elif nm.startswith("k2_bwd"):
            l = max(L - 1 - n_bwd, 0)
            k["flops"] = bwd_f[l]; k["what"] = f"bwd layer {l}"; n_bwd += 1
        elif nm.startswith("k5_"):
            k["bytes"] = 20.0 * e.layout.w_len; k["what"] = "sgmcmc update"
        elif nm.startswith("k3_"):
            k["what"] = "likelihood seed"
        else:
            k["what"] = "operand prep"
        k["share"] = k["avg_us"] / step_us
    dom = max((k for k in kernels if "flops" in k or "bytes" in k), key=lambda k: k["avg_us"])
    if "flops" in dom:
        ach = dom["flops"] / (dom["avg_us"] * 1e-6) / 1e12
        traffic = next((v for pre, v in NCU_TRAFFIC.items() if dom["kernel"].startswith(pre)), (None, None))
        roof = {"bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                "traffic": traffic[0],
                "traffic_source": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full of this kernel on this workload ({traffic[1]})",
                "kernel": f'{dom["kernel"]} ({dom["what"]})', "avg_us": dom["avg_us"],
                "peak_source": "cuBLAS TF32 matmul 8192^3 measured live in this run (best of 10)",
                "binding_resource": "issue slots / dependent-phase latency, NOT the tensor pipe: configs[1] is 0.18 GFLOP and 0.4 MB of "
                                    "parameters per step; ncu (profiles/r02_ncu_summary.md): tensor pipe 14 % active (legacy mma.sync tf32 "
                                    "measures 1024 FLOP/clk/SM on B200, a third of that with the 3xTF32 split), issue slots 44 %, 16 warps/SM",
                "note": "roofline_tc_layer and roofline_k5_256MiB in this line give the tcgen05 and update kernels at throughput-relevant sizes"}
    else:
        ach = dom["bytes"] / (dom["avg_us"] * 1e-6) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": None, "kernel": dom["kernel"], "avg_us": dom["avg_us"], "peak_source": pk_src}

    # ---- the per-sample cost besides the cycle of steps: the evaluation forward over the test split (10 % of N) -------------
    n_test = N // 10
    Xe, Ye = X[:n_test].contiguous(), Y[:n_test].contiguous()
    for _ in range(3):
        model.eval_log_likelihood_and_se([(Xe, Ye)])
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s0.record(stream)
    for _ in range(10):
        model.eval_log_likelihood_and_se([(Xe, Ye)])
    s1.record(stream)
    torch.cuda.synchronize()
    eval_ms = s0.elapsed_time(s1) / 10

    # ---- one epoch of sampling steps as ONE CUDA-graph launch (the sampler drivers' graph=True mode) ----------
    from experiments.utils_dataset import DeviceDataset
    from experiments.utils_training import EpochGraph
    ds_g = DeviceDataset(X, Y, B, shuffle=True, drop_remainder=True, seed=1)
    ds_g.reshuffle()
    eg = EpochGraph(model, ds_g, N, [CFG["lr"]] * len(ds_g), CFG["momentum_decay"], CFG["temperature"], False, False)
    for _ in range(2):
        ds_g.reshuffle(); eg.replay()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    g0.record(stream)
    G_EPOCHS = 10
    for _ in range(G_EPOCHS):
        ds_g.reshuffle()                 # device-side permutation + gather, inside the timed region
        eg.replay()
    g1.record(stream)
    torch.cuda.synchronize()
    graph_it_s = G_EPOCHS * len(ds_g) / (g0.elapsed_time(g1) * 1e-3)
    graph_epoch = {"value": graph_it_s, "unit": UNIT, "steps_per_launch": len(ds_g), "epochs_timed": G_EPOCHS,
                   "note": "regression_train(..., graph=True): one captured CUDA graph per epoch (device-resident shuffled dataset, "
                           "Philox step base in device memory); the per-epoch reshuffle is inside the timed region"}
    del eg, ds_g

    # ---- K5 on a >= 256 MiB flat buffer: the honest HBM number for the update kernel --------------
    n_big = 16 << 20
    Cn = 4                       # 4 chains x 16 Mi parameters x 4 B = 256 MiB per buffer
    th = torch.randn(Cn, n_big, device=dev); mo = torch.randn(Cn, n_big, device=dev); gr = torch.randn(Cn, n_big, device=dev)
    segs = _ffi.make_segments([(0, n_big, 1.0, 1)])
    Lb = _ffi.lib()

    def big(stepno):
        _ffi.check(Lb.dgprf_sgmcmc_update(th.data_ptr(), mo.data_ptr(), n_big, n_big, Cn, gr.data_ptr(), n_big, 1, 0,
                                          segs, 1, 1e-4, float(N), 0.9, 1.0, 0, 7, stepno, None, None, stream.cuda_stream))
    for i in range(3):
        big(i)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    c0.record(stream)
    for i in range(10):
        big(10 + i)
    c1.record(stream)
    torch.cuda.synchronize()
    k5_ms = c0.elapsed_time(c1) / 10
    k5_gbs = 20.0 * Cn * n_big / (k5_ms * 1e-3) / 1e9
    del th, mo, gr

    # ---- tensor-core kernels at configs[4] layer scale (tf32 mode): the GEMM rooflines ------------------
    # One [RF -> GP] layer, B=65536 rows, input width 120, M=4096 features, n_gp=30: forward (3xTF32 phase
    # GEMM + sincos epilogue + Phi.W, saved features stored by TMA) and backward (three tf32 UMMAs per tile,
    # saved features streamed back by TMA).  Kernel times from the library's CUDA-event hook.
    tB, td, tM, tg = 65536, 120, 4096, 30
    tspec = ModelSpec.build(td, tg, [tM], [tg], ["RBF"], False, False, "gaussian")
    te = Engine(tspec, 1, precision=_ffi.PREC_TF32)
    te.theta_w.normal_()
    te.theta_h[:, te.layout.off_lik_log_var] = -2.0
    tX = torch.randn(tB, td, device=dev); tY = torch.randn(tB, tg, device=dev)
    for _ in range(3):
        te.gradients(tX, tY, 1e5, hyper=False, prior_w=True, prior_h=False)
    torch.cuda.synchronize()
    _ffi.profile_start()
    TREP = 5
    for _ in range(TREP):
        te.gradients(tX, tY, 1e5, hyper=False, prior_w=True, prior_h=False)
    tper = {}
    for nm, ms in _ffi.profile_stop():
        tper.setdefault(nm, []).append(ms)
    tF = 2 * tM
    phi_bytes = 4.0 * tB * tF

    def tc_entry(name, alg_flops, exe_flops, what):
        if name not in tper:
            return None
        us = 1e3 * statistics.mean(tper[name])
        return {"kernel": name, "what": what, "avg_us": us,
                "algorithmic_tflops": alg_flops / us / 1e6, "executed_tensor_tflops": exe_flops / us / 1e6,
                "tensor_frac_algorithmic": alg_flops / us / 1e6 / tf32_peak, "tensor_frac_executed": exe_flops / us / 1e6 / tf32_peak,
                "saved_feature_gbs": phi_bytes / us / 1e3, "hbm_frac": phi_bytes / us / 1e3 / pk["hbm_gbs"],
                "hbm_peak_gbs": pk["hbm_gbs"], "hbm_peak_kind": f"{pk_src} copy peak (MEASURED_PEAKS.json hbm_gbs)"}
    tc_layer = {
        "workload": f"one [RF->GP] layer at configs[4] scale: B={tB}, d={td}, M={tM}, n_gp={tg}, RBF, tf32 mode",
        "tf32_peak_tflops": tf32_peak, "tf32_peak_source": "cuBLAS TF32 matmul 8192^3, measured live (best of 10)",
        "hbm_copy_peak_gbs": pk["hbm_gbs"],
        "traffic_ncu_bytes": {"fwd": 2.105e9 + 84.0e6, "bwd": 2.173e9 + 63.1e6, "algorithmic_phi_bytes": phi_bytes,
                              "source": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full of layer 1 of a 3-layer configs[4]-shaped "
                                        "model (same tile shapes; the backward also writes the dF slabs of the layer below): "
                                        "profiles/r02_ncu_summary.md, profiles/r02_ncu_tc_final_raw.csv"},
        "fwd": tc_entry("k1_fwd_tc2", 2.0 * tB * (td * tM + tF * tg), 2.0 * tB * (3 * 128 * tM + tF * tg),
                        "3xTF32 phase GEMM (A in TMEM) + sincos epilogue + Phi.W; Phi stored by TMA as 32-column sub-tiles through a ring of three "
                        "buffers (tile-blocked layout in HBM); bound by the in-order tensor pipe + the TMA store, see profiles/r02_summary.md"),
        "bwd": tc_entry("k2_bwd_tc2", 4.0 * tB * tF * tg, 4.0 * tB * tF * 32,
                        "dPhi = dF.W^T, gW += Phi^T.dF (accumulators resident in TMEM); Phi loaded (TMA ring)"),
    }
    del te, tX, tY
    torch.cuda.empty_cache()
    clk = clocks.stop() if clocks else None

    it_s = world * K / t_dev
    prec_note = ("fp32: one cooperative cluster-split step kernel per iteration (thread-block clusters of 4 over 32-row tiles, 3xTF32 "
                 "mma.sync GEMM chains in registers, DSMEM exchange, update behind a grid barrier)" if args.precision == "fp32" else
                 "tf32: tcgen05 forward (3xTF32 phase GEMM, tf32 Phi*W) and tcgen05 backward, layered launches")
    line = {
        "metric": METRIC, "value": it_s, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": CFG,
        "timing": "value: per-step CUDA events on the launch stream, L2 flushed between timed steps (256 MiB write), the launch queue primed "
                  "behind a 1 ms device-side sleep so that host launch jitter is not counted as device time; max over ranks",
        "precision": prec_note, "parallelism": f"{world} independent chain(s), 1 per GPU, no data-path collective",
        "posterior_samples_per_second": world / (50 * nb * (t_dev / K) + eval_ms * 1e-3),
        "samples_note": f"one posterior sample = a cycle of 50 epochs x {nb} it (SURVEY 8d) + the evaluation forward over the {n_test}-point "
                        f"test split ({eval_ms:.3f} ms, models/regression_model.py:33-50)",
        "graph_epoch": graph_epoch,
        "warm_loop": {"value": world * K / t_warm, "unit": UNIT, "note": "back-to-back steps, no L2 flush, CPU launch cost included"},
        "e2e": {"value": world * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 4 * B * (CFG["D"] + 1), "d2h_bytes_per_step": 4,
                "sync": "end of loop: the sampler never waits on a step; the H2D copies of every step's pinned minibatch run on a side stream "
                        "under the previous step's kernel (double-buffered staging, event-ordered), sum_i ll_i is written to pinned host "
                        "memory by the kernel; one barrier + synchronize after the K steps"},
        "gpu_launches": launches_per_step * K,
        "launches_per_step": launches_per_step,
        "roofline": roof,
        "kernels": kernels,
        "roofline_k5_256MiB": {"bound": "hbm", "achieved": k5_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                               "frac": k5_gbs / pk["hbm_gbs"], "ms": k5_ms, "bytes": 20.0 * Cn * n_big,
                               "peak_source": pk_src},
        "cfg4_chains": cfg4_out,
        "cfg5_dp": cfg5_out,
        "roofline_tc_layer": tc_layer,
        "cpu_baseline": cpu_base,
        "clocks": clk,
        "wall_s_flushed_loop": wall_flushed,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DGPRF_PRECISION", "fp32"), choices=["fp32", "tf32"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device for --impl ours (no CPU fallback); use --impl reference for the CPU arm")
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
