"""configs[4]: 5-layer RBF RF-DGP, M=4096, global minibatch 65536 rows split over the ranks (data parallel),
one NCCL all-reduce of the flat W gradient per step.  Launch: python -m torch.distributed.run --nnodes=1
--nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/dp_bench.py [steps]   (N = 1 works without torchrun).
Rank 0 prints one JSON line; every run also checks that the replicas hold bit-identical parameters."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
import torch
import torch.distributed as dist
from dgprf import _ffi, dist as D
from dgprf.engine import Engine, ModelSpec

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 20
PREC = os.environ.get("DGPRF_PRECISION", "tf32")
B, D_in, M, L, N = int(os.environ.get("DP_GLOBAL_BATCH", 65536)), 90, 4096, 5, 515345
n_gp = [30, 30, 30, 30, 1]
spec = ModelSpec.build(D_in, 1, [M] * L, n_gp, ["RBF"] * L, True, False, "gaussian")
torch.manual_seed(0)                                  # identical initial replicas
e = Engine(spec, 1, device=dev, precision={"fp32": _ffi.PREC_FP32, "tf32": _ffi.PREC_TF32}[PREC])
e.theta_w.normal_()
e.theta_h[:, e.layout.off_lik_log_var] = -2.0
g = torch.Generator(device=dev).manual_seed(1)
X = torch.randn(B, D_in, device=dev, generator=g)
Y = torch.randn(B, 1, device=dev, generator=g)
Xl, Yl = D.row_shard(X, Y, rank, world)
Xl, Yl = Xl.contiguous(), Yl.contiguous()
kw = dict(global_rows=B, data_size=N, lr=1e-4, momentum_decay=0.9, seed=3)
# A/B switches of the gradient reduction: DP_OVERLAP=0|1 (default: the library's), DP_MIN_BUCKET=floats per bucket,
# DP_MAX_CTAS=n: all-reduce on a dedicated NCCL communicator limited to n CTAs (so that its kernels fit on the SMs the
# one-CTA-per-SM backward kernels leave free and never delay the residency of the next backward grid)
if "DP_OVERLAP" in os.environ:
    kw["overlap"] = os.environ["DP_OVERLAP"] == "1"
if "DP_MIN_BUCKET" in os.environ:
    kw["min_bucket_floats"] = int(os.environ["DP_MIN_BUCKET"])
if world > 1 and "DP_MAX_CTAS" in os.environ:
    kw["group"] = D.gradient_group(max_ctas=int(os.environ["DP_MAX_CTAS"]))
for i in range(3):
    D.data_parallel_step(e, Xl, Yl, step=i, **kw)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(STEPS):
    D.data_parallel_step(e, Xl, Yl, step=10 + i, **kw)
b.record()
torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b) / STEPS], device=dev)
chk = e.theta_w.double().sum().reshape(1)
same = True
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(lo.item() == hi.item())
if rank == 0 and os.environ.get("DGPRF_BREAKDOWN"):
    _ffi.profile_start()
    D.data_parallel_step(e, Xl, Yl, step=999, **kw)
    recs = _ffi.profile_stop()
    print("kernel sum %.1f us:" % (1e3 * sum(t for _, t in recs)), " ".join(f"{nm}={t * 1e3:.0f}" for nm, t in recs))
if rank == 0:
    print(json.dumps({"workload": "configs[4] data-parallel step", "n_gpus": world, "precision": PREC, "global_batch": B,
                      "rows_per_gpu": Xl.shape[0], "ms_per_step": ms.item(), "it_per_s": 1e3 / ms.item(), "scaling": "strong",
                      "allreduce_bytes": 4 * (e.layout.w_len + 1), "replicas_bit_identical": same,
                      "switches": {k: os.environ[k] for k in ("DP_OVERLAP", "DP_MIN_BUCKET", "DP_MAX_CTAS") if k in os.environ},
                      "finite": bool(torch.isfinite(e.theta_w).all())}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
