"""A/B of the gradient reduction of the configs[4] data-parallel step inside ONE torchrun job (same replicas, same clocks):
one all-reduce of the whole flat buffer after the reverse pass against the per-layer buckets reduced under the reverse pass,
on the default communicator and on communicators limited to a few CTAs (dgprf/dist.py: gradient_group).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/dp_overlap_ab.py
Rank 0 prints one line per variant: max-over-ranks ms per step (device events), three repeats."""
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
assert world > 1, "launch under torchrun with at least two ranks"
dist.init_process_group("nccl", device_id=dev)
B, D_in, M, L, N = int(os.environ.get("DP_GLOBAL_BATCH", 65536)), 90, 4096, 5, 515345
spec = ModelSpec.build(D_in, 1, [M] * L, [30, 30, 30, 30, 1], ["RBF"] * L, True, False, "gaussian")
torch.manual_seed(0)
e = Engine(spec, 1, device=dev, precision=_ffi.PREC_TF32)
e.theta_w.normal_()
e.theta_h[:, e.layout.off_lik_log_var] = -2.0
g = torch.Generator(device=dev).manual_seed(1)
X = torch.randn(B, D_in, device=dev, generator=g)
Y = torch.randn(B, 1, device=dev, generator=g)
Xl, Yl = D.row_shard(X, Y, rank, world)
Xl, Yl = Xl.contiguous(), Yl.contiguous()
base = dict(global_rows=B, data_size=N, lr=1e-4, momentum_decay=0.9, seed=3)
groups = {0: None}
for n in (4, 8, 16):
    groups[n] = D.gradient_group(max_ctas=n)
VARIANTS = [  # (name, overlap, max_ctas, min_bucket_floats)
    ("one all-reduce after the reverse pass", False, 0, 65536),
    ("PEER two-shot all-reduce over NVLink peer memory (k11)", "peer", 0, 65536),
    ("one all-reduce after the reverse pass (again)", False, 0, 65536),
    ("PEER two-shot all-reduce (again)", "peer", 0, 65536),
    ("one all-reduce, 16 CTAs", False, 16, 65536),
    ("per-layer buckets under the reverse pass", True, 0, 65536),
    ("per-layer buckets, 16 CTAs", True, 16, 65536),
    ("per-layer buckets, 8 CTAs", True, 8, 65536),
    ("per-layer buckets, 4 CTAs", True, 4, 65536),
    ("two buckets [L4 L3 L2 | L1 L0], 8 CTAs", True, 8, 600000),
    ("one bucket issued from the hook of layer 0, 8 CTAs", True, 8, 1 << 30),
]
if os.environ.get("DP_AB_ONLY"):                      # e.g. DP_AB_ONLY="one all-reduce after,PEER": substring filter
    keys = os.environ["DP_AB_ONLY"].split(",")
    VARIANTS = [v for v in VARIANTS if any(k in v[0] for k in keys)]
STEPS = 20
step = 0
for name, ov, ctas, mb in VARIANTS:
    kw = dict(base, overlap=ov, min_bucket_floats=mb, group=groups[ctas])
    if ov == "peer":
        kw = dict(base, reduction="peer")
    res = []
    for rep in range(3):
        for i in range(3):
            D.data_parallel_step(e, Xl, Yl, step=step, **kw); step += 1
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(STEPS):
            D.data_parallel_step(e, Xl, Yl, step=step, **kw); step += 1
        b.record(); torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / STEPS], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res.append(round(ms.item(), 4))
    chk = e.theta_w.double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"variant": name, "n_gpus": world, "rows_per_gpu": Xl.shape[0], "ms_per_step": res,
                          "replicas_bit_identical": bool(lo.item() == hi.item()), "finite": bool(torch.isfinite(e.theta_w).all())}), flush=True)
dist.barrier()
dist.destroy_process_group()
