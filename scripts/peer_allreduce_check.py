"""Peer-memory all-reduce (csrc/k11_peer_allreduce.cu, dgprf/dist.py: PeerAllReduce) against NCCL, under torchrun:
  1. random buffers: reduced half == dist.all_reduce (bit for bit at 2 ranks; to fp32 rounding beyond), identical on every rank;
  2. the configs[4] data-parallel step with reduction="peer" against reduction="nccl" from the same state;
  3. the all-reduce of the 4 MB flat gradient alone, and the whole step, both ways (max over ranks, device events).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/peer_allreduce_check.py"""
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
dist.init_process_group("nccl", device_id=dev)
out = {"n_gpus": world}


def max_ms(ms):
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def timed(fn, n, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return max_ms(a.elapsed_time(b) / n)


# ---- 1. random buffers of awkward sizes -------------------------------------------------------------------------
ok = True
for n in (1, 33, 4097, 991233):
    pr = D.PeerAllReduce(n, dev)
    for rep in range(3):
        g = torch.Generator(device=dev).manual_seed(100 * rank + rep + n)
        x = torch.randn(n, device=dev, generator=g)
        pr.grad.copy_(x)
        pr()
        ref = x.clone()
        dist.all_reduce(ref)
        got = pr.reduced.clone()
        err = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        chk = got.clone()
        dist.broadcast(chk, src=0)
        same = bool(torch.equal(chk, got))
        ok &= same and err <= (0.0 if world == 2 else 2e-6)
        if rank == 0:
            print(f"n={n} rep={rep}: max rel diff vs NCCL {err:.2e}  identical on all ranks: {same}", flush=True)
    ok &= pr.status() == 0
    del pr
out["random_buffers_ok"] = bool(ok)

# ---- 2./3. the data-parallel step ----------------------------------------------------------------------------------
B, D_in, M, L, N = int(os.environ.get("DP_GLOBAL_BATCH", 65536)), 90, 4096, 5, 515345
spec = ModelSpec.build(D_in, 1, [M] * L, [30, 30, 30, 30, 1], ["RBF"] * L, True, False, "gaussian")
engines = []
for _ in range(2):
    torch.manual_seed(0)
    e = Engine(spec, 1, device=dev, precision=_ffi.PREC_TF32)
    e.theta_w.normal_()
    e.theta_h[:, e.layout.off_lik_log_var] = -2.0
    engines.append(e)
ea, eb = engines
g = torch.Generator(device=dev).manual_seed(1)
X = torch.randn(B, D_in, device=dev, generator=g)
Y = torch.randn(B, 1, device=dev, generator=g)
Xl, Yl = D.row_shard(X, Y, rank, world)
Xl, Yl = Xl.contiguous(), Yl.contiguous()
kw = dict(global_rows=B, data_size=N, lr=1e-4, momentum_decay=0.9, seed=3)
for step in range(3):
    lla = D.data_parallel_step(ea, Xl, Yl, step=step, reduction="nccl", **kw).clone()
    llb = D.data_parallel_step(eb, Xl, Yl, step=step, reduction="peer", **kw).clone()
torch.cuda.synchronize()
rel = float((ea.theta_w - eb.theta_w).abs().max() / ea.theta_w.abs().max())
relm = float((ea.mom_w - eb.mom_w).abs().max() / ea.mom_w.abs().max())
chk = eb.theta_w.double().sum().reshape(1)
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
out["step_theta_rel_diff_peer_vs_nccl"] = rel
out["step_mom_rel_diff_peer_vs_nccl"] = relm
out["step_ll_rel_diff"] = float((lla - llb).abs().max() / lla.abs().max())
out["peer_replicas_bit_identical"] = bool(lo.item() == hi.item())
out["peer_status"] = eb._dp_peer[1].status()

flat = torch.zeros(ea.layout.w_len + 1, device=dev)
pr = eb._dp_peer[1]
out["allreduce_us"] = {"nccl": 1e3 * timed(lambda: dist.all_reduce(flat), 50), "peer": 1e3 * timed(lambda: pr(), 50)}
res = {"nccl": [], "peer": []}
s = 100
for rep in range(3):
    for red, e in (("nccl", ea), ("peer", eb)):
        def one():
            global s
            D.data_parallel_step(e, Xl, Yl, step=s, reduction=red, **kw); s += 1
        res[red].append(round(timed(one, 20, 3), 4))
out["ms_per_step"] = res
out["rows_per_gpu"] = int(Xl.shape[0])
out["peer_status_end"] = pr.status()
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
