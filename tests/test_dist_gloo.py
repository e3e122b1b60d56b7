"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in dgprf/dist.py: chain sharding,
the single flat-gradient all-reduce of the data-parallel step, and the sharded predictive average.
The arithmetic stand-in on CPU is the oracle (tests only)."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _worker_dp(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.join(here, "..", "dgp-rf-mcmc_b200"), os.path.join(here, "..", "oracle")):
        sys.path.insert(0, p)
    import dgprf_oracle as O
    from dgprf import dist as D
    _init(rank, world, port)
    torch.manual_seed(0)
    N, B = 1000, 37                       # odd batch: ranks get 19 / 18 rows
    p = O.init_params(4, 1, 2, 12, [3, 1], None, True, "gaussian", seed=5)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(B, 4, generator=g, dtype=torch.float64)
    Y = torch.randn(B, 1, generator=g, dtype=torch.float64)
    # single-process truth on the whole minibatch
    u_ref, g_ref = O.grads_autograd(p, X, Y, N, full_bayesian=True)
    # data-parallel: rank-local gradient of the data term only (prior off), rescaled, one all-reduce
    Xl, Yl = D.row_shard(X, Y, rank, world)
    _, g_loc = O.grads_analytic(p, Xl, Yl, N, full_bayesian=False, hyper=True, allow_gradient_from_W=False)
    names_w = [n for n in g_loc if n.startswith("W_")]
    names_h = [n for n in g_loc if not n.startswith("W_")]
    s = D.dp_scale(Xl.shape[0], B)
    gW = torch.cat([g_loc[n].reshape(-1) for n in names_w]) * s
    gH = torch.cat([g_loc[n].reshape(-1) for n in names_h]) * s
    ll = O.log_likelihood(p, Xl, Yl).sum().reshape(1)
    gW, gH, ll = D.allreduce_flat_gradient(gW, gH, ll)
    # prior added once after the reduction
    named = dict(p.w_named() + p.hyper_named())
    off = 0
    ok = True
    for n in names_w:
        k = named[n].numel()
        full = gW[off:off + k].view_as(named[n]) + named[n] / N
        ok &= torch.allclose(full, g_ref[n], rtol=1e-10, atol=1e-12)
        off += k
    off = 0
    for n in names_h:
        k = named[n].numel()
        full = gH[off:off + k].view_as(named[n]) + named[n] / N
        ok &= torch.allclose(full, g_ref[n].reshape(named[n].shape), rtol=1e-10, atol=1e-12)
        off += k
    ll_ref = O.log_likelihood(p, X, Y).sum()
    ok &= bool(abs(float(ll) - float(ll_ref)) < 1e-9)
    # every rank holds the identical reduced buffer -> identical update without a broadcast
    chk = gW.clone()
    dist.broadcast(chk, src=0)
    ok &= bool(torch.equal(chk, gW))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _worker_buckets(rank, world, port, q):
    """The bucketed reduction of the overlapped data-parallel step (dgprf/dist.py: bucket_layers): all-reducing the flat buffer
    [gW | sum ll] bucket by bucket, top layer first, gives every rank the same buffer as one all-reduce of the whole of it."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.join(here, "..", "dgp-rf-mcmc_b200"), os.path.join(here, "..", "oracle")):
        sys.path.insert(0, p)
    import dgprf_oracle as O
    from dgprf import dist as D
    from dgprf.engine import FlatLayout, ModelSpec
    _init(rank, world, port)
    N, B = 1000, 41
    n_rf, n_gp = [12, 8, 12], [3, 2, 1]
    spec = ModelSpec.build(4, 1, n_rf, n_gp, ["RBF"] * 3, True, False, "gaussian")
    lay = FlatLayout.of(spec)
    p = O.init_params(4, 1, 3, n_rf, n_gp, None, True, "gaussian", seed=6)
    g = torch.Generator().manual_seed(2)
    X = torch.randn(B, 4, generator=g, dtype=torch.float64)
    Y = torch.randn(B, 1, generator=g, dtype=torch.float64)
    Xl, Yl = D.row_shard(X, Y, rank, world)
    _, g_loc = O.grads_analytic(p, Xl, Yl, N, full_bayesian=False, hyper=False, allow_gradient_from_W=False)
    flat = torch.zeros(lay.w_len + 1, dtype=torch.float64)
    for l in range(3):
        w = g_loc[f"W_{l}"].reshape(-1) * D.dp_scale(Xl.shape[0], B)
        flat[lay.off_W[l]:lay.off_W[l] + w.numel()] = w
    flat[lay.w_len] = O.log_likelihood(p, Xl, Yl).sum()
    whole = flat.clone()
    dist.all_reduce(whole)
    ok = True
    for min_bucket in (1, 40, 1 << 30):
        closing = D.bucket_layers(lay.off_W, lay.w_len, min_bucket)
        buf, hi, n_calls = flat.clone(), lay.w_len + 1, 0
        for l in (2, 1, 0):                               # the order the reverse pass reports the layers
            if l in closing:
                dist.all_reduce(buf[lay.off_W[l]:hi])
                hi = lay.off_W[l]
                n_calls += 1
        ok &= hi == 0 and n_calls == len(closing)
        ok &= bool(torch.equal(buf, whole))               # two ranks: one addition per element, identical whatever the split
    ok &= D.bucket_layers(lay.off_W, lay.w_len, 1) == {0, 1, 2} and D.bucket_layers(lay.off_W, lay.w_len, 1 << 30) == {0}
    ok &= D.gradient_group(max_ctas=8) is None            # CTA-limited communicators are an NCCL option: default group on gloo
    # the peer-memory all-reduce needs CUDA symmetric memory: where a rank cannot set it up, EVERY rank agrees on the library
    # collective (None), and an explicit reduction="peer" raises instead of splitting the ranks between two protocols
    ok &= D._make_peer(64, torch.device("cpu"), None, required=False) is None
    try:
        D._make_peer(64, torch.device("cpu"), None, required=True)
        ok = False
    except RuntimeError as exc:
        ok &= "not available on every rank" in str(exc)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _worker_pred(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.join(here, "..", "dgp-rf-mcmc_b200"), os.path.join(here, "..", "oracle")):
        sys.path.insert(0, p)
    import dgprf_oracle as O
    from dgprf import dist as D
    _init(rank, world, port)
    g = torch.Generator().manual_seed(3)
    S, N = 7, 50
    lp = 2.0 * torch.randn(S, N, generator=g, dtype=torch.float64)
    se = torch.rand(S, N, generator=g, dtype=torch.float64)
    a_ref, b_ref = O.predictive_average(lp, se)
    lo, hi = D.shard_range(S, rank, world)            # sample sets sharded per rank (4 / 3 samples)
    a, b = D.combine_predictive(torch.logsumexp(lp[lo:hi], 0), hi - lo, se[lo:hi].sum(), True)
    q.put((rank, abs(a - float(a_ref)) < 1e-12 and abs(b - float(b_ref)) < 1e-12))
    dist.destroy_process_group()


def _run(worker, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=120) for _ in procs]
    for p_ in procs:
        p_.join(timeout=60)
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(ok for _, ok in res), res


def test_shard_range_covers_everything():
    from dgprf.dist import shard_range
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(64, r, 8) for r in (0, 7)] == [(0, 8), (56, 64)]       # cfg4: 64 chains on 8 GPUs


def test_data_parallel_gradient_allreduce_matches_single_process():
    _run(_worker_dp, 29611)


def test_bucketed_gradient_reduction_matches_one_all_reduce():
    _run(_worker_buckets, 29613)


def test_sharded_predictive_average_matches_single_process():
    _run(_worker_pred, 29612)
