"""Multi-GPU host logic (one process per GPU, `torch.distributed`; NCCL on GPUs, gloo in CPU tests).

Three sharding patterns of SURVEY section 8(e); only the third has a data-path collective:

  1. chains            independent (one DGP_RF instance == one chain): `shard_range` splits chain ids,
                       no communication while sampling;
  2. predictive sets   each rank reduces its own stored samples to per-test-point log-sum-exps and
                       sums; `combine_predictive` all-gathers [G, N_test] and finishes with a logsumexp of
                       logsumexps (experiments/utils_training.py:79-85 over the union of all samples);
  3. data parallel     large minibatches split by rows; every rank computes the gradient of its rows
                       scaled by 1/B_global, ONE all-reduce(sum) over the flat [gW | gH | ll_sum] buffer,
                       the prior term theta/N is added once after the reduction, and every rank applies
                       the identical update (same Philox key) -- replicas stay bit-identical, no broadcast.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n_items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def row_shard(X: torch.Tensor, Y: torch.Tensor, rank: int, world: int):
    """Rows of a global minibatch owned by `rank` (data-parallel step)."""
    lo, hi = shard_range(X.shape[0], rank, world)
    return X[lo:hi], Y[lo:hi]


def allreduce_flat_gradient(gW: torch.Tensor, gH: Optional[torch.Tensor], ll_sum: torch.Tensor,
                            group=None) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
    """One all-reduce(sum) over the concatenation [gW | gH | ll_sum] (a single NCCL launch)."""
    parts = [gW.reshape(-1)] + ([gH.reshape(-1)] if gH is not None else []) + [ll_sum.reshape(-1)]
    flat = torch.cat(parts)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    nW = gW.numel()
    nH = gH.numel() if gH is not None else 0
    return (flat[:nW].view_as(gW), flat[nW:nW + nH].view_as(gH) if gH is not None else None,
            flat[nW + nH:].view_as(ll_sum))


def dp_scale(local_rows: int, global_rows: int) -> float:
    """Factor that turns a rank-local dU/dtheta (data term, scaled by 1/B_local inside the kernels)
    into its share of the global-minibatch gradient (1/B_global), models/dgp.py:174."""
    return float(local_rows) / float(global_rows)


def combine_predictive(lse_local: torch.Tensor, n_samples_local: int, aux_sum_local: Optional[torch.Tensor] = None,
                       aux_is_se: bool = True, group=None):
    """lse_local [N_test]: logsumexp over this rank's samples; aux_sum_local: scalar sum of the aux
    matrix (squared errors or accuracies) over this rank's samples and all test points.
    Returns (mean_n(logsumexp_all - log S_total), sqrt(mean se) | mean acc)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    N = lse_local.numel()
    meta = torch.tensor([float(n_samples_local), float(aux_sum_local) if aux_sum_local is not None else 0.0],
                        dtype=torch.float64, device=lse_local.device)
    if world > 1:
        gathered = [torch.empty_like(lse_local) for _ in range(world)]
        dist.all_gather(gathered, lse_local.contiguous(), group=group)
        dist.all_reduce(meta, op=dist.ReduceOp.SUM, group=group)
        stack = torch.stack(gathered, 0)
    else:
        stack = lse_local[None]
    S = float(meta[0])
    lp = (torch.logsumexp(stack.double(), dim=0) - math.log(S)).mean()
    aux = None
    if aux_sum_local is not None:
        m = float(meta[1]) / (S * N)
        aux = math.sqrt(m) if aux_is_se else m
    return float(lp), aux


def data_parallel_step(engine, X_local: torch.Tensor, Y_local: torch.Tensor, global_rows: int, data_size: float,
                       lr: float, momentum_decay: float, temperature: float = 1.0, resample: bool = False,
                       seed: int = 0, step: int = 0, group=None) -> torch.Tensor:
    """One W-only sgmcmc_update (models/dgp.py:184-216) of a minibatch whose rows are split over the ranks.

    Every rank runs forward / likelihood seed / backward on its rows (CUDA kernels, data term only) with the seed scaled
    by 1 / B_global, so its flat gradient is already its share of the global one; ONE all-reduce(sum) of
    [gW | sum_i ll_i] (written in place by the kernels, no staging copy), and every rank applies the update
    kernel to its replica with the same Philox (seed, step): the prior term theta/N is added inside the update,
    once, after the reduction; replicas stay bit-identical without a broadcast.  Returns sum_i ll_i [C]."""
    from . import _ffi
    w_len = engine.layout.w_len
    assert engine.C == 1, "the data-parallel step drives one replica per rank"
    flat = getattr(engine, "_dp_flat", None)
    if flat is None:
        flat = engine._dp_flat = torch.empty(w_len + 1, device=engine.device, dtype=torch.float32)
    # forward / seed / backward of the local rows with the seed already scaled by 1 / B_global: the flat buffer
    # [gW | sum_i ll_i] is written in place by the kernels and is the all-reduce payload as it stands
    engine.gradients(X_local, Y_local, data_size, hyper=False, prior_w=False, prior_h=False, inv_B=1.0 / float(global_rows),
                     out_flat=flat)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    sw, nsw, _, _ = engine._segments()
    _ffi.check(_ffi.lib().dgprf_sgmcmc_update(
        engine.theta_w.data_ptr(), engine.mom_w.data_ptr(), w_len, w_len, 1, flat.data_ptr(), w_len, 1, 0,
        sw, nsw, float(lr), float(data_size), float(momentum_decay), float(temperature), int(bool(resample)),
        int(seed), int(step), None, None, _ffi.stream_ptr()))
    return flat[w_len:]
