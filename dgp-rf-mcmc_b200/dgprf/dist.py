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
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


_DP_OVERLAP_DEFAULT = os.environ.get("DGPRF_DP_OVERLAP", "0") == "1"
_DP_REDUCTION_DEFAULT = os.environ.get("DGPRF_DP_REDUCTION", "auto")   # "auto" | "nccl" | "peer" (csrc/k11_peer_allreduce.cu)


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


def gradient_group(max_ctas: int = 16, ranks=None):
    """A dedicated NCCL communicator for the gradient all-reduce of the data-parallel step, limited to `max_ctas` CTAs.
    The pipelined backward kernels run one CTA per SM (227 KB of shared memory each); at the row counts of an 8-GPU split
    they leave ~20 SMs free.  A collective that spreads over more SMs than that is not slower itself, but the NEXT backward
    grid then cannot become resident at once and runs a second wave -- the overlapped reduction must stay inside the free
    SMs.  Returns None (the default group) on backends without the option (gloo)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend() != "nccl":
        return None
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = int(max_ctas)
    opts.config.min_ctas = 1
    return dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)


class PeerAllReduce:
    """Flat-gradient all-reduce of the data-parallel step over NVLink peer memory (csrc/k11_peer_allreduce.cu) instead of a
    library collective: `torch.distributed._symmetric_memory` only allocates and maps the buffer [2][n_pad] (+ a zeroed signal
    pad) on every rank of the group; the reduction itself is three launches of this library on the caller's stream.
    `grad` (this rank's [n] gradient, written in place by the step's kernels) and `reduced` ([n], identical bits on every
    rank after `__call__`) are views of the symmetric buffer."""

    def __init__(self, n: int, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        q = 4 * self.world
        self.n, self.n_pad = int(n), (int(n) + q - 1) // q * q
        self.buf = symm.empty(2 * self.n_pad, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        assert self.hdl.world_size == self.world and self.hdl.rank == self.rank
        self._bufs = (C.c_void_p * self.world)(*[int(p) for p in self.hdl.buffer_ptrs])
        self._sigs = (C.c_void_p * self.world)(*[int(p) for p in self.hdl.signal_pad_ptrs])
        assert self.hdl.signal_pad_size >= 4 * (self.SIG_WORD + 2 * self.world)
        self.grad = self.buf[:self.n]
        self.reduced = self.buf[self.n_pad:self.n_pad + self.n]
        self.epoch = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)                      # every rank's buffer is zeroed and mapped before the first peer access

    SIG_WORD = 256                               # first signal word used (torch's own barrier / put_signal channels sit below)

    def __call__(self):
        from . import _ffi
        self.epoch += 1
        _ffi.check(_ffi.lib().dgprf_peer_allreduce(self._bufs, self._sigs, self.rank, self.world, self.n_pad,
                                                   self.epoch, self.SIG_WORD, _ffi.stream_ptr()))

    def status(self) -> int:
        """0 = healthy; otherwise (phase + 1) << 8 | peer of the first wait that timed out (synchronises the device)."""
        import ctypes as C
        from . import _ffi
        s = C.c_uint32(0)
        _ffi.check(_ffi.lib().dgprf_peer_allreduce_status(C.byref(s)))
        return int(s.value)


def _make_peer(n: int, device, group, required: bool):
    """PeerAllReduce, or None on EVERY rank when any rank cannot set it up (the ranks agree through one small all-reduce, so
    a partial failure can never leave some ranks in NCCL and others in the peer protocol)."""
    pr, err = None, None
    try:
        pr = PeerAllReduce(n, device, group)
    except Exception as exc:                         # symmetric memory not supported here (driver, topology, permissions)
        err = exc
    ok = torch.tensor([1 if pr is not None else 0], device=device, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 1:
        return pr
    if required:
        raise RuntimeError(f"reduction='peer' is not available on every rank of the group: {err!r}")
    return None


def bucket_layers(off_W, w_len: int, min_bucket_floats: int):
    """Layers whose hook closes a bucket of the overlapped reduction.  Buckets are contiguous runs [off_W[l], hi) of the flat
    buffer [gW | sum ll] (w_len + 1 floats), built from the top layer down -- the order the reverse pass retires the layers:
    a layer closes a bucket once the run holds min_bucket_floats, layer 0 closes the last one.  Returns the set of closing
    layers; together the buckets tile [0, w_len + 1) exactly once."""
    flush_at = set()
    hi = w_len + 1
    for l in range(len(off_W) - 1, -1, -1):
        if l == 0 or hi - off_W[l] >= min_bucket_floats:
            flush_at.add(l)
            hi = off_W[l]
    return flush_at


class _OverlapState:
    """Side stream, per-layer events and bucket boundaries of the overlapped data-parallel step (one per engine)."""

    def __init__(self, engine, min_bucket_floats: int):
        lay = engine.layout
        self.comm = torch.cuda.Stream(device=engine.device)
        self.events = [torch.cuda.Event() for _ in lay.off_W]
        self.flush_at = bucket_layers(lay.off_W, lay.w_len, min_bucket_floats)


def data_parallel_step(engine, X_local: torch.Tensor, Y_local: torch.Tensor, global_rows: int, data_size: float,
                       lr: float, momentum_decay: float, temperature: float = 1.0, resample: bool = False,
                       seed: int = 0, step: int = 0, group=None, overlap: Optional[bool] = None,
                       min_bucket_floats: int = 65536, reduction: Optional[str] = None) -> torch.Tensor:
    """One W-only sgmcmc_update (models/dgp.py:184-216) of a minibatch whose rows are split over the ranks.

    Every rank runs forward / likelihood seed / backward on its rows (CUDA kernels, data term only) with the seed scaled
    by 1 / B_global, so its flat gradient is already its share of the global one; the flat buffer [gW | sum_i ll_i]
    (written in place by the kernels, no staging copy) is all-reduced (sum), and every rank applies the update kernel to
    its replica with the same Philox (seed, step): the prior term theta/N is added inside the update, once, after the
    reduction; replicas stay bit-identical without a broadcast.  Returns sum_i ll_i [C].

    reduction: "auto" (default; DGPRF_DP_REDUCTION) = "peer" on NCCL process groups when symmetric memory can be set up on every
    rank, else "nccl".  "peer" is this library's two-shot all-reduce over NVLink peer memory (csrc/k11_peer_allreduce.cu): every
    rank sums its 1/world slice of all ranks' gradients in rank order and pushes the result to every rank -- the replicas hold
    identical bits by construction.  "nccl" is one `dist.all_reduce` of the whole buffer after the reverse pass.

    overlap=True (opt-in, NCCL only; DGPRF_DP_OVERLAP=1 makes it the default for world > 1): the reverse pass retires the layers
    top-down, and a layer's gradient slice is final as soon as its backward kernel has run -- so the slab sum of that slice and
    its all-reduce are issued on a side stream from a per-layer host hook (dgprf_set_backward_hook) and run UNDER the backward
    kernels of the layers below; only the bucket of layer 0 is exposed.  Every element is still reduced exactly once, so the
    replicas stay bit-identical.  MEASURED on 2 / 4 / 8 B200s over NVSwitch (profiles/r02_dp_overlap_ab_{2,4,8}gpu.txt, A/B
    inside one job): not faster -- the whole 4 MB all-reduce is 35-55 us of a 1-3 ms step and latency-bound, four 1 MB
    all-reduces cost ~45 us each, their kernels slow the one-CTA-per-SM backward kernels they run beside, and the event
    records between the backward kernels give up their programmatic launch overlap (8 GPUs: 0.965 ms with one all-reduce,
    0.988 ms with per-layer buckets, 1.03-1.12 ms on communicators limited to 8 / 4 CTAs).  Hence opt-in: the reduction runs
    once, after the reverse pass (peer two-shot 0.938 ms, one NCCL all-reduce 0.961 ms)."""
    from . import _ffi
    w_len = engine.layout.w_len
    assert engine.C == 1, "the data-parallel step drives one replica per rank"
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if reduction is None:
        reduction = _DP_REDUCTION_DEFAULT
    assert reduction in ("auto", "nccl", "peer"), reduction
    if overlap is None:
        overlap = world > 1 and _DP_OVERLAP_DEFAULT
    pr = None
    if reduction != "nccl" and world > 1 and not overlap and dist.get_backend(group) == "nccl":
        # reduction="peer" (what "auto" resolves to on NCCL process groups): the library's own two-shot all-reduce over NVLink
        # peer memory (PeerAllReduce above): the step's kernels write [gW | sum ll] straight into this rank's half of the
        # symmetric buffer, three small launches reduce it, the update reads the reduced half -- no NCCL call on the step
        # (8 GPUs: 37 us against NCCL's 55 us for the 4 MB buffer, configs[4] step 0.961 -> 0.938 ms)
        pr = getattr(engine, "_dp_peer", None)
        if pr is None or pr[0] is not group:
            pr = engine._dp_peer = (group, _make_peer(w_len + 1, engine.device, group, required=reduction == "peer"))
        pr = pr[1]                                   # None: symmetric memory unavailable on some rank -> NCCL on every rank
    if pr is not None:
        engine.gradients(X_local, Y_local, data_size, hyper=False, prior_w=False, prior_h=False,
                         inv_B=1.0 / float(global_rows), out_flat=pr.grad)
        pr()
        sw, nsw, _, _ = engine._segments()
        _ffi.check(_ffi.lib().dgprf_sgmcmc_update(
            engine.theta_w.data_ptr(), engine.mom_w.data_ptr(), w_len, w_len, 1, pr.reduced.data_ptr(), w_len, 1, 0,
            sw, nsw, float(lr), float(data_size), float(momentum_decay), float(temperature), int(bool(resample)),
            int(seed), int(step), None, None, _ffi.stream_ptr()))
        return pr.reduced[w_len:]
    flat = getattr(engine, "_dp_flat", None)
    if flat is None:
        flat = engine._dp_flat = torch.zeros(w_len + 1, device=engine.device, dtype=torch.float32)
    if overlap:
        key = ("_dp_overlap", int(min_bucket_floats))
        st = getattr(engine, "_dp_overlap_state", None)
        if st is None or st[0] != key:
            st = engine._dp_overlap_state = (key, _OverlapState(engine, int(min_bucket_floats)))
        st = st[1]
        main = torch.cuda.current_stream()
        off_W = engine.layout.off_W
        hi = [w_len + 1]

        def hook(l, finalize_layer):
            ev = st.events[l]
            ev.record(main)
            st.comm.wait_event(ev)
            finalize_layer(l, st.comm.cuda_stream)
            if l in st.flush_at:
                if world > 1:
                    with torch.cuda.stream(st.comm):
                        dist.all_reduce(flat[off_W[l]:hi[0]], op=dist.ReduceOp.SUM, group=group)
                hi[0] = off_W[l]

        engine.gradients(X_local, Y_local, data_size, hyper=False, prior_w=False, prior_h=False,
                         inv_B=1.0 / float(global_rows), out_flat=flat, layer_hook=hook)
        assert hi[0] == 0, "a layer of the reverse pass did not report"
        main.wait_stream(st.comm)
    else:
        # forward / seed / backward of the local rows with the seed already scaled by 1 / B_global: the flat buffer
        # [gW | sum_i ll_i] is written in place by the kernels and is the all-reduce payload as it stands
        engine.gradients(X_local, Y_local, data_size, hyper=False, prior_w=False, prior_h=False,
                         inv_B=1.0 / float(global_rows), out_flat=flat)
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    sw, nsw, _, _ = engine._segments()
    _ffi.check(_ffi.lib().dgprf_sgmcmc_update(
        engine.theta_w.data_ptr(), engine.mom_w.data_ptr(), w_len, w_len, 1, flat.data_ptr(), w_len, 1, 0,
        sw, nsw, float(lr), float(data_size), float(momentum_decay), float(temperature), int(bool(resample)),
        int(seed), int(step), None, None, _ffi.stream_ptr()))
    return flat[w_len:]
