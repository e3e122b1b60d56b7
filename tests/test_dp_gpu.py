"""Data-parallel sampling step (BASELINE configs[4]) on the CUDA path, single GPU:
the rank-local pieces must add up to the single-process step.  (The NCCL run over 2 / 4 GPUs is
scripts/dp_bench.py -- results in profiles/; the collective logic itself is covered on gloo in test_dist_gloo.py.)"""
import pytest
import torch

from dgprf import _ffi, dist as D
from dgprf.engine import Engine, ModelSpec
from helpers import rel_err

pytestmark = pytest.mark.gpu


def _engine(prec, seed=0):
    spec = ModelSpec.build(7, 2, [256, 256], [6, 2], ["RBF", "RBF"], True, False, "gaussian")
    torch.manual_seed(seed)
    e = Engine(spec, 1, precision=prec)
    e.theta_w.normal_()
    e.theta_h[:, e.layout.off_lik_log_var] = -1.0
    return e


@pytest.mark.parametrize("prec", [_ffi.PREC_FP32, _ffi.PREC_TF32])
def test_row_shards_sum_to_the_global_gradient(prec):
    e = _engine(prec)
    g = torch.Generator().manual_seed(3)
    B, N = 901, 20000.0
    X = torch.randn(B, 7, generator=g).cuda()
    Y = torch.randn(B, 2, generator=g).cuda()
    ll, gW, _ = e.gradients(X, Y, N, hyper=False, prior_w=False, prior_h=False)
    ll, gW = ll.clone(), gW.clone()
    acc_g, acc_ll = torch.zeros_like(gW), torch.zeros_like(ll)
    world = 3
    for rank in range(world):
        Xl, Yl = D.row_shard(X, Y, rank, world)
        ll_l, g_l, _ = e.gradients(Xl.contiguous(), Yl.contiguous(), N, hyper=False, prior_w=False, prior_h=False)
        acc_g += g_l * D.dp_scale(Xl.shape[0], B)
        acc_ll += ll_l
    tol = 1e-4 if prec == _ffi.PREC_FP32 else 3e-3
    assert rel_err(acc_g, gW) < tol
    assert rel_err(acc_ll, ll) < tol


def test_single_rank_data_parallel_step_equals_the_layered_step():
    """world = 1: gradients + flat 'all-reduce' + update kernel == dgprf_sgmcmc_step on the same seed / step (tf32 mode
    takes the layered kernels in both paths, so the two are the same arithmetic and agree to fp32 rounding)."""
    a, b = _engine(_ffi.PREC_TF32), _engine(_ffi.PREC_TF32)
    assert torch.equal(a.theta_w, b.theta_w)
    g = torch.Generator().manual_seed(4)
    B, N = 640, 5000.0
    X = torch.randn(B, 7, generator=g).cuda()
    Y = torch.randn(B, 2, generator=g).cuda()
    for step in range(3):
        a.step(X, Y, N, 1e-3, 0.9, 1.0, False, False, 11, step)
        D.data_parallel_step(b, X, Y, B, N, 1e-3, 0.9, 1.0, False, 11, step)
    assert torch.isfinite(a.theta_w).all()
    assert rel_err(b.theta_w, a.theta_w) < 1e-5
    assert rel_err(b.mom_w, a.mom_w) < 1e-4


@pytest.mark.parametrize("min_bucket", [1, 3000, 1 << 30])
def test_overlapped_reduction_is_bit_identical_to_the_single_all_reduce(min_bucket):
    """overlap=True: the per-layer hook of the reverse pass sums a layer's gradient slabs on a side stream (where the
    all-reduce of that bucket is issued when world > 1) while the layers below run; every element is still produced by the
    same fixed-order slab sum, so the step must equal the one-buffer path bit for bit (world = 1 here; the NCCL run is
    bench.py --gpus N: cfg5_dp.replicas_bit_identical / equals_single_all_reduce)."""
    spec = ModelSpec.build(7, 2, [256, 128, 256], [6, 5, 2], ["RBF", "RBF", "RBF"], True, False, "gaussian")
    engines = []
    for _ in range(2):
        torch.manual_seed(5)
        e = Engine(spec, 1, precision=_ffi.PREC_TF32)
        e.theta_w.normal_()
        e.theta_h[:, e.layout.off_lik_log_var] = -1.0
        engines.append(e)
    a, b = engines
    g = torch.Generator().manual_seed(6)
    B, N = 640, 5000.0
    X = torch.randn(B, 7, generator=g).cuda()
    Y = torch.randn(B, 2, generator=g).cuda()
    for step in range(3):
        lla = D.data_parallel_step(a, X, Y, B, N, 1e-3, 0.9, 1.0, False, 11, step, overlap=False).clone()
        llb = D.data_parallel_step(b, X, Y, B, N, 1e-3, 0.9, 1.0, False, 11, step, overlap=True,
                                   min_bucket_floats=min_bucket).clone()
        assert torch.equal(lla, llb)
    torch.cuda.synchronize()
    assert torch.isfinite(a.theta_w).all()
    assert torch.equal(a.theta_w, b.theta_w) and torch.equal(a.mom_w, b.mom_w)
    st = b._dp_overlap_state[1]
    assert 0 in st.flush_at and (len(st.flush_at) == 3) == (min_bucket == 1)


def test_layer_hook_errors_surface_in_the_caller():
    e = _engine(_ffi.PREC_TF32)
    X = torch.randn(64, 7).cuda()
    Y = torch.randn(64, 2).cuda()

    def bad_hook(l, finalize_layer):
        raise ValueError(f"layer {l}")

    with pytest.raises(ValueError, match="layer 1"):
        e.gradients(X, Y, 100.0, hyper=False, prior_w=False, prior_h=False, layer_hook=bad_hook)
    # the hook is cleared again: a plain call works and never reports layers
    ll, gW, _ = e.gradients(X, Y, 100.0, hyper=False, prior_w=False, prior_h=False)
    assert torch.isfinite(gW).all()
