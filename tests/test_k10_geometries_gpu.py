"""The cluster-split step kernel (csrc/k10_step_cluster.cu) in EVERY geometry, not only the one its planner picks.

The planner chooses (rows per cluster, cluster size) from the batch and feature counts, so the small parity shapes all
land on single-CTA clusters; the exchange over distributed shared memory, the column split of a layer over 2 / 4 / 8
CTAs (including CTAs that own no column at all), the trainable-mean row-sum column, arc-cosine layers and the softmax
seed are forced here through DGPRF_K10_MT / DGPRF_K10_CL and checked against the fp64 oracle with injected noise
(models/dgp.py:184-216), fused update and stand-alone update alike, plus chain batches against single chains."""
import os

import pytest
import torch

import dgprf_oracle as O
from dgprf import _ffi
from dgprf.chains import ChainEnsemble
from helpers import CONFIGS, assert_close, make_model, oracle_params, rel_err

pytestmark = pytest.mark.gpu
GEOMS = [(1, 1), (2, 1), (1, 2), (2, 2), (1, 4), (2, 4), (1, 8), (2, 8)]
ELIGIBLE = [n for n, c in CONFIGS.items() if max(c["n_gp"]) <= 32]          # K10 takes n_gp <= 32 (wide_gp stays on K9)


@pytest.fixture
def geometry(request):
    mt, cl = request.param
    os.environ["DGPRF_K10_MT"], os.environ["DGPRF_K10_CL"] = str(mt), str(cl)
    yield mt, cl
    os.environ.pop("DGPRF_K10_MT", None)
    os.environ.pop("DGPRF_K10_CL", None)
    os.environ.pop("DGPRF_NO_FUSED_UPDATE", None)


@pytest.mark.parametrize("geometry", GEOMS, indirect=True, ids=[f"mt{m}cl{c}" for m, c in GEOMS])
@pytest.mark.parametrize("name", ELIGIBLE)
@pytest.mark.parametrize("mode", ["sghmc", "resample_unfused"])
def test_step_matches_the_oracle_in_every_geometry(geometry, name, mode):
    if mode == "resample_unfused":
        os.environ["DGPRF_NO_FUSED_UPDATE"] = "1"
    model, X, Y, c = make_model(name)
    model.precond_update(None, c["N"], precond_type="identity")
    e = model._engine
    p = oracle_params(model)
    names = e.names(False)
    g = torch.Generator().manual_seed(7)
    mom = {n: e.view(n, "mom").double().cpu().clone() for n in names}
    eps = {n: torch.randn(e.view(n).shape, generator=g, dtype=torch.float64) for n in names}
    res = {n: torch.randn(e.view(n).shape, generator=g, dtype=torch.float64) for n in names} if mode != "sghmc" else None
    _, _, p_new, m_new = O.sgmcmc_step(p, mom, X.double(), Y.double(), c["N"], lr=0.02, momentum_decay=0.9, temperature=1.0,
                                       eps=eps, resample=res)
    _ffi.profile_start()
    model.sgmcmc_update(X, Y, c["N"], lr=0.02, momentum_decay=0.9, temperature=1.0, eps=eps, resample=res)
    kernels = [nm for nm, _ in _ffi.profile_stop()]
    assert kernels[0] == "k10_step_cluster", kernels                          # the forced geometry really ran
    assert (len(kernels) == 1) == (mode == "sghmc"), kernels                  # fused update | K10 + K5
    new = dict(p_new.w_named())
    for n in names:
        assert rel_err(e.view(n), new[n]) < 1e-4, ("theta", n)
        assert rel_err(e.view(n, "mom"), m_new[n]) < 1e-4, ("moments", n)
        assert_close(e.view(n), new[n], what=f"theta {n}")
    live = torch.zeros_like(e.theta_w[0], dtype=torch.bool)
    for off, ln, _, _ in e.seg_w.values():
        live[off:off + ln] = True
    assert float(e.theta_w[0][~live].abs().sum()) == 0.0 and float(e.mom_w[0][~live].abs().sum()) == 0.0


@pytest.mark.parametrize("geometry", [(2, 4), (1, 2)], indirect=True, ids=["mt2cl4", "mt1cl2"])
def test_chain_batch_equals_single_chains_bit_for_bit(geometry):
    """Chain c of a batched launch (multi-wave grid, stand-alone update) == the same chain alone (fused update)."""
    kw = dict(d_in=9, d_out=1, n_hidden_layers=3, n_rf=128, n_gp=[9, 9, 1], input_cat=True, seed=11, precision="fp32")
    g = torch.Generator().manual_seed(0)
    X = torch.randn(6, 200, 9, generator=g).cuda()
    Y = torch.randn(6, 200, 1, generator=g).cuda()
    full = ChainEnsemble(n_chains=6, chain_base=0, **kw)
    one = ChainEnsemble(n_chains=1, chain_base=4, **kw)
    for step in range(3):
        full.sgmcmc_update(X, Y, 5000, lr=0.01, momentum_decay=0.9, resample_moments=(step == 1))
        one.sgmcmc_update(X[4:5], Y[4:5], 5000, lr=0.01, momentum_decay=0.9, resample_moments=(step == 1))
    assert torch.equal(full.engine.theta_w[4], one.engine.theta_w[0])
    assert torch.equal(full.engine.mom_w[4], one.engine.mom_w[0])
    assert not torch.equal(full.engine.theta_w[0], full.engine.theta_w[1])
