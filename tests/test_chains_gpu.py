"""Chain batching / sharding: 4 chains in one process == 2 + 2 chains in two "ranks", bit for bit, with the
in-kernel Philox noise switched on ("chain sharding == serial chains bit-for-bit")."""
import pytest
import torch

from dgprf.chains import ChainEnsemble

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_sharded_chains_equal_one_ensemble(precision):
    kw = dict(d_in=9, d_out=1, n_hidden_layers=3, n_rf=128, n_gp=[9, 9, 1], input_cat=True, seed=11, precision=precision)
    g = torch.Generator().manual_seed(0)
    X = torch.randn(4, 200, 9, generator=g).cuda()
    Y = torch.randn(4, 200, 1, generator=g).cuda()
    full = ChainEnsemble(n_chains=4, chain_base=0, **kw)
    lo = ChainEnsemble(n_chains=2, chain_base=0, **kw)
    hi = ChainEnsemble(n_chains=2, chain_base=2, **kw)
    for step in range(3):
        resample = step == 1
        full.sgmcmc_update(X, Y, 5000, lr=0.01, momentum_decay=0.9, resample_moments=resample)
        lo.sgmcmc_update(X[:2], Y[:2], 5000, lr=0.01, momentum_decay=0.9, resample_moments=resample)
        hi.sgmcmc_update(X[2:], Y[2:], 5000, lr=0.01, momentum_decay=0.9, resample_moments=resample)
    assert torch.equal(full.engine.theta_w[:2], lo.engine.theta_w)
    assert torch.equal(full.engine.theta_w[2:], hi.engine.theta_w)
    assert torch.equal(full.engine.mom_w[2:], hi.engine.mom_w)
    assert not torch.equal(full.engine.theta_w[0], full.engine.theta_w[1])      # chains really are independent
    ll, se = full.evaluate(X[0], Y[0])                                          # shared test set for all chains
    ll_hi, _ = hi.evaluate(X[0], Y[0])
    assert ll.shape == (4, 200) and torch.equal(ll[2:], ll_hi)
