"""Host-minibatch path of sgmcmc_update (dgprf_sgmcmc_step_host) and the random_fixed=False mode of the sampler.

* the reference's sgmcmc_update consumes its batch by value (models/dgp.py:184): a caller may drop the pinned batch
  right after the call -- the engine holds it until the GPU is done (engine.py:365);
* a softmax Y with more than one column is read as labels = Y[:, 0] (likelihoods/softmax.py:14) on the host path too;
* random_fixed=False redraws z inside every potential / step (layers/rf_layers.py:39-41)."""
import gc

import pytest
import torch

from helpers import make_model, rel_err

pytestmark = pytest.mark.gpu


def _pair(name):
    a, X, Y, c = make_model(name, seed=3)
    b, _, _, _ = make_model(name, seed=3)
    for m in (a, b):
        m.precond_update(None, c["N"], precond_type="identity")
        m.seed(77)
    b._engine.mom_w.copy_(a._engine.mom_w)
    assert torch.equal(a._engine.theta_w, b._engine.theta_w)
    return a, b, X, Y, c


def test_pinned_batch_may_be_dropped_after_the_call():
    a, b, X, Y, c = _pair("protein_small")
    for step in range(12):
        g = torch.Generator().manual_seed(step)
        Xs = X + 0.01 * torch.randn(X.shape, generator=g)
        a.sgmcmc_update(Xs.cuda(), Y.cuda(), c["N"], lr=0.01, momentum_decay=0.9)       # device path
        xp, yp = Xs.clone().pin_memory(), Y.clone().pin_memory()
        b.sgmcmc_update(xp, yp, c["N"], lr=0.01, momentum_decay=0.9)                    # host path, zero copy
        del xp, yp                                                                      # dropped while the GPU may still read
        gc.collect()
        junk = torch.full((X.shape[0], X.shape[1]), 1e6).pin_memory()                   # would reuse the freed pinned block
        del junk
    torch.cuda.synchronize()
    assert torch.equal(a._engine.theta_w, b._engine.theta_w)
    assert len(b._engine._inflight) <= 64


def test_softmax_host_batch_with_a_wide_label_matrix():
    a, b, X, Y, c = _pair("mnist_small")
    Ywide = torch.cat([Y, torch.full((Y.shape[0], 3), 7.0)], dim=1).contiguous()        # labels in column 0
    a.sgmcmc_update(X.cuda(), Y.cuda(), c["N"], lr=0.01, momentum_decay=0.0)
    b.sgmcmc_update(X.clone().pin_memory(), Ywide.pin_memory(), c["N"], lr=0.01, momentum_decay=0.0)
    torch.cuda.synchronize()
    assert torch.equal(a._engine.theta_w, b._engine.theta_w)
    with pytest.raises(AssertionError):
        r, Xr, Yr, cr = make_model("ragged_mixed")                                     # Gaussian, d_out = 2
        r.precond_update(None, cr["N"], precond_type="identity")
        r.sgmcmc_update(Xr.clone().pin_memory(), Yr[:, :1].contiguous().pin_memory(), cr["N"])   # too narrow


def test_random_fixed_false_redraws_z_inside_the_sampler():
    m, X, Y, c = make_model("protein_small", seed=5)
    m.precond_update(None, c["N"], precond_type="identity")
    Xd, Yd = X.cuda(), Y.cuda()
    z_before = [z.clone() for z in m._engine.z]
    u_fixed = [float(m.U(Xd, Yd, c["N"])) for _ in range(2)]
    assert u_fixed[0] == u_fixed[1]
    m.set_random_fixed(False)
    u_free = [float(m.U(Xd, Yd, c["N"])) for _ in range(3)]
    assert len(set(u_free)) == 3 and all(u != u_fixed[0] for u in u_free)               # a fresh draw per call
    # U agrees with log_likelihood's layer-wise resampling in distribution only; here: same order of magnitude
    w0 = m._engine.theta_w.clone()
    m.sgmcmc_update(Xd, Yd, c["N"], lr=0.01, momentum_decay=0.9)
    assert not torch.equal(w0, m._engine.theta_w) and torch.isfinite(m._engine.theta_w).all()
    U, g = m.grad_U(Xd, Yd, c["N"])
    assert all(torch.isfinite(t).all() for t in g.values())
    for z, keep in zip(m._engine.z, z_before):                                          # the fixed draw is restored
        assert torch.equal(z, keep)
    m.set_random_fixed(True)
    assert float(m.U(Xd, Yd, c["N"])) != u_fixed[0]                                     # W moved ...
    m._engine.theta_w.copy_(w0)
    assert float(m.U(Xd, Yd, c["N"])) == u_fixed[0]                                     # ... and nothing else did
