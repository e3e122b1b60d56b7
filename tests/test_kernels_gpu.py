"""GPU tests of the stand-alone kernels: Philox noise (K5), predictive reduce (K7), Adam,
log-prior (K4), plus size-independent properties at BASELINE.json's full sizes."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

import dgprf_oracle as O
from dgprf import _ffi
from dgprf.engine import Engine, ModelSpec
from helpers import rel_err

pytestmark = pytest.mark.gpu


def _philox(n, seed, chain, step, stream):
    o = torch.empty(n, device="cuda")
    _ffi.check(_ffi.lib().dgprf_philox_normal(o.data_ptr(), n, seed, chain, step, stream, _ffi.stream_ptr()))
    return o


def test_philox_matches_host_restatement_and_is_standard_normal():
    from philox_ref import normal4
    got = _philox(64, 2024, 3, 17, 0).cpu().numpy()
    ref = np.array([normal4(2024, 3, i, 17, 0) for i in range(16)]).reshape(-1)
    assert np.allclose(got, ref, rtol=2e-4, atol=2e-5)
    x = _philox(1 << 22, 99, 0, 1, 0).double()
    n = x.numel()
    assert abs(float(x.mean())) < 5 / math.sqrt(n)
    assert abs(float(x.var()) - 1.0) < 5 * math.sqrt(2.0 / n)
    assert abs(float((x ** 3).mean())) < 5 * math.sqrt(15.0 / n)
    assert abs(float((x ** 4).mean()) - 3.0) < 5 * math.sqrt(96.0 / n)
    # streams / chains / steps are decorrelated
    for other in (_philox(1 << 22, 99, 0, 1, 1), _philox(1 << 22, 99, 1, 1, 0), _philox(1 << 22, 99, 0, 2, 0)):
        assert abs(float((x * other.double()).mean())) < 5 / math.sqrt(n)
    assert torch.equal(_philox(4096, 5, 2, 9, 0), _philox(4096, 5, 2, 9, 0))      # pure function of its key


def test_update_kernel_in_kernel_noise_statistics():
    """theta' - theta = h/M * m' with m' = beta m - hN g + sqrt(2(1-beta)T M) eps, eps from Philox."""
    n = 1 << 20
    theta = torch.zeros(1, n, device="cuda"); mom = torch.zeros(1, n, device="cuda"); grad = torch.zeros(1, n, device="cuda")
    segs = _ffi.make_segments([(0, n // 2, 1.0, 0), (n // 2, n // 2, 4.0, 0)])
    lr, N, beta, T = 0.01, 1000.0, 0.9, 1.0
    _ffi.check(_ffi.lib().dgprf_sgmcmc_update(theta.data_ptr(), mom.data_ptr(), n, n, 1, grad.data_ptr(), n, 1, 0,
                                              segs, 2, lr, N, beta, T, 0, 1234, 1, None, None, _ffi.stream_ptr()))
    m = mom[0].double()
    s0 = float(m[: n // 2].std()); s1 = float(m[n // 2:].std())
    assert s0 == pytest.approx(math.sqrt(2 * (1 - beta) * T * 1.0), rel=0.01)
    assert s1 == pytest.approx(math.sqrt(2 * (1 - beta) * T * 4.0), rel=0.01)
    h = math.sqrt(lr / N)
    assert rel_err(theta[0, : n // 2], h * m[: n // 2]) < 1e-5
    assert rel_err(theta[0, n // 2:], h / 4.0 * m[n // 2:]) < 1e-5
    # resample: momentum replaced by a fresh N(0,1) draw before the step
    mom.fill_(100.0)
    _ffi.check(_ffi.lib().dgprf_sgmcmc_update(theta.data_ptr(), mom.data_ptr(), n, n, 1, grad.data_ptr(), n, 1, 0,
                                              segs, 2, lr, N, beta, 0.0, 1, 1234, 2, None, None, _ffi.stream_ptr()))
    assert float(mom[0].double().std()) == pytest.approx(beta, rel=0.01) and abs(float(mom.mean())) < 0.01


def test_update_kernel_full_size_linearity():
    """cfg5-sized flat buffer (991232 parameters x 4 chains): T=0 update is linear in the gradient."""
    n, Cn = 991232, 4
    g = torch.Generator(device="cuda").manual_seed(0)
    th0 = torch.randn(Cn, n, device="cuda", generator=g); m0 = torch.randn(Cn, n, device="cuda", generator=g)
    g1 = torch.randn(Cn, n, device="cuda", generator=g); g2 = torch.randn(Cn, n, device="cuda", generator=g)
    segs = _ffi.make_segments([(0, n, 1.0, 1)])

    def run(grad):
        th, m = th0.clone(), m0.clone()
        _ffi.check(_ffi.lib().dgprf_sgmcmc_update(th.data_ptr(), m.data_ptr(), n, n, Cn, grad.data_ptr(), n, 1, 0,
                                                  segs, 1, 0.01, 515345.0, 0.9, 0.0, 0, 1, 1, None, None, _ffi.stream_ptr()))
        return th, m
    (ta, ma), (tb, mb), (tc, mc) = run(g1), run(g2), run(0.5 * (g1 + g2))
    assert rel_err(tc, 0.5 * (ta.double() + tb.double())) < 1e-6
    ref_m = 0.9 * m0.double() - math.sqrt(0.01 / 515345.0) * 515345.0 * (g1.double() + th0.double() / 515345.0)
    assert rel_err(ma, ref_m) < 1e-5


@pytest.mark.parametrize("S,N", [(1, 1), (7, 33), (100, 4573), (16, 51535)])
def test_predictive_reduce(S, N):
    g = torch.Generator().manual_seed(S * 1000 + N)
    lp = (3.0 * torch.randn(S, N, generator=g) - 2.0)
    se = torch.rand(S, N, generator=g)
    a_ref, b_ref = O.predictive_average(lp.double(), se.double())
    lpd, sed = lp.cuda(), se.cuda()
    outv = torch.empty(2, device="cuda"); lse = torch.empty(N, device="cuda")
    scratch = torch.empty(2 * ((N + 255) // 256) + 2, device="cuda")
    L = _ffi.lib()
    _ffi.check(L.dgprf_predictive_reduce(lpd.data_ptr(), sed.data_ptr(), S, N, N, math.log(S), 1, lse.data_ptr(),
                                         outv.data_ptr(), scratch.data_ptr(), _ffi.stream_ptr()))
    assert float(outv[0]) == pytest.approx(float(a_ref), rel=1e-5, abs=1e-6)
    assert float(outv[1]) == pytest.approx(float(b_ref), rel=1e-5)
    assert rel_err(lse, torch.logsumexp(lp.double(), 0)) < 1e-5
    # sharded sample sets: logsumexp of per-shard logsumexps gives the same answer
    if S >= 2:
        h = S // 2
        parts = torch.empty(2, N, device="cuda")
        for i, (lo, hi) in enumerate(((0, h), (h, S))):
            _ffi.check(L.dgprf_predictive_reduce(lpd[lo:hi].contiguous().data_ptr(), None, hi - lo, N, N, 0.0, 1,
                                                 parts[i].data_ptr(), outv.data_ptr(), scratch.data_ptr(), _ffi.stream_ptr()))
        _ffi.check(L.dgprf_predictive_reduce(parts.data_ptr(), None, 2, N, N, math.log(S), 1, None, outv.data_ptr(),
                                             scratch.data_ptr(), _ffi.stream_ptr()))
        assert float(outv[0]) == pytest.approx(float(a_ref), rel=1e-5, abs=1e-6)


def test_adam_and_log_prior():
    g = torch.Generator().manual_seed(0)
    th = torch.randn(577, generator=g); gr = torch.randn(577, generator=g)
    m = torch.zeros(577); v = torch.zeros(577)
    thd, grd, md, vd = th.cuda(), gr.cuda(), m.cuda(), v.cuda()
    th64, m64, v64 = th.double(), m.double(), v.double()
    for t in (1, 2, 3):
        _ffi.check(_ffi.lib().dgprf_adam_step(thd.data_ptr(), grd.data_ptr(), md.data_ptr(), vd.data_ptr(), 577,
                                              0.01, 0.9, 0.999, 1e-7, t, _ffi.stream_ptr()))
        th64, m64, v64 = O.adam_step(th64, gr.double(), m64, v64, t)
    assert rel_err(thd, th64) < 1e-5
    x = torch.randn(3, 1000, generator=g).cuda()
    o = torch.empty(3, device="cuda")
    _ffi.check(_ffi.lib().dgprf_log_prior(x.data_ptr(), 1000, 1000, 3, o.data_ptr(), _ffi.stream_ptr()))
    assert rel_err(o, O.log_gaussian(x.double().cpu()).sum(-1)) < 1e-5


def test_full_size_protein_step_properties():
    """BASELINE.json configs[1] at full size (N=45730, D=9, M=512, B=1000): size-independent
    properties instead of an oracle run -- determinism, chain batching == single chains, and
    dU/dW consistent with a finite difference of U along a random direction."""
    spec = ModelSpec.build(9, 1, [512] * 3, [9, 9, 1], ["RBF"] * 3, True, False, "gaussian")
    torch.manual_seed(0)
    Cn, B, N = 3, 1000, 45730.0
    e = Engine(spec, Cn, shared_z=False)
    e.theta_w.normal_()
    for off, ln, _, _ in e.seg_w.values():
        pass
    live = torch.zeros(e.layout.w_len, dtype=torch.bool, device="cuda")
    for off, ln, _, _ in e.seg_w.values():
        live[off:off + ln] = True
    e.theta_w[:, ~live] = 0
    for l, s in enumerate(spec.layers):
        e.view(f"log_inv_ls_{l}").fill_(-0.5 * math.log(s.d))
        for c in range(1, Cn):
            e.view(f"log_inv_ls_{l}", chain=c).fill_(-0.5 * math.log(s.d))
    e.theta_h[:, e.layout.off_lik_log_var] = math.log(0.1)
    X = torch.randn(Cn, B, 9, device="cuda"); Y = torch.randn(Cn, B, 1, device="cuda")
    tot, gW, _ = e.gradients(X, Y, N, hyper=False, prior_w=True, prior_h=False)
    tot2, gW2, _ = e.gradients(X, Y, N, hyper=False, prior_w=True, prior_h=False)
    assert torch.equal(gW, gW2) and torch.equal(tot, tot2)                   # bit-deterministic
    # chain c of the batch == a single-chain engine holding chain c's state
    c = 1
    e1 = Engine(spec, 1, z=[z[c:c + 1].clone() for z in e.z])
    e1.theta_w.copy_(e.theta_w[c:c + 1]); e1.theta_h.copy_(e.theta_h[c:c + 1])
    t1, g1, _ = e1.gradients(X[c], Y[c], N, hyper=False, prior_w=True, prior_h=False)
    assert torch.equal(g1[0], gW[c]) and torch.equal(t1[0], tot[c])
    # directional derivative (fp32 forward differences are noisy: loose tolerance)
    v = torch.zeros_like(e1.theta_w); v[0, live] = torch.randn(int(live.sum()), device="cuda")
    v /= v.norm()
    def Uval():
        _, _, t = e1.evaluate(X[c], Y[c])
        return -float(t[0]) / B + 0.5 * float((e1.theta_w.double() ** 2).sum()) / N
    base = e1.theta_w.clone(); h = 1e-2
    e1.theta_w.copy_(base + h * v); up = Uval()
    e1.theta_w.copy_(base - h * v); dn = Uval()
    fd = (up - dn) / (2 * h)
    assert fd == pytest.approx(float((g1.double() * v.double()).sum()), rel=2e-2, abs=1e-4)
