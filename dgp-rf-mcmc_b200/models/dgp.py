"""DGP_RF: random-feature deep GP with SG-MCMC over the GP weights (models/dgp.py:8-304).

Same constructor, properties and methods as the reference class; the arithmetic runs in
libdgprf's CUDA kernels through ``dgprf.engine.Engine``:

    U                 forward (K1) + likelihood (K3) + log-prior (K4)            dgp.py:161-182
    sgmcmc_update     K1 -> K3 -> K2 -> K5 in ONE C call, no Python per tensor   dgp.py:184-216
    precond_update    identity | RMSprop/Welford scalar mass per tensor (K6)     dgp.py:218-299

Extensions (keyword-only, default = reference behaviour): ``eps`` / ``resample`` dictionaries
inject the N(0,1) draws by parameter NAME for parity tests; ``grad_U`` exposes the gradients
that the reference obtains from tf.GradientTape.
"""
import contextlib
import math
import os

import numpy as np
import torch

from dgprf import _ffi
from dgprf.engine import Engine, ModelSpec
from dgprf.variable import Variable, out
from kernels import RBFKernel, ARCKernel
from layers import RBFLayer, ARCLayer, GPLayer
from likelihoods import Softmax, Gaussian
from utils import BNN_from_list, BNN_from_list_input_cat

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def _as_int_list(v, n, what):
    vals = [int(v)] * n if np.ndim(v) == 0 else [int(x) for x in np.asarray(v).reshape(-1)]
    assert len(vals) == n, f"Error in #{what} layers!"
    return vals


class DGP_RF:
    def __init__(self, d_in, d_out, n_hidden_layers=1, n_rf=20, n_gp=2, likelihood=None,
                 kernel_type_list=None, kernel_trainable=True, random_fixed=True, input_cat=False,
                 set_nonzero_mean=False, name=None):
        self.name = name
        self.d_in, self.d_out = int(d_in), int(d_out)
        self.n_hidden_layers = int(n_hidden_layers)
        self.random_fixed = random_fixed
        self.input_cat = input_cat
        self.set_nonzero_mean = set_nonzero_mean
        self.kernel_trainable = kernel_trainable
        # The reference's default `likelihood=Softmax()` is evaluated once at import and shared
        # between models; a fresh instance per model is created here instead.
        self.likelihood = self._default_likelihood() if likelihood is None else likelihood
        self.n_rf = _as_int_list(n_rf, self.n_hidden_layers, "random feature")
        self.n_gp = _as_int_list(n_gp, self.n_hidden_layers, "hidden GP")
        if kernel_type_list is None:
            self.kernel_type_list = ['RBF'] * self.n_hidden_layers
        else:
            assert len(kernel_type_list) == self.n_hidden_layers, "Kernel type list's length does not match!"
            self.kernel_type_list = list(kernel_type_list)
        self.kernel_list = self.transform_kernel_list()
        self.BNN = self.transformed_BNN()
        self._adopt()

    @staticmethod
    def _default_likelihood():
        return Softmax()

    # ---- construction (dgp.py:74-115) -----------------------------------------------------------
    def transform_kernel_list(self):
        widths = [self.d_in] + [g + (self.d_in if self.input_cat else 0) for g in self.n_gp[:-1]]
        table = {'RBF': RBFKernel, 'ARC': ARCKernel}
        kernels_ = []
        for width, kind in zip(widths, self.kernel_type_list):
            if kind not in table:
                raise NotImplementedError
            kernels_.append(table[kind](n_feature=width, trainable=self.kernel_trainable, is_ard=True,
                                        length_scale=None))
        return kernels_

    def transformed_BNN(self):
        chain = []
        for l, k in enumerate(self.kernel_list):
            if k.kernel_type == "RBF":
                rf = RBFLayer(k, self.n_rf[l], random_fixed=self.random_fixed, set_nonzero_mean=self.set_nonzero_mean)
            elif k.kernel_type == "ARC":
                rf = ARCLayer(k, self.n_rf[l], random_fixed=self.random_fixed, set_nonzero_mean=self.set_nonzero_mean)
            else:
                raise NotImplementedError
            chain.extend([rf, GPLayer(rf.n_rf, self.n_gp[l])])
        return BNN_from_list_input_cat(chain) if self.input_cat else BNN_from_list(chain)

    def _adopt(self):
        """Move every parameter into the engine's flat buffers and keep the Variables as views."""
        if isinstance(self.likelihood, Gaussian):
            lik = "gaussian"
        elif isinstance(self.likelihood, Softmax):
            lik = "softmax"
        else:
            raise NotImplementedError
        assert self.n_gp[-1] == self.d_out, "n_gp[-1] must equal d_out for the likelihood to line up"
        spec = ModelSpec.build(self.d_in, self.d_out, self.n_rf, self.n_gp, self.kernel_type_list, self.input_cat,
                               self.set_nonzero_mean, lik)
        rf_layers = [self.BNN.layers[2 * l] for l in range(self.n_hidden_layers)]
        dev = rf_layers[0].z.device
        self._engine = Engine(spec, 1, dev, z=[rf.z.unsqueeze(0) for rf in rf_layers])
        self._vars = {}
        trainable_h = []

        def adopt(name, var):
            var._rebind(self._engine.view(name))
            var._owner, var._seg = self, name
            self._vars[name] = var

        for l, rf in enumerate(rf_layers):
            adopt(f"W_{l}", self.BNN.layers[2 * l + 1].W)
            adopt(f"log_amp_{l}", rf.kernel.log_amplitude)
            adopt(f"log_inv_ls_{l}", rf.kernel.log_inv_length_scale)
            if rf.kernel.log_amplitude.trainable:
                trainable_h += [f"log_amp_{l}", f"log_inv_ls_{l}"]
            if self.set_nonzero_mean:
                adopt(f"mean_{l}", rf.mean)
                trainable_h.append(f"mean_{l}")
        if lik == "gaussian":
            adopt("lik_log_var", self.likelihood.lik_log_var)
            if self.likelihood.lik_log_var.trainable:
                trainable_h.append("lik_log_var")
        self._engine.trainable_h = trainable_h
        self.BNN._engine_owner = self
        self._seed = int.from_bytes(os.urandom(7), "little")
        self._step = 0
        self._sampler_ready = {}

    # hooks used by Variable.moments / Variable.M
    def _moment_view(self, var):
        return self._engine.view(var._seg, "mom")

    def _set_mass(self, var, mass):
        self._engine.set_mass(var._seg, mass)

    def set_precision(self, name):
        """'fp32' (SIMT, parity mode) | 'tf32' (tcgen05 tensor-core kernels).  Extension."""
        self._engine.set_precision(name)

    def seed(self, seed):
        """Seed of the in-kernel Philox noise (extension; the reference is unseeded)."""
        self._seed, self._step = int(seed), 0

    # ---- variable groups (dgp.py:54-72) -----------------------------------------------------------
    @property
    def trainable_variables(self):
        order = []
        for l in range(self.n_hidden_layers):
            order += [f"log_amp_{l}", f"log_inv_ls_{l}"]
        order.append("lik_log_var")
        for l in range(self.n_hidden_layers):
            order += [f"mean_{l}", f"W_{l}"]
        return [self._vars[n] for n in order if n in self._vars and self._vars[n].trainable]

    @property
    def Likelihood_hyperparams(self):
        return list(self.likelihood.trainable_variables)

    @property
    def Omega_hyperparams(self):
        params = []
        for l in range(self.n_hidden_layers):
            params.extend(self.BNN.layers[2 * l].trainable_variables)
        return params

    @property
    def W_mcmc(self):
        return [self.BNN.layers[2 * l + 1].W for l in range(self.n_hidden_layers)]

    def assign_W(self, W_value_list):
        for gp_layer, W_value in zip(self.BNN.gp_layers, W_value_list):
            gp_layer.assign_W(W_value)

    # ---- densities (dgp.py:118-182) ------------------------------------------------------------------
    def log_likelihood(self, X, Y, allow_gradient_from_W=True):
        """log p(y_i | x_i, params) per row -> [N]."""
        if not self.BNN._fused_ok():
            return self.likelihood.log_prob(self.BNN(X), Y)
        ll, _, _ = self._engine.evaluate(X, Y)
        return out(ll[0])

    def _log_prior_of(self, variables):
        tot = None
        for v in variables:
            t = self._engine.log_prior(v.tensor.reshape(-1))
            tot = t if tot is None else tot + t
        return tot

    def prior_W(self):
        e = self._engine
        pad = e.layout.w_len - sum(s[1] for s in e.seg_w.values())
        return out(e.log_prior(e.theta_w[0]) + pad * _HALF_LOG_2PI)     # zero padding adds only the constant

    def prior_kernel_params(self):
        vs = []
        for l in range(self.n_hidden_layers):
            k = self.BNN.layers[2 * l].kernel
            vs += [k.log_amplitude, k.log_inv_length_scale]
        return out(self._log_prior_of(vs))

    def prior_likelihood_params(self):
        if isinstance(self.likelihood, Softmax):
            return 0.
        if isinstance(self.likelihood, Gaussian):
            vs = self.likelihood.trainable_variables
            return out(self._log_prior_of(vs)) if vs else 0.
        raise NotImplementedError

    @contextlib.contextmanager
    def _z_for_call(self):
        """random_fixed=False (layers/rf_layers.py:39-41, 85-87): the reference draws a fresh z inside EVERY forward, also
        the one under the GradientTape of U / sgmcmc_update / precond_update.  The fused kernels read z from the engine
        buffers, so one fresh draw is put there for the duration of one potential / gradient / step evaluation (forward
        and backward of a step see the same draw, as under one tape) and the fixed draw is restored afterwards."""
        if self.BNN._fused_ok() or getattr(self, "_z_redrawn", False):
            yield
            return
        e = self._engine
        rf = [self.BNN.layers[2 * l] for l in range(self.n_hidden_layers)]
        saved = [z.clone() for z in e.z]
        for l, layer in enumerate(rf):
            if not layer.random_fixed:
                e.z[l].normal_()
        self._z_redrawn = True
        try:
            yield
        finally:
            self._z_redrawn = False
            for z, keep in zip(e.z, saved):
                z.copy_(keep)

    def U(self, X_batch, Y_batch, data_size, full_bayesian=False, allow_gradient_from_W=True):
        """Minibatch potential  -(log_prior / N + sum_i ll_i / B)."""
        with self._z_for_call():
            _, _, tot = self._engine.evaluate(X_batch, Y_batch)
        B = float(np.shape(X_batch)[0])
        N = float(data_size)
        if not full_bayesian:
            log_prior = self.prior_W() / N if allow_gradient_from_W else 0.
        else:
            assert allow_gradient_from_W == True, "Full Bayes should allow gradients from W!"
            log_prior = self.prior_W()
            hyp = [v for v in self.trainable_variables if not v._seg.startswith("W_")]
            if hyp:
                log_prior = log_prior + self._log_prior_of(hyp)
            log_prior = log_prior / N
        return out(-(log_prior + tot[0] / B))

    def grad_U(self, X_batch, Y_batch, data_size, full_bayesian=False, allow_gradient_from_W=True, hyper=None):
        """(U, {name: dU/dparam}) -- what tape.gradient(U, watched) yields (dgp.py:194-204; with
        allow_gradient_from_W=False and hyper=True: the M-step gradients, utils_training.py:341-354)."""
        hyper = full_bayesian if hyper is None else hyper
        e = self._engine
        with self._z_for_call():                       # U and its gradient share one draw of z
            tot, gW, gH = e.gradients(X_batch, Y_batch, data_size, hyper=hyper,
                                      prior_w=allow_gradient_from_W, prior_h=full_bayesian)
            g = {n: t.view(e.view(n).shape) for n, t in e.named_from_flat(gW, "w").items()}
            if hyper:
                g.update({n: t.view(e.view(n).shape) for n, t in e.named_from_flat(gH, "h").items()})
            return self.U(X_batch, Y_batch, data_size, full_bayesian, allow_gradient_from_W), g

    # ---- SG-MCMC (dgp.py:184-216) ---------------------------------------------------------------------
    def sgmcmc_update(self, X_batch, Y_batch, data_size, lr=0.01, momentum_decay=0.95,
                      resample_moments=False, temperature=1., full_bayesian=False, *, eps=None, resample=None,
                      u_host=None):
        """One SGHMC step (SGLD when momentum_decay == 0) on W (and on every trainable
        hyper-parameter when full_bayesian)."""
        if not self._sampler_ready.get(bool(full_bayesian), False):       # checked once, then cached
            watched = self.trainable_variables if full_bayesian else self.W_mcmc
            for param in watched:
                assert param._mom is not None, "Trainable Params do not have attr moments!"
                assert param._M is not None, "Trainable Params do not have attr preconditioner M!"
            self._sampler_ready[bool(full_bayesian)] = True
        e = self._engine
        self._step += 1
        if not self.BNN._fused_ok():                    # random_fixed=False: fresh z for this step (see _z_for_call)
            with self._z_for_call():
                return self._sgmcmc_update_impl(X_batch, Y_batch, data_size, lr, momentum_decay, resample_moments,
                                                temperature, full_bayesian, eps, resample, None)
        return self._sgmcmc_update_impl(X_batch, Y_batch, data_size, lr, momentum_decay, resample_moments, temperature,
                                        full_bayesian, eps, resample, u_host)

    def _sgmcmc_update_impl(self, X_batch, Y_batch, data_size, lr, momentum_decay, resample_moments, temperature,
                            full_bayesian, eps, resample, u_host):
        e = self._engine
        if (eps is None and resample is None and type(X_batch) is torch.Tensor and type(Y_batch) is torch.Tensor
                and not X_batch.is_cuda and X_batch.dtype is torch.float32 and Y_batch.dtype is torch.float32
                and X_batch.dim() == 2 and Y_batch.dim() == 2 and X_batch.is_contiguous() and Y_batch.is_contiguous()
                and e.device.type == "cuda"):
            # host minibatch: H2D copies + step enqueued by one C call (dgprf_sgmcmc_step_host)
            e.step_host(X_batch, Y_batch, float(data_size), float(lr), float(momentum_decay), float(temperature),
                        bool(resample_moments), bool(full_bayesian), self._seed, self._step, u_host=u_host)
            return
        inj = {}
        if eps is not None:
            inj["eps_w"] = e.flat_from_named(eps, "w")
            inj["eps_h"] = e.flat_from_named(eps, "h") if full_bayesian else None
        if resample is not None:
            inj["res_w"] = e.flat_from_named(resample, "w")
            inj["res_h"] = e.flat_from_named(resample, "h") if full_bayesian else None
            resample_moments = True
        e.step(X_batch, Y_batch, float(data_size), float(lr), float(momentum_decay), float(temperature),
               bool(resample_moments), bool(full_bayesian), self._seed, self._step, **inj)

    # ---- preconditioner (dgp.py:218-299) ------------------------------------------------------------------
    def precond_update(self, ds, data_size, K_batches=32, full_bayesian=False,
                       precond_type='rmsprop', second_moment_centered=False):
        """Attach `M` and `moments` to the sampled variables; 'rmsprop' re-estimates one scalar
        mass per tensor from the gradient noise of the first K_batches minibatches."""
        variables = self.trainable_variables if full_bayesian else self.W_mcmc
        e = self._engine
        for param in variables:
            if not hasattr(param, "M"):
                param.M = 1.
            if not hasattr(param, "moments"):
                param.moments = torch.randn(param.shape, device=e.device)
        if precond_type == 'identity':
            return None
        if precond_type != 'rmsprop':
            raise NotImplementedError
        L = _ffi.lib()
        st = _ffi.stream_ptr()
        for param in variables:
            param.m_c = param.moments.as_subclass(torch.Tensor) / math.sqrt(param.M)
        zeros = lambda n: torch.zeros(n, device=e.device, dtype=torch.float32)
        mean_w, m2_w = zeros(e.layout.w_len), zeros(e.layout.w_len)
        mean_h, m2_h = zeros(e.layout.h_len), zeros(e.layout.h_len)
        k = 0
        for X_batch, Y_batch in ds:
            with self._z_for_call():
                _, gW, gH = e.gradients(X_batch, Y_batch, data_size, hyper=full_bayesian, prior_w=True,
                                        prior_h=full_bayesian)
            k += 1
            _ffi.check(L.dgprf_welford_update(gW.data_ptr(), mean_w.data_ptr(), m2_w.data_ptr(), e.layout.w_len, k, st))
            if full_bayesian:
                _ffi.check(L.dgprf_welford_update(gH.data_ptr(), mean_h.data_ptr(), m2_h.data_ptr(), e.layout.h_len, k, st))
            if k == K_batches:
                break
        assert k == K_batches, f"Estimating M ends before we use {K_batches} batches, we actually use {k} batches!"
        est = {}
        for names, segs, mean, m2 in ((list(e.seg_w), e.seg_w, mean_w, m2_w),
                                      (e.trainable_h if full_bayesian else [], e.seg_h, mean_h, m2_h)):
            if not names:
                continue
            table = _ffi.make_segments([tuple(segs[n]) for n in names])
            mass = torch.empty(len(names), device=e.device, dtype=torch.float32)
            _ffi.check(L.dgprf_mass_estimate(mean.data_ptr(), m2.data_ptr(), table, len(names), int(K_batches),
                                             int(bool(second_moment_centered)), mass.data_ptr(), st))
            est.update(dict(zip(names, mass.cpu().tolist())))
        mass_min = min(est[p._seg] for p in variables)
        for param in variables:
            param.mean_pre = out(e.named_from_flat((mean_w if param._seg in e.seg_w else mean_h)[None], "w" if param._seg in e.seg_w else "h")[param._seg].view(param.shape))
            param.m2_pre = out(e.named_from_flat((m2_w if param._seg in e.seg_w else m2_h)[None], "w" if param._seg in e.seg_w else "h")[param._seg].view(param.shape))
            param.mass_estimate = est[param._seg]
            param.M = est[param._seg] / mass_min
            param.moments = math.sqrt(param.M) * param.m_c
        return None

    def set_random_fixed(self, state):
        for l in range(self.n_hidden_layers):
            self.BNN.layers[2 * l].set_random_fixed(state)
