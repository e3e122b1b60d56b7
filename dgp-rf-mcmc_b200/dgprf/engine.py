"""Host-side engine: flat HBM layout of one DGP (or a batch of C independent chains), the
``dgprf_model`` descriptor, workspace cache, and the per-step C-ABI calls.

Layout (all fp32, every tensor starts on a 16-byte boundary so K5 can use 128-bit lanes):

    W buffer      [C][w_len]   W_0 | W_1 | ... | W_{L-1}            (sampled parameters)
    hyper buffer  [C][h_len]   per layer: log_amp(4) | log_inv_ls(d_l) | [mean(d_l)] ; lik_log_var(4)
    momentum buffers mirror both; z_l [C?][d_l, M_l] are fixed draws and live outside.

The padding between tensors is kept at exactly zero by every kernel.
"""
from __future__ import annotations

import os

import collections
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _ffi
from .variable import out


# How pinned host minibatches reach the kernels (read once, at import: these are A/B switches, not per-call options):
# 2 = copied on a side stream under the previous step's kernels (default), 1 = read in place over the bus
# (DGPRF_ZERO_COPY_READ), 0 = in-order staging copies (DGPRF_NO_ZERO_COPY).
_ZC_MODE = 0 if os.environ.get("DGPRF_NO_ZERO_COPY") else (1 if os.environ.get("DGPRF_ZERO_COPY_READ") else 2)


def _r4(n: int) -> int:
    return (n + 3) // 4 * 4


@dataclass
class LayerSpec:
    kind: str          # 'RBF' | 'ARC'
    d_prev: int
    d_x: int
    M: int
    g: int
    has_mean: bool = False

    @property
    def d(self):
        return self.d_prev + self.d_x

    @property
    def F(self):
        return 2 * self.M if self.kind == "RBF" else self.M


@dataclass
class ModelSpec:
    d_in: int
    d_out: int
    layers: List[LayerSpec]
    likelihood: str    # 'gaussian' | 'softmax'

    @staticmethod
    def build(d_in, d_out, n_rf: Sequence[int], n_gp: Sequence[int], kinds: Sequence[str], input_cat: bool,
              has_mean: bool, likelihood: str) -> "ModelSpec":
        """Shape wiring of models/dgp.py:74-115: d_0 = d_in, d_l = n_gp[l-1] (+ d_in if input_cat)."""
        layers = []
        for l, (M, g, k) in enumerate(zip(n_rf, n_gp, kinds)):
            d_prev = 0 if l == 0 else int(n_gp[l - 1])
            d_x = d_in if (l == 0 or input_cat) else 0
            layers.append(LayerSpec(k, d_prev, d_x, int(M), int(g), has_mean))
        return ModelSpec(int(d_in), int(d_out), layers, likelihood)


@dataclass
class FlatLayout:
    off_W: List[int] = field(default_factory=list)
    off_log_amp: List[int] = field(default_factory=list)
    off_log_inv_ls: List[int] = field(default_factory=list)
    off_mean: List[int] = field(default_factory=list)
    off_lik_log_var: int = -1
    w_len: int = 0
    h_len: int = 0

    @staticmethod
    def of(spec: ModelSpec) -> "FlatLayout":
        lay = FlatLayout()
        cw = ch = 0
        for s in spec.layers:
            lay.off_W.append(cw)
            cw += _r4(s.F * s.g)
            lay.off_log_amp.append(ch); ch += 4
            lay.off_log_inv_ls.append(ch); ch += _r4(s.d)
            if s.has_mean:
                lay.off_mean.append(ch); ch += _r4(s.d)
            else:
                lay.off_mean.append(-1)
        if spec.likelihood == "gaussian":
            lay.off_lik_log_var = ch; ch += 4
        lay.w_len, lay.h_len = cw, max(ch, 4)
        return lay


class Engine:
    """Owns the flat buffers of C chains of one architecture and drives libdgprf."""

    def __init__(self, spec: ModelSpec, n_chains: int = 1, device: Optional[torch.device] = None,
                 precision: Optional[int] = None, z: Optional[List[torch.Tensor]] = None, shared_z: bool = True):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
                else torch.device("cpu")
        if precision is None:
            from . import default_precision
            precision = default_precision()
        self.spec, self.C, self.device, self.precision = spec, int(n_chains), device, precision
        self.layout = FlatLayout.of(spec)
        f32 = dict(device=device, dtype=torch.float32)
        self.theta_w = torch.zeros(self.C, self.layout.w_len, **f32)
        self.mom_w = torch.zeros(self.C, self.layout.w_len, **f32)
        self.theta_h = torch.zeros(self.C, self.layout.h_len, **f32)
        self.mom_h = torch.zeros(self.C, self.layout.h_len, **f32)
        if z is None:
            zc = 1 if shared_z else self.C
            z = [torch.randn(zc, s.d, s.M, **f32) for s in spec.layers]
        self.z = z
        self._ws: Dict[Tuple[int, int], torch.Tensor] = {}
        self._staging: Dict = {}
        self._inflight = collections.deque()     # host minibatches the GPU may still be reading: (event | None, tensors)
        self._held = 0
        self._u_pinned = None
        self._model = None
        self._scratch = torch.zeros(max(4, self.C), **f32)
        # segment tables: name -> (offset, length, mass, flags)
        self.seg_w: Dict[str, List] = {f"W_{l}": [self.layout.off_W[l], s.F * s.g, 1.0, 1]
                                       for l, s in enumerate(spec.layers)}
        self.seg_h: Dict[str, List] = {}
        for l, s in enumerate(spec.layers):
            self.seg_h[f"log_amp_{l}"] = [self.layout.off_log_amp[l], 1, 1.0, 1]
            self.seg_h[f"log_inv_ls_{l}"] = [self.layout.off_log_inv_ls[l], s.d, 1.0, 1]
            if s.has_mean:
                self.seg_h[f"mean_{l}"] = [self.layout.off_mean[l], s.d, 1.0, 1]
        if spec.likelihood == "gaussian":
            self.seg_h["lik_log_var"] = [self.layout.off_lik_log_var, 1, 1.0, 1]
        self.trainable_h: List[str] = list(self.seg_h.keys())
        self._seg_cache = None

    # ---- views ---------------------------------------------------------------------------------
    def view(self, name: str, buf: str = "theta", chain: int = 0) -> torch.Tensor:
        """View of one named tensor inside the flat buffers (buf: 'theta' | 'mom')."""
        if name in self.seg_w:
            base = self.theta_w if buf == "theta" else self.mom_w
            off, n = self.seg_w[name][0], self.seg_w[name][1]
            l = int(name.split("_")[1])
            return base[chain, off:off + n].view(self.spec.layers[l].F, self.spec.layers[l].g)
        base = self.theta_h if buf == "theta" else self.mom_h
        off, n = self.seg_h[name][0], self.seg_h[name][1]
        v = base[chain, off:off + n]
        if name.startswith("log_amp") or name == "lik_log_var":
            return v.view(())
        if name.startswith("mean"):
            return v.view(n, 1)
        return v

    def names(self, full_bayesian: bool) -> List[str]:
        return list(self.seg_w.keys()) + (list(self.trainable_h) if full_bayesian else [])

    def flat_from_named(self, named: Dict[str, torch.Tensor], which: str) -> torch.Tensor:
        """Scatter {name: tensor} into a zero flat buffer laid out like the W / hyper buffer."""
        segs, n = (self.seg_w, self.layout.w_len) if which == "w" else (self.seg_h, self.layout.h_len)
        flat = torch.zeros(self.C, n, device=self.device, dtype=torch.float32)
        for name, t in named.items():
            if name in segs:
                off, ln = segs[name][0], segs[name][1]
                t = _ffi.as_dev(t, self.device).reshape(-1, ln)
                flat[:, off:off + ln] = t if t.shape[0] == self.C else t.expand(self.C, ln)
        return flat

    def named_from_flat(self, flat: torch.Tensor, which: str, chain: int = 0) -> Dict[str, torch.Tensor]:
        segs = self.seg_w if which == "w" else self.seg_h
        return {n: flat[chain, s[0]:s[0] + s[1]].clone() for n, s in segs.items()}

    def set_precision(self, name: str):
        self.precision = {"fp32": _ffi.PREC_FP32, "tf32": _ffi.PREC_TF32}[name]
        self._model = None
        self._ws.clear()            # the workspace layout depends on the precision mode
        self._staging.clear()

    def set_mass(self, name: str, mass: float):
        (self.seg_w if name in self.seg_w else self.seg_h)[name][2] = float(mass)
        self._seg_cache = None

    def _segments(self):
        if self._seg_cache is None:
            sw = _ffi.make_segments([tuple(v) for v in self.seg_w.values()])
            hv = [tuple(self.seg_h[n]) for n in self.trainable_h]
            sh = _ffi.make_segments(hv)
            self._seg_cache = (sw, len(self.seg_w), sh, len(hv))
        return self._seg_cache

    # ---- descriptor / workspace ------------------------------------------------------------------
    def model(self, w_base: Optional[torch.Tensor] = None, w_cs: Optional[int] = None,
              n_chains: Optional[int] = None, h_cs: Optional[int] = None) -> _ffi.Model:
        """The dgprf_model descriptor.  Overrides let stored W samples act as chains that share
        this model's hyper-parameters (stochastic-EM M-step, predictive sample sets)."""
        default = w_base is None and n_chains is None
        if default and self._model is not None:
            return self._model
        m = _ffi.Model()
        sp, lay = self.spec, self.layout
        m.n_layers, m.d_in, m.d_out = len(sp.layers), sp.d_in, sp.d_out
        m.likelihood = _ffi.LIK_GAUSSIAN if sp.likelihood == "gaussian" else _ffi.LIK_SOFTMAX
        m.n_chains = self.C if n_chains is None else int(n_chains)
        m.precision = self.precision
        m.w_base = (self.theta_w if w_base is None else w_base).data_ptr()
        m.w_cs = lay.w_len if w_cs is None else int(w_cs)
        m.h_base = self.theta_h.data_ptr()
        m.h_cs = lay.h_len if h_cs is None else int(h_cs)
        m.off_lik_log_var = lay.off_lik_log_var
        for l, s in enumerate(sp.layers):
            y = m.layer[l]
            y.kind = _ffi.KIND_RBF if s.kind == "RBF" else _ffi.KIND_ARC
            y.d_prev, y.d_x, y.M, y.g, y.has_mean = s.d_prev, s.d_x, s.M, s.g, int(s.has_mean)
            y.off_W, y.off_log_amp, y.off_log_inv_ls = lay.off_W[l], lay.off_log_amp[l], lay.off_log_inv_ls[l]
            y.off_mean = lay.off_mean[l]
            y.z = self.z[l].data_ptr()
            y.z_cs = 0 if (self.z[l].shape[0] == 1 or not default) else s.d * s.M
        if default:
            self._model = m
        return m

    def workspace(self, m: _ffi.Model, B: int, mode: int) -> torch.Tensor:
        key = (int(B), int(mode), int(m.n_chains))
        ws = self._ws.get(key)
        if ws is None:
            _ffi.require_cuda()
            nbytes = C.c_size_t(0)
            _ffi.check(_ffi.lib().dgprf_workspace_bytes(C.byref(m), B, mode, C.byref(nbytes)))
            if len(self._ws) >= 6:                       # keep the cache bounded
                self._ws.pop(next(iter(self._ws)))
            ws = torch.zeros(max(nbytes.value, 256), dtype=torch.uint8, device=self.device)   # zero once
            self._ws[key] = ws
        return ws

    # ---- tensor plumbing ---------------------------------------------------------------------------
    def _xy(self, X, Y=None, n_chains=None):
        """X: [B, d_in] shared by all chains or [C, B, d_in]; returns (X, x_cs, Y, y_cs, B)."""
        Cn = self.C if n_chains is None else n_chains
        X = _ffi.as_dev(X, self.device)
        if X.ndim == 2:
            x_cs, B = 0, X.shape[0]
        else:
            assert X.shape[0] == Cn, "leading dim of X must be the chain count"
            x_cs, B = X.shape[1] * X.shape[2], X.shape[1]
        assert X.shape[-1] == self.spec.d_in, f"X has width {X.shape[-1]}, model expects {self.spec.d_in}"
        y_cs = 0
        if Y is not None:
            Y = _ffi.as_dev(Y, self.device)
            yw = self.spec.d_out if self.spec.likelihood == "gaussian" else 1
            if Y.ndim == 1:
                Y = Y[:, None]
            if self.spec.likelihood == "softmax" and Y.shape[-1] != 1:
                Y = Y[..., :1].contiguous()
            assert Y.shape[-1] == yw and Y.shape[-2] == B, "Y shape does not match the likelihood / batch"
            if Y.ndim == 3:
                y_cs = Y.shape[1] * Y.shape[2]
        return X, x_cs, Y, y_cs, B

    # ---- hot path ------------------------------------------------------------------------------------
    def forward(self, X, m: Optional[_ffi.Model] = None, mode: int = _ffi.MODE_EVAL, want_F: bool = True):
        """BNN(X): [C, B, d_out] (utils.py:10-16, 32-44).  mode=TRAIN also saves the features."""
        m = self.model() if m is None else m
        X, x_cs, _, _, B = self._xy(X, None, m.n_chains)
        ws = self.workspace(m, B, mode)
        F = torch.empty(m.n_chains, B, self.spec.d_out, device=self.device, dtype=torch.float32) if want_F else None
        _ffi.check(_ffi.lib().dgprf_forward(C.byref(m), X.data_ptr(), x_cs, B, mode, ws.data_ptr(),
                                            ws.numel(), _ffi.ptr(F), _ffi.stream_ptr()))
        return F

    def evaluate(self, X, Y, m: Optional[_ffi.Model] = None):
        """Per-point log-likelihood and (squared error | correct flag): ([C,B], [C,B], ll_sum [C])."""
        m = self.model() if m is None else m
        X, x_cs, Y, y_cs, B = self._xy(X, Y, m.n_chains)
        ws = self.workspace(m, B, _ffi.MODE_EVAL)
        L = _ffi.lib()
        st = _ffi.stream_ptr()
        ll = torch.empty(m.n_chains, B, device=self.device, dtype=torch.float32)
        aux = torch.empty(m.n_chains, B, device=self.device, dtype=torch.float32)
        tot = torch.empty(m.n_chains, device=self.device, dtype=torch.float32)
        _ffi.check(L.dgprf_forward(C.byref(m), X.data_ptr(), x_cs, B, _ffi.MODE_EVAL, ws.data_ptr(), ws.numel(),
                                   None, st))
        _ffi.check(L.dgprf_loglik(C.byref(m), Y.data_ptr(), y_cs, B, _ffi.MODE_EVAL, ws.data_ptr(), ws.numel(),
                                  ll.data_ptr(), aux.data_ptr(), tot.data_ptr(), 0.0, st))
        return ll, aux, tot

    def gradients(self, X, Y, data_size: float, hyper: bool, prior_w: bool, prior_h: bool,
                  m: Optional[_ffi.Model] = None, inv_B: Optional[float] = None, out_flat: Optional[torch.Tensor] = None,
                  fused: bool = True, layer_hook=None):
        """Forward + likelihood seed + backward; returns (ll_sum [C], gW [C,w_len], gH [C,h_len]|None).
        layer_hook(l, finalize_layer): called on the host as soon as the reverse pass of layer l is enqueued (top layer
        first; layered kernels only); finalize_layer(l, cuda_stream) sums that layer's partial slabs into gW on the stream
        given.  The hook owns the ordering against the current stream (dgprf/dist.py: data_parallel_step).
        gW/gH are dU/dtheta of models/dgp.py:161-182 (prior terms theta/N added on request).  ONE C call (dgprf_gradients):
        W-only gradients of a model the row-fused step kernel takes run in one launch (fused=False keeps the layered kernels)."""
        assert prior_w or not prior_h, "hyper prior without W prior is not a mode of the reference"
        m = self.model() if m is None else m
        X, x_cs, Y, y_cs, B = self._xy(X, Y, m.n_chains)
        mode = _ffi.MODE_HYPER if hyper else _ffi.MODE_TRAIN
        ws = self.workspace(m, B, mode)
        L = _ffi.lib()
        st = _ffi.stream_ptr()
        Cn = m.n_chains
        if out_flat is not None:
            # data-parallel step: gW and sum_i ll_i land in ONE flat buffer [w_len + 1] (the all-reduce payload); inv_B is the
            # 1 / B_global the likelihood seed is scaled by (models/dgp.py:174 over the GLOBAL minibatch)
            assert Cn == 1 and out_flat.numel() == self.layout.w_len + 1 and out_flat.is_contiguous()
            gW = out_flat[:self.layout.w_len].view(1, -1)
            tot = out_flat[self.layout.w_len:]
        else:
            tot = torch.empty(Cn, device=self.device, dtype=torch.float32)
            gW = torch.empty(Cn, self.layout.w_len, device=self.device, dtype=torch.float32)
        gH = torch.empty(Cn, self.layout.h_len, device=self.device, dtype=torch.float32) if hyper else None
        inv_N = 1.0 / float(data_size)
        if layer_hook is None:
            _ffi.check(L.dgprf_gradients(C.byref(m), X.data_ptr(), x_cs, Y.data_ptr(), y_cs, B, mode, ws.data_ptr(), ws.numel(),
                                         gW.data_ptr(), gW.shape[1], gH.data_ptr() if hyper else None, self.layout.h_len,
                                         inv_N if (prior_w or prior_h) else 0.0, int(prior_h), tot.data_ptr(),
                                         0.0 if inv_B is None else float(inv_B), int(bool(fused)), st))
            return tot, gW, gH
        # layered reverse pass with a per-layer host hook (data-parallel step): the caller finalizes gW slice by slice
        # (finalize_layer below) as the layers retire, so no whole-buffer finalize here (gW = NULL)
        assert not hyper, "the per-layer hook drives W-only gradients"
        prior = inv_N if prior_w else 0.0
        gw_ptr, gw_cs, ws_ptr, ws_n = gW.data_ptr(), gW.shape[1], ws.data_ptr(), ws.numel()
        failure = []

        def finalize_layer(l, stream_ptr):
            _ffi.check(L.dgprf_grad_finalize_layer(C.byref(m), l, B, mode, ws_ptr, ws_n, gw_ptr, gw_cs, prior, stream_ptr))

        def _hook(l, _user):
            try:
                layer_hook(l, finalize_layer)
            except BaseException as exc:        # a ctypes callback cannot raise through the C frames
                failure.append(exc)

        cb = _ffi.LAYER_HOOK(_hook)
        L.dgprf_set_backward_hook(C.cast(cb, C.c_void_p), None)
        try:
            _ffi.check(L.dgprf_gradients(C.byref(m), X.data_ptr(), x_cs, Y.data_ptr(), y_cs, B, mode, ws_ptr, ws_n,
                                         None, gw_cs, None, self.layout.h_len, 0.0, 0, tot.data_ptr(),
                                         0.0 if inv_B is None else float(inv_B), 0, st))
        finally:
            L.dgprf_set_backward_hook(None, None)
        if failure:
            raise failure[0]
        return tot, gW, gH

    def step(self, X, Y, data_size: float, lr: float, momentum_decay: float, temperature: float,
             resample: bool, full_bayesian: bool, seed: int, step: int,
             eps_w=None, res_w=None, eps_h=None, res_h=None, u_out: Optional[torch.Tensor] = None,
             step_base: Optional[torch.Tensor] = None):
        """One sgmcmc_update (models/dgp.py:184-216) for all C chains: a single C call.
        step_base (int64 device scalar): the Philox step becomes step + *step_base (CUDA-graph replays)."""
        m = self.model()
        X, x_cs, Y, y_cs, B = self._xy(X, Y)
        mode = _ffi.MODE_HYPER if full_bayesian else _ffi.MODE_TRAIN
        ws = self.workspace(m, B, mode)
        sw, nsw, sh, nsh = self._segments()
        p = _ffi.ptr
        if step_base is not None:
            assert eps_w is None and res_w is None and eps_h is None and res_h is None
            assert step_base.dtype == torch.int64 and step_base.is_cuda and step_base.numel() == 1
            _ffi.check(_ffi.lib().dgprf_sgmcmc_step_graph(
                C.byref(m), X.data_ptr(), x_cs, Y.data_ptr(), y_cs, B, int(full_bayesian),
                self.theta_w.data_ptr(), self.mom_w.data_ptr(), self.layout.w_len, sw, nsw,
                self.theta_h.data_ptr(), self.mom_h.data_ptr(), self.layout.h_len, sh, nsh,
                float(lr), float(data_size), float(momentum_decay), float(temperature), int(bool(resample)),
                int(seed), int(step), step_base.data_ptr(), ws.data_ptr(), ws.numel(), p(u_out), _ffi.stream_ptr()))
            return
        _ffi.check(_ffi.lib().dgprf_sgmcmc_step(
            C.byref(m), X.data_ptr(), x_cs, Y.data_ptr(), y_cs, B, int(full_bayesian),
            self.theta_w.data_ptr(), self.mom_w.data_ptr(), self.layout.w_len, sw, nsw,
            self.theta_h.data_ptr(), self.mom_h.data_ptr(), self.layout.h_len, sh, nsh,
            float(lr), float(data_size), float(momentum_decay), float(temperature), int(bool(resample)),
            int(seed), int(step), p(eps_w), p(res_w), p(eps_h), p(res_h),
            ws.data_ptr(), ws.numel(), p(u_out), _ffi.stream_ptr()))

    def step_host(self, X_host: torch.Tensor, Y_host: torch.Tensor, data_size: float, lr: float, momentum_decay: float,
                  temperature: float, resample: bool, full_bayesian: bool, seed: int, step: int,
                  u_host: Optional[torch.Tensor] = None):
        """sgmcmc_update from a HOST minibatch (fp32 contiguous CPU tensors, ideally pinned): the H2D copies,
        the step and the optional D2H read of sum_i ll_i are enqueued by ONE C call; nothing synchronises.
        Everything that does not change between steps is converted to ctypes once and cached.

        Lifetime contract (the reference's sgmcmc_update consumes its batch by value, models/dgp.py:184): PINNED
        buffers are read by the GPU in place, after this call has returned.  The engine therefore keeps a reference to
        X_host / Y_host / u_host until a CUDA event recorded after the step has completed, so a caller that drops its
        batch right away (a DataLoader(pin_memory=True) loop) cannot have the block recycled under the kernel.
        What the caller must not do is WRITE into the same pinned tensors before the step has run (synchronise, or
        use a fresh / pageable tensor per step -- pageable batches are staged by value by the driver)."""
        B, dx = X_host.shape
        if self.spec.likelihood == "softmax" and Y_host.shape[1] != 1:
            Y_host = Y_host[:, :1].contiguous()         # labels live in column 0 (likelihoods/softmax.py:14), as _xy does
            if X_host.is_pinned():
                Y_host = Y_host.pin_memory()
        yc = Y_host.shape[1]
        assert yc == (self.spec.d_out if self.spec.likelihood == "gaussian" else 1), \
            f"Y has width {yc}, the likelihood expects {self.spec.d_out if self.spec.likelihood == 'gaussian' else 1}"
        assert Y_host.shape[0] == B and dx == self.spec.d_in, "X / Y shape does not match the model / batch"
        key = (B, yc, full_bayesian)
        st = self._staging.get(key)
        if st is None:
            assert dx == self.spec.d_in and Y_host.shape[0] == B and self.C == 1
            m = self.model()
            mode = _ffi.MODE_HYPER if full_bayesian else _ffi.MODE_TRAIN
            ws = self.workspace(m, B, mode)
            xd = torch.empty(2 * B, dx, device=self.device)        # two halves: pipelined staging (zero_copy mode 2)
            yd = torch.empty(2 * B, yc, device=self.device)
            ud = torch.zeros(max(1, self.C), device=self.device)
            sw, nsw, sh, nsh = self._segments()
            fn = _ffi.lib().dgprf_sgmcmc_step_host
            head = (C.byref(m),)
            mid = (C.c_int(yc), C.c_int(B), C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr())),
            mid2 = (C.c_int(int(full_bayesian)),
                   C.c_void_p(self.theta_w.data_ptr()), C.c_void_p(self.mom_w.data_ptr()), C.c_int64(self.layout.w_len), sw,
                   C.c_int(nsw), C.c_void_p(self.theta_h.data_ptr()), C.c_void_p(self.mom_h.data_ptr()),
                   C.c_int64(self.layout.h_len), sh, C.c_int(nsh))
            tail = (C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()), C.c_void_p(ud.data_ptr()))
            mid = (mid[0], mid2)
            st = (fn, head, mid, tail, (m, ws, xd, yd, ud, sw, sh), self._seg_cache)
            self._staging[key] = st
        elif st[5] is not self._seg_cache:                 # masses changed: rebuild the cached segment tables
            del self._staging[key]
            return self.step_host(X_host, Y_host, data_size, lr, momentum_decay, temperature, resample, full_bayesian,
                                  seed, step, u_host)
        fn, head, mid, tail = st[0], st[1], st[2], st[3]
        # pinned host tensors: copied on a side stream under the previous step's kernels (mode 2; DGPRF_ZERO_COPY_READ=1:
        # read in place by the kernels, mode 1); pageable ones go through in-order staging copies (mode 0)
        zc = _ZC_MODE if (X_host.is_pinned() and Y_host.is_pinned() and (u_host is None or self._pinned_out(u_host))) else 0
        rc = fn(*head, X_host.data_ptr(), Y_host.data_ptr(), *mid[0], zc, *mid[1], lr, data_size, momentum_decay, temperature,
                1 if resample else 0, seed, step, *tail, u_host.data_ptr() if u_host is not None else None,
                _ffi.stream_ptr())
        if rc:
            _ffi.check(rc)
        if zc:
            self._hold(X_host, Y_host, u_host)

    def _pinned_out(self, u_host) -> bool:
        """is_pinned() of the output word, remembered for the tensor last seen (a sampler passes the same one every step)."""
        c = self._u_pinned
        if c is not None and c[0] is u_host:
            return c[1]
        self._u_pinned = (u_host, u_host.is_pinned())
        return self._u_pinned[1]

    def _hold(self, *tensors):
        """Keep zero-copy host buffers alive until the GPU is done with them: every 8th step records an event; entries
        up to the newest completed event are released (at most 64 steps are ever outstanding)."""
        self._held += 1
        ev = None
        if (self._held & 7) == 0:
            ev = torch.cuda.Event()
            ev.record()
        q = self._inflight
        q.append((ev, tensors))
        done = -1
        for i, (e, _) in enumerate(q):
            if e is not None:
                if len(q) - i > 64:
                    e.synchronize()
                if e.query():
                    done = i
                else:
                    break
        for _ in range(done + 1):
            q.popleft()

    def log_prior(self, t: torch.Tensor) -> torch.Tensor:
        """sum log N(t; 0, 1) over a contiguous tensor -> scalar tensor (models/dgp.py:129-136)."""
        _ffi.require_cuda()
        t = t.contiguous()
        o = torch.empty(1, device=self.device, dtype=torch.float32)
        _ffi.check(_ffi.lib().dgprf_log_prior(t.data_ptr(), 0, t.numel(), 1, o.data_ptr(), _ffi.stream_ptr()))
        return o[0]
