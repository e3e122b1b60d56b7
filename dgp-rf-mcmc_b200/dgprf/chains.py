"""Independent SG-MCMC chains batched on one GPU and sharded over GPUs (BASELINE.json configs[3]).

In the reference a chain is one `DGP_RF` instance (own z, W, moments, hyper-parameters) run through one
sampler loop; nothing is shared between instances.  `ChainEnsemble` holds C such chains of one architecture in
the engine's flat buffers ([C][w_len] ...), so every kernel launch advances all C chains (grid.z / grid.y =
chain).  Chains are identified by a GLOBAL id `chain_base + c`: initial state and Philox noise are functions
of (seed, global id) only, so sharding 64 chains as 8 per GPU gives bit-for-bit what one process would
compute for all 64 -- no communication during sampling (dgprf/dist.py combines the predictive statistics).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

from . import _ffi
from .engine import Engine, ModelSpec
from .variable import out

_K1 = 0x9E3779B97F4A7C15          # chain stride of the Philox key (csrc/philox.cuh: philox_key)
_MASK = (1 << 64) - 1


class ChainEnsemble:
    def __init__(self, d_in: int, d_out: int, n_hidden_layers: int, n_rf, n_gp, kernel_type_list: Optional[Sequence[str]] = None,
                 input_cat: bool = False, likelihood: str = "gaussian", lik_variance: float = 0.1, n_chains: int = 8,
                 chain_base: int = 0, seed: int = 0, precision: Optional[str] = None, device=None):
        L = n_hidden_layers
        n_rf = [int(n_rf)] * L if isinstance(n_rf, int) else [int(v) for v in n_rf]
        n_gp = [int(n_gp)] * L if isinstance(n_gp, int) else [int(v) for v in n_gp]
        kinds = ["RBF"] * L if kernel_type_list is None else list(kernel_type_list)
        assert n_gp[-1] == d_out and len(n_rf) == L and len(n_gp) == L and len(kinds) == L
        self.spec = ModelSpec.build(d_in, d_out, n_rf, n_gp, kinds, input_cat, False, likelihood)
        self.C, self.chain_base, self.seed = int(n_chains), int(chain_base), int(seed)
        prec = None if precision is None else {"fp32": _ffi.PREC_FP32, "tf32": _ffi.PREC_TF32}[precision]
        self.engine = e = Engine(self.spec, self.C, device, precision=prec, shared_z=False)
        self._step = 0
        # defaults of the reference: z, W, moments ~ N(0,1); log_amp = 0; log_inv_ls = -0.5 log d_l; lik var 0.1
        for c in range(self.C):
            g = torch.Generator().manual_seed((self.seed * 1000003 + self.chain_base + c) & 0x7FFFFFFF)
            for l, s in enumerate(self.spec.layers):
                e.z[l][c].copy_(torch.randn(s.d, s.M, generator=g))
                e.view(f"W_{l}", chain=c).copy_(torch.randn(s.F, s.g, generator=g))
                e.view(f"W_{l}", "mom", chain=c).copy_(torch.randn(s.F, s.g, generator=g))
                e.view(f"log_inv_ls_{l}", chain=c).fill_(-0.5 * math.log(s.d))
            if likelihood == "gaussian":
                e.view("lik_log_var", chain=c).fill_(math.log(lik_variance))

    @property
    def philox_seed(self) -> int:
        """Kernels key the noise by (seed, local chain); shifting the seed by chain_base * K1 makes it a
        function of the GLOBAL chain id."""
        return (self.seed + self.chain_base * _K1) & _MASK

    def sgmcmc_update(self, X, Y, data_size, lr=0.01, momentum_decay=0.95, resample_moments=False, temperature=1.0):
        """One SGHMC / SGLD step of every chain.  X: [B, d_in] (shared minibatch) or [C, B, d_in]."""
        self._step += 1
        self.engine.step(X, Y, float(data_size), float(lr), float(momentum_decay), float(temperature),
                         bool(resample_moments), False, self.philox_seed, self._step)

    def evaluate(self, X, Y):
        """Per-point log-likelihood and squared error / correct flag of every chain: ([C, N], [C, N])."""
        ll, aux, _ = self.engine.evaluate(X, Y)
        return out(ll), out(aux)

    def forward(self, X):
        return out(self.engine.forward(X))

    def W(self, chain: int) -> List[torch.Tensor]:
        return [self.engine.view(f"W_{l}", chain=chain) for l in range(len(self.spec.layers))]
