"""K7 (posterior-predictive averaging, experiments/utils_training.py:79-85) at a bandwidth-relevant size:
S stored samples x N test points of log-probabilities (+ the squared-error matrix), reduced to
(mean_n logsumexp_s - log S, sqrt(mean se)).  Algorithmic bytes: 4 S N (x2 with the aux matrix)."""
import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
from dgprf import _ffi
L = _ffi.lib()
dev = torch.device("cuda")
for S, N, with_aux in ((64, 1 << 21, False), (64, 1 << 21, True), (200, 515345 // 10, True)):
    lp = torch.randn(S, N, device=dev) - 1.0
    ax = torch.rand(S, N, device=dev) if with_aux else None
    res = torch.empty(2, device=dev)
    scratch = torch.empty(2 * ((N + 255) // 256) + 2, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def run():
        _ffi.check(L.dgprf_predictive_reduce(lp.data_ptr(), ax.data_ptr() if with_aux else None, S, N, N, math.log(S), 1, None,
                                             res.data_ptr(), scratch.data_ptr(), st))
    for _ in range(3): run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    byts = 4.0 * S * N * (2 if with_aux else 1)
    ref = float((torch.logsumexp(lp.double(), 0) - math.log(S)).mean())
    print(f"S={S} N={N} aux={with_aux}: {ms:.3f} ms  {byts / ms / 1e6:.0f} GB/s   value {float(res[0]):.6f} (torch fp64 {ref:.6f})")
