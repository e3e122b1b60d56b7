"""HBM bandwidth of the fused update kernel on a 256 MiB-per-buffer flat parameter vector."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
import torch
from dgprf import _ffi
n, C = 16 << 20, 4
th = torch.randn(C, n, device="cuda"); mo = torch.randn(C, n, device="cuda"); gr = torch.randn(C, n, device="cuda")
segs = _ffi.make_segments([(0, n, 1.0, 1)])
L = _ffi.lib(); st = torch.cuda.current_stream().cuda_stream
def run(step, T=1.0):
    _ffi.check(L.dgprf_sgmcmc_update(th.data_ptr(), mo.data_ptr(), n, n, C, gr.data_ptr(), n, 1, 0, segs, 1, 1e-4, 45730.0, 0.9, T, 0, 7, step, None, None, st))
for T in (1.0, 0.0):
    for i in range(3): run(i, T)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for i in range(10): run(10 + i, T)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"T={T}: {ms:.4f} ms  {20.0 * C * n / ms / 1e6:.1f} GB/s")
