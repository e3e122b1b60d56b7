"""Per-role clock timeline of one CTA of the pipelined forward at configs[4] layer scale (DGPRF_TC2_TIMELINE=<call>)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
import torch
from dgprf import _ffi
from dgprf.engine import Engine, ModelSpec
B, d, M, g = int(os.environ.get("TL_ROWS", 65536)), int(os.environ.get("TL_D", 120)), 4096, 30
spec = ModelSpec.build(d, g, [M], [g], ["RBF"], False, False, "gaussian")
X = torch.randn(B, d, device="cuda")
e = Engine(spec, 1, precision=_ffi.PREC_TF32)
e.theta_w.normal_()
mode = _ffi.MODE_TRAIN if (len(sys.argv) < 2 or sys.argv[1] == "train") else _ffi.MODE_EVAL
for _ in range(6):
    e.forward(X, mode=mode, want_F=False)
torch.cuda.synchronize()
