"""Two gradient passes of a 3-layer configs[4]-shaped model (B = 65536, M = 4096, n_gp = [30, 30, 1], input concatenation, tf32)
for ncu captures: launches per pass are k1_fwd_tc2 x 3 (layers 0, 1, 2), k2_bwd_tc2 x 3 (layers 2, 1, 0); layer 1 (input width
120 = 30 + 90, n_gp = 30, T = dP z^T formed for the layer below) is the representative one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
import torch
from dgprf import _ffi
from dgprf.engine import Engine, ModelSpec
B, M = 65536, 4096
spec = ModelSpec.build(90, 1, [M] * 3, [30, 30, 1], ["RBF"] * 3, True, False, "gaussian")
torch.manual_seed(0)
e = Engine(spec, 1, precision=_ffi.PREC_TF32)
e.theta_w.normal_()
e.theta_h[:, e.layout.off_lik_log_var] = -2.0
X = torch.randn(B, 90, device="cuda"); Y = torch.randn(B, 1, device="cuda")
for _ in range(2):
    e.gradients(X, Y, 1e5, hyper=False, prior_w=True, prior_h=False)
torch.cuda.synchronize()
print("ok")
