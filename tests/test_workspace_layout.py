"""dgprf_workspace_bytes is pure host logic (no GPU needed): the layout must be valid for every BASELINE config in every
mode and precision, grow with the mode, and never shrink when the tensor-core operand buffers are added."""
import ctypes as C

import pytest

from dgprf import _ffi


def _model(d_in, d_out, n_rf, n_gp, kinds, cat, lik, chains, prec):
    m = _ffi.Model()
    L = len(n_gp)
    m.n_layers, m.likelihood, m.d_in, m.d_out, m.n_chains, m.precision = L, lik, d_in, d_out, chains, prec
    fake = 0x10000                       # non-NULL, 256-byte aligned; workspace_bytes never dereferences it
    m.w_base, m.h_base = fake, fake
    off_w = off_h = 0
    for l in range(L):
        y = m.layer[l]
        y.kind = _ffi.KIND_RBF if kinds[l] == "RBF" else _ffi.KIND_ARC
        y.d_prev = 0 if l == 0 else n_gp[l - 1]
        y.d_x = d_in if (l == 0 or cat) else 0
        y.M, y.g, y.has_mean = n_rf[l], n_gp[l], 0
        d = y.d_prev + y.d_x
        F = 2 * y.M if kinds[l] == "RBF" else y.M
        y.off_W = off_w
        off_w += (F * y.g + 3) // 4 * 4
        y.off_log_amp = off_h
        y.off_log_inv_ls = off_h + 4
        y.off_mean = 0
        off_h += 4 + (d + 3) // 4 * 4
        y.z, y.z_cs = fake, 0
    m.off_lik_log_var = off_h
    m.w_cs, m.h_cs = off_w, off_h + 4
    return m


CONFIGS = {
    "cfg1": (1, 1, [100] * 2, [1, 1], ["RBF"] * 2, False, _ffi.LIK_GAUSSIAN, 1, 20),
    "cfg2": (9, 1, [512] * 3, [9, 9, 1], ["RBF"] * 3, True, _ffi.LIK_GAUSSIAN, 1, 1000),
    "cfg3": (784, 10, [512] * 3, [30, 30, 10], ["ARC"] * 3, True, _ffi.LIK_SOFTMAX, 1, 2048),
    "cfg4": (90, 1, [512] * 3, [30, 30, 1], ["RBF"] * 3, True, _ffi.LIK_GAUSSIAN, 8, 1000),
    "cfg5": (90, 1, [4096] * 5, [30, 30, 30, 30, 1], ["RBF"] * 5, True, _ffi.LIK_GAUSSIAN, 1, 65536),
}


def _bytes(m, B, mode):
    n = C.c_size_t(0)
    rc = _ffi.lib().dgprf_workspace_bytes(C.byref(m), B, mode, C.byref(n))
    assert rc == 0, _ffi.lib().dgprf_last_error().decode()
    return n.value


@pytest.mark.parametrize("name", list(CONFIGS))
def test_workspace_sizes_are_consistent(name):
    d_in, d_out, n_rf, n_gp, kinds, cat, lik, chains, B = CONFIGS[name]
    size = {}
    for prec in (_ffi.PREC_FP32, _ffi.PREC_TF32):
        m = _model(d_in, d_out, n_rf, n_gp, kinds, cat, lik, chains, prec)
        e, t, h = (_bytes(m, B, mode) for mode in (_ffi.MODE_EVAL, _ffi.MODE_TRAIN, _ffi.MODE_HYPER))
        assert 0 < e <= t <= h
        assert e % 256 == 0 and t % 256 == 0 and h % 256 == 0
        # the saved features dominate a TRAIN workspace: at least B * F * 4 bytes per layer and chain
        feat = sum(4 * B * (2 * M if k == "RBF" else M) for M, k in zip(n_rf, kinds)) * chains
        assert t >= feat
        size[prec] = (e, t, h)
    # tensor-core mode only ever adds operand buffers on top of the fp32 layout's saved features
    assert size[_ffi.PREC_TF32][1] >= 0.5 * size[_ffi.PREC_FP32][1]
    assert size[_ffi.PREC_TF32][1] <= 2.0 * size[_ffi.PREC_FP32][1] + (64 << 20)


def test_workspace_rejects_bad_arguments():
    m = _model(*CONFIGS["cfg2"][:8], _ffi.PREC_FP32)
    n = C.c_size_t(0)
    L = _ffi.lib()
    assert L.dgprf_workspace_bytes(C.byref(m), 0, _ffi.MODE_TRAIN, C.byref(n)) != 0          # B = 0
    assert L.dgprf_workspace_bytes(C.byref(m), 1000, 7, C.byref(n)) != 0                     # unknown mode
    m.layer[1].d_prev = 5                                                                    # does not chain with n_gp[0] = 9
    assert L.dgprf_workspace_bytes(C.byref(m), 1000, _ffi.MODE_TRAIN, C.byref(n)) != 0
    assert b"chain" in L.dgprf_last_error()
