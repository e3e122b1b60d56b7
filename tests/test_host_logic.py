"""Host-side logic of the drop-in surface that needs no GPU: shapes, flat layout, schedule,
error behaviour, and that compute fails loudly without CUDA (no CPU fallback)."""
import ctypes

import numpy as np
import pytest
import torch

import dgprf_oracle as O
from dgprf import _ffi
from dgprf.engine import FlatLayout, ModelSpec
from kernels import ARCKernel, RBFKernel
from layers import ARCLayer, GPLayer, RBFLayer
from likelihoods import Gaussian, Softmax
from models.classification_model import ClassificationDGP
from models.regression_model import RegressionDGP
from utils import BNN_from_list, BNN_from_list_input_cat, cyclical_step_rate


def test_baseline_config_parameter_counts():
    """nW of SURVEY section 8's config table: cfg2 19456, cfg3 35840, cfg4 62464, cfg5 991232."""
    def nW(d_in, n_rf, n_gp, kinds, L):
        spec = ModelSpec.build(d_in, n_gp[-1], [n_rf] * L, n_gp, kinds, True, False, "gaussian")
        return sum(s.F * s.g for s in spec.layers), [s.d for s in spec.layers]
    assert nW(9, 512, [9, 9, 1], ["RBF"] * 3, 3) == (19456, [9, 18, 18])
    assert nW(784, 512, [30, 30, 10], ["ARC"] * 3, 3) == (35840, [784, 814, 814])
    assert nW(90, 512, [30, 30, 1], ["RBF"] * 3, 3) == (62464, [90, 120, 120])
    assert nW(90, 4096, [30, 30, 30, 30, 1], ["RBF"] * 5, 5) == (991232, [90, 120, 120, 120, 120])


def test_flat_layout_is_16_byte_aligned_and_disjoint():
    spec = ModelSpec.build(3, 2, [30, 70, 65], [5, 2, 2], ["RBF", "ARC", "RBF"], True, True, "gaussian")
    lay = FlatLayout.of(spec)
    offs = lay.off_W + lay.off_log_amp + lay.off_log_inv_ls + lay.off_mean + [lay.off_lik_log_var]
    assert all(o % 4 == 0 for o in offs)
    assert lay.w_len % 4 == 0 and lay.h_len % 4 == 0
    ends = [o + s.F * s.g for o, s in zip(lay.off_W, spec.layers)]
    assert all(e <= o for e, o in zip(ends[:-1], lay.off_W[1:])) and ends[-1] <= lay.w_len


def test_class_surface_shapes_and_views():
    m = RegressionDGP(9, 1, n_hidden_layers=3, n_rf=64, n_gp=[9, 9, 1], input_cat=True)
    assert [w.shape for w in m.W_mcmc] == [(128, 9), (128, 9), (128, 1)]
    assert [k.n_feature for k in m.kernel_list] == [9, 18, 18]
    assert isinstance(m.BNN, BNN_from_list_input_cat) and len(m.BNN.layers) == 6
    assert isinstance(m.likelihood, Gaussian)
    assert len(m.Omega_hyperparams) == 6 and len(m.Likelihood_hyperparams) == 1
    assert len(m.trainable_variables) == 10
    # the Variables are views into the flat buffers
    w1 = m.W_mcmc[1]
    w1.assign(torch.ones(128, 9))
    off = m._engine.layout.off_W[1]
    assert float(m._engine.theta_w[0, off:off + 128 * 9].sum()) == 128 * 9
    m.assign_W([np.zeros((128, 9), np.float32), np.zeros((128, 9), np.float32), np.zeros((128, 1), np.float32)])
    assert float(m._engine.theta_w.abs().sum()) == 0.0
    # default initialisation (kernels/RBF.py:16-17,39-40; likelihoods/gaussian.py:7)
    assert float(m.kernel_list[1].log_amplitude) == 0.0
    assert np.allclose(m.kernel_list[1].length_scale.numpy(), np.sqrt(18.0))
    assert float(m.likelihood.variance) == pytest.approx(0.1)


def test_sampler_state_attributes_follow_the_reference():
    m = RegressionDGP(2, 1, n_hidden_layers=2, n_rf=8, n_gp=[2, 1])
    w = m.W_mcmc[0]
    assert not hasattr(w, "moments") and not hasattr(w, "M")
    with pytest.raises(AssertionError):
        m.sgmcmc_update(torch.zeros(4, 2), torch.zeros(4, 1), 10)      # dgp.py:208
    m.precond_update(None, 10, precond_type="identity")
    assert w.M == 1.0 and w.moments.shape == (16, 2)
    assert not hasattr(m.kernel_list[0].log_amplitude, "moments")       # W-only unless full_bayesian
    m.precond_update(None, 10, precond_type="identity", full_bayesian=True)
    assert hasattr(m.kernel_list[0].log_amplitude, "moments")
    with pytest.raises(NotImplementedError):
        m.precond_update(None, 10, precond_type="adam")


def test_kernel_constructor_semantics():
    k = RBFKernel(n_feature=3, is_ard=True)
    assert k.is_ard and k.log_inv_length_scale.shape == (3,)
    k = RBFKernel(n_feature=3, length_scale=[1., 2., 4.])
    assert k.is_ard and np.allclose(k.inv_length_scale.numpy(), [1., .5, .25])
    k = RBFKernel(n_feature=3)
    assert not k.is_ard and k.log_inv_length_scale.shape == ()
    with pytest.raises(ValueError):
        RBFKernel(n_feature=3, length_scale=[1., 2.])
    with pytest.raises(ValueError):
        RBFKernel(n_feature=2, length_scale=[[1., 2.]])
    with pytest.raises(NotImplementedError):
        ARCKernel(n_feature=2, degree=2)
    with pytest.raises(AssertionError):
        RBFLayer(ARCKernel(n_feature=2), 4)
    assert RBFLayer(RBFKernel(2), 5).n_rf == 10 and ARCLayer(ARCKernel(2), 5).n_rf == 5
    with pytest.raises(NotImplementedError):
        RegressionDGP(2, 1, kernel_type_list=["MATERN"], n_gp=1)
    with pytest.raises(AssertionError):
        RegressionDGP(2, 1, n_hidden_layers=2, n_rf=[4], n_gp=[1, 1])


def test_cyclical_step_rate_matches_oracle():
    for cyc in (3, 50, 150):
        for s in (1, 2, cyc - 1, cyc, cyc + 1, 3 * cyc):
            for sched in ("cosine", "glide", "flat"):
                a, ea = cyclical_step_rate(s, cyc, sched, 0.001)
                b, eb = O.cyclical_step_rate(s, cyc, sched, 0.001)
                assert float(a) == float(b) and ea == eb
    with pytest.raises(ValueError):
        cyclical_step_rate(0, 10)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_cuda():
    m = ClassificationDGP(4, 3, n_hidden_layers=1, n_rf=8, n_gp=[3])
    m.precond_update(None, 10, precond_type="identity")
    with pytest.raises(_ffi.DgprfError):
        m.U(torch.zeros(5, 4), torch.zeros(5, 1), 10)
    with pytest.raises(_ffi.DgprfError):
        m.sgmcmc_update(torch.zeros(5, 4), torch.zeros(5, 1), 10)
    with pytest.raises(_ffi.DgprfError):
        GPLayer(4, 2)(torch.zeros(3, 4))
    with pytest.raises(_ffi.DgprfError):
        Softmax().log_prob(torch.zeros(3, 4), torch.zeros(3, 1))


def test_philox_host_reference_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (the generator of the update kernel)."""
    from philox_ref import philox4x32_10
    assert philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_driver_signatures_match_the_reference():
    """Positional parameter names and defaults of the drivers, as written in the reference
    (experiments/utils_training.py:11-16, 93-98, 174-177, 258-261, 339, 361-363, 381-383, 431-433;
    experiments/utils_training_demo.py:10-14, 87-89, 171, 193-195, 215-217)."""
    import inspect
    from experiments import utils_training as T, utils_training_demo as D

    def pos(fn):
        return [(n, p.default) for n, p in inspect.signature(fn).parameters.items()
                if p.kind is inspect.Parameter.POSITIONAL_OR_KEYWORD]

    train = [("lr_0", 0.01), ("momentum_decay", 0.9), ("full_bayesian", True), ("precond_type", "identity"), ("K_batches", None),
             ("second_moment_centered", None), ("resample_in_cycle_head", False), ("total_epochs", 5000),
             ("start_sampling_epoch", 2000), ("epochs_per_cycle", 50), ("print_epoch_cycle", 100)]
    E = inspect.Parameter.empty
    assert pos(T.regression_train) == [("model", E), ("dataset_name", "boston"), ("batch_size", 200), ("data_dir", "./data/")] + train
    assert pos(T.classification_train) == [("model", E), ("dataset_name", "mnist"), ("batch_size", 200),
                                           ("data_dir", "./tensorflow_datasets/")] + train
    smp = [("lr_0", 0.01), ("momentum_decay", 0.9), ("precond_type", "identity"), ("K_batches", None),
           ("second_moment_centered", None), ("resample_in_cycle_head", True), ("start_sampling_epoch", 2000), ("epochs_per_cycle", 50)]
    assert pos(T.MCEM_sampler_UCI) == [("model", E), ("dataset_name", "boston"), ("batch_size", 200), ("data_dir", "./data/")] + smp
    assert pos(T.MCEM_sampler_classification)[4:] == smp
    assert [n for n, _ in pos(T.MCEM_Q_maximizer)] == ["model", "data_size", "optimizer"]
    em = [("sampler_EM", E), ("maximizer", E), ("sampler_fixing_hyper", E), ("total_EM_steps", E), ("ds_train", E)]
    assert pos(T.MCEM) == em + [("num_samples_EM", 100), ("num_samples_fixing_hyper", 200), ("print_epoch_cycle_EM", 100),
                                ("print_epoch_cycle_fixing", 100)]
    win = em + [("num_samples_fixing_hyper", 200), ("window_size", 300), ("print_epoch_cycle_EM", 100), ("print_epoch_cycle_fixing", 100)]
    assert pos(T.MCEM_windows) == win and pos(T.MCEM_increasing_windows) == win
    demo = [("model_demo", E), ("ds_train", E), ("ds_test", E), ("train_size", E), ("batch_size", E), ("X_test", E)]
    assert pos(D.regression_train_demo) == demo + [("lr_0", 0.01), ("momentum_decay", 0.9), ("resample_in_cycle_head", True),
                                                   ("total_epochs", 5000), ("start_sampling_epoch", 2000), ("epochs_per_cycle", 50),
                                                   ("print_epoch_cycle", 100)]
    assert pos(D.MCEM_sampler_demo) == demo + [("lr_0", 0.01), ("momentum_decay", 0.9), ("resample_in_cycle_head", False),
                                               ("start_sampling_epoch", 0), ("epochs_per_cycle", 50)]
    assert [n for n, _ in pos(D.MCEM_Q_maximizer_demo)] == ["model_demo", "data_size", "optimizer"]
    assert pos(D.MCEM_demo) == pos(T.MCEM)
    assert pos(D.MCEM_windows_demo) == em + [("num_samples_fixing_hyper", 200), ("window_size", 50), ("print_epoch_cycle_EM", 100),
                                             ("print_epoch_cycle_fixing", 100)]
