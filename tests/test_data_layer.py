"""Data layer: UCI CSV reader / split / normaliser and the device-resident minibatch source
(experiments/datasets.py:26-87, experiments/utils_dataset.py:7-65 of the reference).  CPU tests: host logic only."""
import json
import os

import numpy as np
import pytest
import torch

from experiments.datasets import Dataset, Datasets
from experiments.utils_dataset import DeviceDataset, normalize_MNIST

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSV_DIR = "/root/reference/Baselines/RF_DGP/data/"


def _write_csv(tmp_path, name, N, D, seed=0):
    rng = np.random.RandomState(seed)
    data = np.concatenate([rng.randn(N, D) * rng.rand(1, D) * 5 + rng.randn(1, D), 3.0 * rng.randn(N, 1) + 7.0], 1)
    np.savetxt(os.path.join(tmp_path, f"{name}.csv"), data, delimiter=",")
    return data


def test_split_and_normalisation(tmp_path):
    tmp = str(tmp_path) + "/"
    raw = _write_csv(tmp, "toy", 103, 4)
    ds = Dataset("toy", 103, 4, "regression", tmp)
    a, b = ds.get_data(), ds.get_data()
    assert a["X"].shape == (92, 4) and a["Xs"].shape == (11, 4) and a["Y"].shape == (92, 1)      # int(N * 0.9)
    assert np.array_equal(a["X"], b["X"])                                                       # seeded split
    assert not np.array_equal(a["X"], ds.get_data(split=1)["X"])
    # the split is np.random.seed(seed + split); shuffle(arange(N))  (datasets.py:58-72)
    ind = np.arange(103); np.random.seed(0); np.random.shuffle(ind)
    Xtr = raw[ind[:92], :-1]
    assert np.allclose(a["X_mean"], Xtr.mean(0)) and np.allclose(a["X_std"], Xtr.std(0) + 1e-6)
    assert np.allclose(a["X"], (Xtr - Xtr.mean(0)) / (Xtr.std(0) + 1e-6))
    assert np.allclose(a["Xs"], (raw[ind[92:], :-1] - Xtr.mean(0)) / (Xtr.std(0) + 1e-6))          # TRAIN statistics on the test part
    assert abs(a["X"].mean()) < 1e-9 and np.allclose(a["X"].std(0), 1.0, atol=1e-5)
    Ytr = raw[ind[:92], -1:]
    assert np.allclose(a["Y_mean"], Ytr.mean(0)) and np.allclose(a["Y_std"], Ytr.std(0) + 1e-6)
    assert np.allclose(a["Y"], (Ytr - Ytr.mean(0)) / (Ytr.std(0) + 1e-6))


def test_missing_csv_fails_loudly(tmp_path):
    with pytest.raises(FileNotFoundError):
        Datasets(str(tmp_path) + "/").all_datasets["boston"].get_data()
    assert set(Datasets(str(tmp_path) + "/").all_datasets) == {"boston", "concrete", "energy", "kin8nm", "naval", "power",
                                                                "protein", "wine_red", "wine_white"}


@pytest.mark.skipif(not os.path.isfile(REF_CSV_DIR + "boston.csv"), reason="the reference's CSV is only present in the build container")
def test_boston_matches_the_reference_notebook_output():
    """train_regression_UCI.ipynb cell 1 printed: D 13, N 455, Ns 51, X_mean[...], Y_mean [22.656263], Y_std [9.32293]."""
    kat = json.load(open(os.path.join(HERE, "golden", "uci_boston_kat.json")))
    nb = kat["notebook"]
    assert "Y_std: [9.32293]" in kat["notebook_stdout"] and "N: 455, Ns: 51" in kat["notebook_stdout"]
    d = Datasets(REF_CSV_DIR).all_datasets["boston"].get_data()
    assert d["X"].shape == (nb["N"], nb["D"]) and d["Xs"].shape == (nb["Ns"], nb["D"])
    assert np.float32(d["Y_mean"])[0] == pytest.approx(nb["Y_mean"], rel=2e-7)
    assert np.float32(d["Y_std"])[0] == pytest.approx(nb["Y_std"], rel=2e-7)
    assert np.allclose(np.float32(d["X_mean"]), np.float32(nb["X_mean"]), rtol=2e-7)
    assert np.allclose(np.float32(d["X"][0]), kat["derived"]["X_row0"], rtol=1e-6)
    assert np.allclose(np.float32(d["Y"][:5, 0]), kat["derived"]["Y_head"], rtol=1e-6)


def test_device_dataset_batching_semantics():
    cpu = torch.device("cpu")
    X = torch.arange(23, dtype=torch.float32)[:, None].repeat(1, 3)
    Y = torch.arange(23, dtype=torch.float32)
    tr = DeviceDataset(X, Y, 5, shuffle=True, drop_remainder=True, seed=3, device=cpu)
    assert len(tr) == 4 and tr.shape == (23, 3)
    p1 = [(x.clone(), y.clone()) for x, y in tr]
    p2 = [(x.clone(), y.clone()) for x, y in tr]
    assert all(x.shape == (5, 3) and y.shape == (5, 1) for x, y in p1)
    assert all(torch.equal(x[:, 0], y[:, 0]) for x, y in p1)                                     # rows stay paired
    seen1 = torch.cat([y for _, y in p1]).reshape(-1)
    assert len(set(seen1.tolist())) == 20                                                       # no repeats inside a pass
    assert not torch.equal(seen1, torch.cat([y for _, y in p2]).reshape(-1))                     # reshuffled every pass
    ptrs = [x.data_ptr() for x, _ in tr]
    assert ptrs == [x.data_ptr() for x, _ in tr]                                                # fixed addresses (graph capture)
    again = DeviceDataset(X, Y, 5, shuffle=True, drop_remainder=True, seed=3, device=cpu)
    assert torch.equal(torch.cat([y for _, y in again]).reshape(-1), seen1)                      # seeded
    te = DeviceDataset(X, Y, 5, shuffle=False, drop_remainder=False, device=cpu)
    sizes = [x.shape[0] for x, _ in te]
    assert sizes == [5, 5, 5, 5, 3] and torch.equal(torch.cat([y for _, y in te]).reshape(-1), Y)
    big = DeviceDataset(X, Y, 100, shuffle=False, drop_remainder=False, device=cpu)
    assert [x.shape[0] for x, _ in big] == [23]                                                 # batch > N: the whole set once
    assert len(DeviceDataset(X, Y, 100, drop_remainder=True, device=cpu)) == 0                   # ... or nothing, as tf.data
    it = tr.repeat()
    assert sum(1 for _ in zip(range(11), it)) == 11                                             # endless


def test_normalize_mnist():
    img = np.arange(2 * 28 * 28, dtype=np.uint8).reshape(2, 28, 28)
    x, y = normalize_MNIST(img, np.array([3, 9]))
    assert x.shape == (2, 784) and x.dtype == np.float32 and y.shape == (2, 1) and y.dtype == np.float32
    assert x.min() >= -0.5 and x.max() <= 0.5 and x[0, 1] == pytest.approx(1 / 255. - 0.5)
    assert y[:, 0].tolist() == [3.0, 9.0]
