"""Minibatch pipelines of the drivers (experiments/utils_dataset.py:7-65 of the reference), device resident.

The reference builds `tf.data` pipelines: shuffle over the whole set (re-drawn every epoch), `batch(batch_size,
drop_remainder=True)` for the training part, `drop_remainder=False` for the test part.  Here the arrays live in HBM and
a `DeviceDataset` yields views of a permuted copy: every `__iter__` draws a new permutation on the device and gathers the
rows into a buffer at a FIXED address, so the batches of every epoch sit at the same pointers -- which is what lets a
whole epoch of sampling steps be captured once as a CUDA graph and replayed (experiments/utils_training.py).

    download_UCI_data_info    utils_dataset.py:17-25
    load_UCI_dataset          utils_dataset.py:27-45
    load_tf_dataset           utils_dataset.py:47-60   (tensorflow_datasets is not available offline: reads an .npz)
    normalize_MNIST           utils_dataset.py:62-65
"""
import os

import numpy as np
import torch

from dgprf import _ffi
from experiments.datasets import Datasets


class DeviceDataset:
    """Re-iterable minibatch source over device-resident arrays: yields (X_b, Y_b) views.

    shuffle=True redraws the permutation at the start of every pass (tf.data `shuffle(N)` with its default
    reshuffle_each_iteration); drop_remainder as in `tf.data.Dataset.batch`."""

    def __init__(self, X, Y, batch_size, shuffle=True, drop_remainder=True, seed=None, device=None):
        dev = _ffi.require_cuda() if device is None else device
        self.X, self.Y = _ffi.as_dev(X, dev), _ffi.as_dev(Y, dev)
        if self.Y.ndim == 1:
            self.Y = self.Y[:, None].contiguous()
        assert self.X.shape[0] == self.Y.shape[0], "X and Y hold different numbers of rows"
        self.N = int(self.X.shape[0])
        self.batch_size = int(batch_size)
        self.shuffle, self.drop_remainder = bool(shuffle), bool(drop_remainder)
        self._gen = torch.Generator(device=dev)
        if seed is not None:
            self._gen.manual_seed(int(seed))
        self._Xp = torch.empty_like(self.X) if self.shuffle else self.X      # fixed-address permuted copies
        self._Yp = torch.empty_like(self.Y) if self.shuffle else self.Y
        self.passes = 0

    def __len__(self):
        n, r = divmod(self.N, self.batch_size)
        return n if (self.drop_remainder or r == 0) else n + 1

    def reshuffle(self):
        """A new permutation of the rows, gathered into the fixed buffers (two device kernels, no host sync)."""
        if self.shuffle:
            perm = torch.randperm(self.N, device=self.X.device, generator=self._gen)
            torch.index_select(self.X, 0, perm, out=self._Xp)
            torch.index_select(self.Y, 0, perm, out=self._Yp)
        self.passes += 1

    def batch(self, i):
        lo = i * self.batch_size
        hi = min(self.N, lo + self.batch_size)
        return self._Xp[lo:hi], self._Yp[lo:hi]

    def __iter__(self):
        self.reshuffle()
        for i in range(len(self)):
            yield self.batch(i)

    def repeat(self):
        """`ds.repeat()` of tf.data: an endless stream of batches, pass after pass."""
        while True:
            yield from self

    @property
    def shape(self):
        return tuple(self.X.shape)


def download_UCI_data_info(name, data_path='./data/'):
    dataset = Datasets(data_path=data_path).all_datasets[name]
    data = dataset.get_data()
    X, Y, Xs, Ys, X_mean, Y_mean, Y_std = [np.float32(data[_]) for _ in ['X', 'Y', 'Xs', 'Ys', 'X_mean', 'Y_mean', 'Y_std']]
    assert dataset.N == X.shape[0] + Xs.shape[0], f"N + Ns does not match dataset.N (should be {X.shape[0] + Xs.shape[0]})! "
    assert dataset.D == X.shape[1], f"D does not match dataset.D(should be {X.shape[1]})!"
    return X, Y, Xs, Ys, X_mean, Y_mean, Y_std                     # Y_mean, Y_std shape [1,]


def load_UCI_dataset(dataset_name, batch_size=128, transform_fn=None, data_dir='./data/', drop_train_remainder=True,
                     seed=None, verbose=True):
    X, Y, Xs, Ys, X_mean, Y_mean, Y_std = download_UCI_data_info(dataset_name, data_path=data_dir)
    if verbose:
        print('#' * 30 + f" Getting data info:dataset name: {dataset_name} " + '#' * 30)
        print(f"D: {X.shape[1]}, N: {X.shape[0]}, Ns: {Xs.shape[0]}")
        print(f"X_mean: {X_mean}, Y_mean: {Y_mean}, Y_std: {Y_std}")
        print('#' * 70)
    train_shape, test_shape = np.shape(X), np.shape(Xs)
    if transform_fn is not None:                                   # element-wise transform of (x, y) rows
        X, Y = transform_fn(X, Y)
        Xs, Ys = transform_fn(Xs, Ys)
    ds_train = DeviceDataset(X, Y, batch_size, shuffle=True, drop_remainder=drop_train_remainder, seed=seed)
    # if the batch is larger than the test set, one batch with the whole data comes back
    ds_test = DeviceDataset(Xs, Ys, batch_size, shuffle=True, drop_remainder=False, seed=None if seed is None else seed + 1)
    return ds_train, ds_test, train_shape, test_shape


def normalize_MNIST(img, label):
    """img uint8 [..., 28, 28] -> float32 [..., 784] in [-0.5, 0.5]; label -> float32 [..., 1]."""
    img = np.asarray(img)
    img = img.reshape(img.shape[:-2] + (28 * 28,)).astype(np.float32) / np.float32(255.) - np.float32(0.5)
    label = np.asarray(label).astype(np.float32).reshape(-1, 1) if np.ndim(label) > 0 else np.float32(label).reshape(1)
    return img, label


def load_tf_dataset(dataset_name, batch_size=128, transform_fn=None, data_dir='./tensorflow_datasets/', seed=None):
    """tensorflow_datasets cannot be used offline: reads `<data_dir>/<dataset_name>.npz` with arrays x_train, y_train,
    x_test, y_test (the layout of keras' mnist.npz) instead.  Returns (ds_train, ds_test, train size, test size)."""
    path = os.path.join(data_dir, f"{dataset_name}.npz")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path} not found: tensorflow_datasets is unavailable offline, provide the arrays as an .npz")
    with np.load(path) as f:
        x_train, y_train, x_test, y_test = f['x_train'], f['y_train'], f['x_test'], f['y_test']
    if transform_fn is not None:
        x_train, y_train = transform_fn(x_train, y_train)
        x_test, y_test = transform_fn(x_test, y_test)
    ds_train = DeviceDataset(x_train, y_train, batch_size, shuffle=True, drop_remainder=True, seed=seed)
    ds_test = DeviceDataset(x_test, y_test, batch_size, shuffle=True, drop_remainder=False,
                            seed=None if seed is None else seed + 1)
    return ds_train, ds_test, int(x_train.shape[0]), int(x_test.shape[0])
