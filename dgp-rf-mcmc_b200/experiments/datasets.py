"""UCI regression data sets: CSV reader, seeded 90/10 split and normalisation (experiments/datasets.py:26-87, 237-257 of
the reference, itself derived from H. Salimbeni's DGP code).

    Dataset.get_data(seed=0, split=0, prop=0.9)   datasets.py:42-56   read -> split -> normalise
    split                                         datasets.py:58-72   np.random.seed(seed + split); shuffle(arange(N))
    normalize                                     datasets.py:74-87   train mean / std (+1e-6) applied to train and test

No downloads (this build runs offline): the CSV must exist under `data_path`; the nine files ship with the reference
under Baselines/RF_DGP/data/.  The reference's `normalize` leaves Y unscaled and never emits `Y_std`, while its own
`utils_dataset.download_UCI_data_info` reads data['Y_std'] and the executed notebook (train_regression_UCI.ipynb cell 1)
prints Y_std = 9.32293 for boston: the notebook behaviour -- the variant kept in Baselines/SGHMC_DGP/datasets.py:74-88,
Y standardised by the training std, `Y_std` emitted -- is what is implemented here.
"""
import os

import numpy as np


class Dataset(object):
    def __init__(self, name, N, D, type='regression', data_path='/data/'):
        assert type in ['regression', 'classification', 'multiclass']
        self.data_path = data_path
        self.name, self.N, self.D = name, N, D
        self.type = type

    def csv_file_path(self, name):
        return '{}{}.csv'.format(self.data_path, name)

    def read_data(self):
        data = np.loadtxt(self.csv_file_path(self.name), delimiter=',', dtype=np.float64, ndmin=2)
        return {'X': data[:, :-1], 'Y': data[:, -1, None]}

    def download_data(self):
        raise FileNotFoundError(f"{self.csv_file_path(self.name)} not found and this build has no network access; copy the "
                                "CSV there (the reference ships them under Baselines/RF_DGP/data/)")

    def get_data(self, seed=0, split=0, prop=0.9):
        if not os.path.isfile(self.csv_file_path(self.name)):
            self.download_data()
        full_data = self.read_data()
        assert full_data['X'].shape == (self.N, self.D), \
            f"{self.name}: CSV holds {full_data['X'].shape}, expected {(self.N, self.D)}"
        split_data = self.split(full_data, seed, split, prop)
        split_data = self.normalize(split_data, 'X')
        if self.type == 'regression':
            split_data = self.normalize(split_data, 'Y')
        return split_data

    def split(self, full_data, seed, split, prop):
        ind = np.arange(self.N)
        np.random.seed(seed + split)          # the legacy global generator, as the reference: the split is reproducible
        np.random.shuffle(ind)
        n = int(self.N * prop)
        return {'X': full_data['X'][ind[:n], :], 'Xs': full_data['X'][ind[n:], :],
                'Y': full_data['Y'][ind[:n], :], 'Ys': full_data['Y'][ind[n:], :]}

    def normalize(self, split_data, X_or_Y):
        m = np.average(split_data[X_or_Y], 0)[None, :]
        s = np.std(split_data[X_or_Y], 0)[None, :] + 1e-6       # statistics of the TRAINING part only
        split_data[X_or_Y] = (split_data[X_or_Y] - m) / s
        split_data[X_or_Y + 's'] = (split_data[X_or_Y + 's'] - m) / s
        split_data.update({X_or_Y + '_mean': m.flatten(), X_or_Y + '_std': s.flatten()})
        return split_data


# name, N, D of the nine UCI sets of the reference (datasets.py:94-234)
_UCI = [('boston', 506, 13), ('concrete', 1030, 8), ('energy', 768, 8), ('kin8nm', 8192, 8), ('naval', 11934, 12),
        ('power', 9568, 4), ('protein', 45730, 9), ('wine_red', 1599, 11), ('wine_white', 4898, 11)]


class Datasets(object):
    def __init__(self, data_path='/data/'):
        self.all_datasets = {name: Dataset(name, N, D, 'regression', data_path) for name, N, D in _UCI}
