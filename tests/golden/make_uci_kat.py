"""Known answers for the data layer.  Two sources:
  * `notebook`: values PRINTED by the reference's executed notebook experiments/train_regression_UCI.ipynb cell 1
    (boston: D 13, N 455, Ns 51, X_mean[13], Y_mean, Y_std) -- copied from its stored stdout;
  * `derived`: what this build's experiments/datasets.py computes from the reference's boston.csv
    (Baselines/RF_DGP/data/boston.csv), so the GPU box (no /root/reference) can still check a second implementation
    against frozen numbers if the CSV is supplied.
Run in the build container: python tests/golden/make_uci_kat.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dgp-rf-mcmc_b200"))
from experiments.datasets import Datasets

nb = json.load(open("/root/reference/experiments/train_regression_UCI.ipynb"))
text = "".join("".join(o.get("text", "")) for o in nb["cells"][1]["outputs"])
d = Datasets("/root/reference/Baselines/RF_DGP/data/").all_datasets["boston"].get_data()
kat = {"notebook_stdout": text,
       "notebook": {"D": 13, "N": 455, "Ns": 51, "Y_mean": 22.656263, "Y_std": 9.32293,
                    "X_mean": [1.7378345e+00, 1.1138461e+01, 1.1224440e+01, 7.2527476e-02, 5.5657738e-01, 6.2923098e+00,
                               6.9054504e+01, 3.6594815e+00, 4.3472528e+00, 4.0854724e+02, 1.8481098e+01, 3.5551254e+02,
                               1.2646418e+01]},
       "derived": {"X_row0": [float(v) for v in np.float32(d["X"][0])], "Y_head": [float(v) for v in np.float32(d["Y"][:5, 0])],
                   "Xs_row0": [float(v) for v in np.float32(d["Xs"][0])], "X_std": [float(v) for v in np.float32(d["X_std"])]}}
json.dump(kat, open(os.path.join(os.path.dirname(__file__), "uci_boston_kat.json"), "w"), indent=1)
print("written", text[:200])
