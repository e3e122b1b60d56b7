"""Extract the per-cycle training printouts of the reference's executed notebook experiments/train_regression_demo_sin.ipynb
(cells 7 and 13: `regression_train_demo` prints, at the end of every cycle, the mean train / test log-likelihood and the
train / test RMSE of the current posterior sample -- experiments/utils_training_demo.py:63-71) into
tests/golden/notebook_sin_demo_traces.json.  Runs in the build container only (reads /root/reference); the fixture travels.

These are the only tensor-derived numbers the reference ever printed for the sampling path: 1040 cycles of a real
TensorFlow run of models/dgp.py.  The data set of the notebook is unseeded, so they pin the sampler statistically (a band),
and -- pair by pair -- the Gaussian likelihood exactly: every printed (log-likelihood, RMSE) pair obeys
    LL = -0.5 log(2 pi var) - RMSE^2 / (2 var)        (likelihoods/gaussian.py:20-25 with the notebook's variance 0.01).
"""
import json
import os
import re

NB = "/root/reference/experiments/train_regression_demo_sin.ipynb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "notebook_sin_demo_traces.json")
RUNS = {
    # cell index of the training call -> the settings of the cells above it (cells 4-6 / 10-12) and of cell 2 (data)
    7: dict(n_hidden_layers=1, n_rf=300, n_gp=1, lik_variance=0.01, lr_0=0.02, momentum_decay=0.99, total_epochs=2000,
            start_sampling_epoch=0, epochs_per_cycle=50, resample_in_cycle_head=True),
    13: dict(n_hidden_layers=2, n_rf=100, n_gp=1, lik_variance=0.01, lr_0=0.01, momentum_decay=0.95, total_epochs=50000,
             start_sampling_epoch=0, epochs_per_cycle=50, resample_in_cycle_head=True),
}
DATA = dict(num_training=60, num_testing=100, batch_size=20, std_noise=0.02,
            recipe="X ~ U(-2,-1) (30 points) and U(1,2) (30 points); Y = sin(pi X) + N(0, std_noise^2), standardised; "
                   "X_test = linspace(-5, 5, 100), Y_test = (sin(pi X_test) - y_mean) / y_std  (notebook cell 1)")

if __name__ == "__main__":
    nb = json.load(open(NB))
    out = {"_provenance": "stdout stored in " + NB.replace("/root/reference/", "") + " (executed by the reference's authors, TensorFlow 2.x on an RTX A5000)",
           "data": DATA, "runs": []}
    for ci, cfg in RUNS.items():
        text = "".join("".join(o.get("text", "")) for o in nb["cells"][ci]["outputs"] if "text" in o)
        ll = re.findall(r"Mean Log Likelihood -- train: ([-\d.e+]+), -- test: ([-\d.e+]+)", text)
        rm = re.findall(r"Root Mean Squared Error -- train: ([-\d.e+]+), -- test: ([-\d.e+]+)", text)
        fin = re.search(r"Number of sampled models: (\d+)\s+Test Log Likelihood of all sampled models: ([-\d.e+]+)\s+"
                        r"Test Root MSE of all sampled models: ([-\d.e+]+)", text)
        assert len(ll) == len(rm) == cfg["total_epochs"] // cfg["epochs_per_cycle"]
        out["runs"].append({"cell": ci, "config": cfg,
                            "train_ll": [float(a) for a, _ in ll], "test_ll": [float(b) for _, b in ll],
                            "train_rmse": [float(a) for a, _ in rm], "test_rmse": [float(b) for _, b in rm],
                            "n_models": int(fin.group(1)), "ensemble_test_ll": float(fin.group(2)),
                            "ensemble_test_rmse": float(fin.group(3))})
    json.dump(out, open(OUT, "w"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
