"""Wall-clock of the reference's hyper-parameter sweep on the B200 path (BASELINE config 4).

One TASK = one (lead week, bootstrap fold, model) = all 18 trials of tune_2MME.py's grid
(n_blocks [3,4,5] x filters [2,3] x ct_kernel [2,3,5], batch 16, lr 1e-3, <=100 epochs, patience 10) through
`utils.training.train_deepnet(..., training_type="tune")`, then predict x3 and RPSS, on the synthetic C1 data set
(16 years of weekly May-Sep starts, 64x64 grid).  Tasks are independent: `--gpus G` shards them one per GPU
(parallel.sweep, no data-path collective).  The full tune_2MME sweep is 3 leads x 10 folds x 2 models = 60 tasks.

    python tools/sweep_demo.py --tasks 2 --gpus 1
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GRID = {"n_blocks": [3, 4, 5], "n_filters": [2, 3], "ct_kernels": [(2, 2), (3, 3), (5, 5)], "batch_sizes": [16],
        "learning_rates": [1e-3], "patience": 10}


def synth(seed, years=range(2003, 2019), M=4, Y=64, X=64):
    from s2s_ismr_unet_b200.labeled import LabeledArray
    rng = np.random.default_rng(seed)
    T = np.concatenate([pd.date_range(f"{y}-05-01", f"{y}-09-30", freq="7D").values for y in years])
    x = rng.gamma(2.0, 3.0, size=(len(T), M, Y, X)).astype(np.float32)
    y = (0.5 * x.mean(1) + 0.5 * rng.gamma(2.0, 3.0, size=(len(T), Y, X))).astype(np.float32)
    co = {"T": T, "Y": np.arange(Y), "X": np.arange(X)}
    return LabeledArray(x, ("T", "M", "Y", "X"), {**co, "M": np.arange(M)}), LabeledArray(y, ("T", "Y", "X"), co)


def run_task(task):
    """task = (task id, epochs, with_elr).  Returns (seconds, fits, mean test RPSS[, ELR seconds, mean ELR test RPSS]).
    with_elr: the task also runs the ELR baseline of the tune scripts (tune_ECMWF_com.py:53-65: bootstrap_splits_ELR ->
    train_elr) and writes both RPSS maps as NetCDF like tune_ECMWF_com.py:114-121."""
    import contextlib
    import io
    tid, epochs, with_elr = task
    from s2s_ismr_unet_b200.labeled import concat
    from s2s_ismr_unet_b200.utils import preprocessing, training
    x, y = synth(100 + tid)
    splits = preprocessing.bootstrap_splits(x, y, n_bootstraps=1)
    cwd = os.getcwd()
    extra = ()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        if with_elr:
            t0 = time.perf_counter()
            rpss_tr_elr, rpss_te_elr, _, _ = training.train_elr(*preprocessing.bootstrap_splits_ELR(x, y, n_bootstraps=1))
            concat(rpss_te_elr, dim="bootstrap").to_netcdf("ELR_rpss_test_wk3-4.nc")
            extra = (time.perf_counter() - t0, float(np.nanmean(rpss_te_elr[0].values)))
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            out = training.train_deepnet(*splits, training_type="tune", architecture="unet", tuning_grid=GRID, predictor="mean",
                                         obs="IMD", modname=f"M{tid}", week="wk3-4", epochs=epochs, batch_size=16, dir="S/")
        concat(out[2], dim="bootstrap").to_netcdf("unet_rpss_test_wk3-4.nc")
        dt = time.perf_counter() - t0
        os.chdir(cwd)
    return (dt, 18, float(np.nanmean(out[2][0].values))) + extra


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=2)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--elr", action="store_true", help="also run the ELR baseline (batched IRLS kernel) in every task")
    a = ap.parse_args()
    tasks = [(i, a.epochs, a.elr) for i in range(a.tasks)]
    t0 = time.perf_counter()
    if a.gpus > 1:
        from s2s_ismr_unet_b200.parallel import sweep
        res = sweep(tasks, run_task, a.gpus)
    else:
        res = [run_task(t) for t in tasks]
    wall = time.perf_counter() - t0
    per_task = float(np.mean([r[0] for r in res]))
    print({"tasks": a.tasks, "gpus": a.gpus, "fits": sum(r[1] for r in res), "wall_s": round(wall, 2), "s_per_task": round(per_task, 2),
           "s_per_fit": round(per_task / 18, 3), "mean_test_rpss": [round(r[2], 4) for r in res],
           "full_tune_2MME_60_tasks_on_8_gpus_s": round(60 * per_task / 8, 1), "full_tune_2MME_on_1_gpu_s": round(60 * per_task, 1),
           **({"elr_s_per_task": round(float(np.mean([r[3] for r in res])), 3),
               "elr_mean_test_rpss": [round(r[4], 4) for r in res]} if a.elr else {})})
