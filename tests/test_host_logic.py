"""Host-side logic that needs no GPU: layout helpers, bootstrap splits, the rolling tercile labeler,
callbacks, the DP/sweep partitioning helpers and the bench's algorithmic-work accounting."""
import numpy as np
import pandas as pd
import pytest

from oracle import skill as so
from s2s_ismr_unet_b200.keras_api.callbacks import EarlyStopping, ModelCheckpoint
from s2s_ismr_unet_b200.keras_api.utils import to_categorical
from s2s_ismr_unet_b200.labeled import LabeledArray
from s2s_ismr_unet_b200.parallel import assign_tasks, shard_batch, shard_weight, trial_cost
from s2s_ismr_unet_b200.utils import preprocessing as pp


def make_xy(years=range(2003, 2019), M=4, Y=8, X=8, seed=0):
    rng = np.random.default_rng(seed)
    T = np.concatenate([pd.date_range(f"{y}-05-01", f"{y}-09-30", freq="7D").values for y in years])
    x = rng.gamma(2.0, 3.0, size=(len(T), M, Y, X)).astype(np.float32)
    y = rng.gamma(2.0, 3.0, size=(len(T), Y, X)).astype(np.float32)
    y[:, 0, 0] = np.nan                                            # an ocean point
    co = {"T": T, "Y": np.arange(Y), "X": np.arange(X)}
    return LabeledArray(x, ("T", "M", "Y", "X"), {**co, "M": np.arange(M)}), LabeledArray(y, ("T", "Y", "X"), co)


def test_bootstrap_splits_are_year_wise_70_20_10_and_seeded():
    x, y = make_xy()
    out = pp.bootstrap_splits(x, y, n_bootstraps=3)
    assert all(len(l) == 3 for l in out)
    for i in range(3):
        yrs = [set(pd.DatetimeIndex(out[k][i]["T"]).year) for k in (0, 2, 4)]
        assert len(yrs[0]) == 12 and len(yrs[1]) == 3 and len(yrs[2]) == 1         # 16 years: int(.2*16)=3, int(.1*16)=1
        assert not (yrs[0] & yrs[1]) and not (yrs[0] & yrs[2]) and not (yrs[1] & yrs[2])
        np.random.seed(i)
        perm = np.random.permutation(np.arange(2003, 2019))
        assert yrs[1] == set(perm[:3]) and yrs[2] == set(perm[3:4])
        assert not np.isnan(out[1][i].values).any()                                   # fillna(0)
        assert np.all(np.diff(out[0][i]["T"].astype("int64")) > 0)                    # sorted by T
        assert len(out[0][i]) == len(out[1][i])


def test_bootstrap_splits_mme_share_the_year_split():
    x, y = make_xy()
    x2, _ = make_xy(seed=1)
    xtr, xva, xte, ytr, yva, yte = pp.bootstrap_splits_mme({"GEFS": x, "IITM": x2}, y, n_bootstraps=2)
    for i in range(2):
        assert np.array_equal(xtr["GEFS"][i]["T"], xtr["IITM"][i]["T"]) and np.array_equal(xtr["GEFS"][i]["T"], ytr[i]["T"])
        assert len(xva["GEFS"][i]) == len(yva[i]) and len(xte["IITM"][i]) == len(yte[i])


def test_to_categorical():
    oh = to_categorical(np.array([[0.0, 2.0], [1.0, np.nan]]), 3)
    assert oh.dtype == np.float32 and oh.shape == (2, 2, 3)
    np.testing.assert_array_equal(oh[0, 1], [0, 0, 1])
    np.testing.assert_array_equal(oh[1, 1], [1, 0, 0])


class FakeModel:
    def __init__(self):
        self.w, self.stop_training, self.optimizer, self.config = 0, False, None, {}

    def get_weights(self):
        return {"w": np.array([self.w])}

    def set_weights(self, w):
        self.w = int(w["w"][0])


def run_callbacks(vals, cbs):
    m = FakeModel()
    for cb in cbs:
        cb.set_model(m)
        cb.on_train_begin()
    n = 0
    for ep, v in enumerate(vals):
        m.w = ep
        n += 1
        for cb in cbs:
            cb.on_epoch_end(ep, {"val_loss": v})
        if m.stop_training:
            break
    for cb in cbs:
        cb.on_train_end()
    return m, n


def test_early_stopping_keras3_semantics():
    vals = [1.0, 0.8, 0.9, 0.85, 0.81, 0.7, 0.9, 0.9, 0.9, 0.9]
    m, n = run_callbacks(vals, [EarlyStopping(monitor="val_loss", patience=3, restore_best_weights=True)])
    assert n == 5 and m.w == 1                 # best at epoch 1, waits 3 non-improving epochs, restores epoch-1 weights
    m, n = run_callbacks(vals, [EarlyStopping(monitor="val_loss", patience=4, restore_best_weights=True)])
    assert n == 10 and m.w == 5                # epoch 5 improves in time; 4 bad epochs end at the last one; best restored
    m, n = run_callbacks([3.0, 2.0, 1.0], [EarlyStopping(patience=1, restore_best_weights=True)])
    assert n == 3 and m.w == 2                 # Keras 3 restores the best weights at train end even without an early stop


def test_model_checkpoint_tracks_the_best_epoch(tmp_path, monkeypatch):
    saved = {}
    import s2s_ismr_unet_b200.model as mod
    monkeypatch.setattr(mod, "save_weights_file", lambda path, cfg, w, opt, o: saved.update(path=path, w=w))
    ck = ModelCheckpoint(str(tmp_path / "a" / "best.keras"), save_best_only=True, monitor="val_loss", mode="min")
    run_callbacks([1.0, 0.5, 0.7, 0.6], [ck])
    assert saved["path"].endswith("best.keras") and int(saved["w"]["w"][0]) == 1


def test_dp_partition_helpers():
    for n, world in [(16, 8), (5, 2), (7, 4), (3, 8)]:
        parts = [shard_batch(n, r, world) for r in range(world)]
        covered = np.concatenate([np.arange(n)[p] for p in parts])
        np.testing.assert_array_equal(covered, np.arange(n))
        assert abs(sum(shard_weight(p.stop - p.start, n) for p in parts) - 1.0) < 1e-12
    costs = [trial_cost(f, nb, k) for nb in (3, 4, 5) for f in (2, 3) for k in (2, 3, 5)]
    plan = assign_tasks(costs, 8)
    assert sorted(i for p in plan for i in p) == list(range(18))
    loads = [sum(costs[i] for i in p) for p in plan]
    assert max(loads) <= 1.05 * max(max(costs), sum(costs) / 8 * 1.34)
    assert abs(trial_cost(2, 3, 3) - 228.85) < 3.0            # 228.85 train MFLOP/sample (C=1)


def test_bench_algorithmic_work_matches_survey():
    import bench
    fl, by = bench.algorithmic_work(dict(H=64, W=64, Cin=1, filters=2, n_blocks=3, ct_kernel=3))
    assert abs(fl / 1e6 - 228.85) < 0.5 and abs(by / 1e6 - 8.307) < 0.05
    fl, by = bench.algorithmic_work(dict(H=64, W=64, Cin=3, filters=2, n_blocks=3, ct_kernel=3))
    assert abs(fl / 1e6 - 232.39) < 0.5 and abs(by / 1e6 - 8.405) < 0.05


def test_unet_mirror_validates_like_the_reference():
    from s2s_ismr_unet_b200.utils.deep_nn_models import Unet
    u = Unet("", ct_kernel=(5, 5), n_blocks=4, filters=3)
    assert (u.filters, u.n_blocks, u.ct_kernel, u.apool, u.bn, u.bs, u.learn_rate) == (3, 4, (5, 5), True, True, 16, 1e-4)
    with pytest.raises(ValueError, match="not divisible"):
        u.build_model((24, 24, 1))


def test_netcdf3_round_trip_of_the_rpss_outputs(tmp_path):
    """outputs/<dir><model>_<obs>/unet_rpss_test_<week>.nc (tune_ECMWF_com.py:114-121): concat over 'bootstrap', write, read."""
    from s2s_ismr_unet_b200.labeled import concat, open_netcdf
    rng = np.random.default_rng(0)
    maps = []
    for b in range(3):
        v = rng.normal(size=(5, 7)).astype(np.float32)
        v[0, 0] = np.nan
        maps.append(LabeledArray(v, ("Y", "X"), {"Y": np.linspace(6.5, 38.5, 5), "X": np.linspace(66.5, 100.0, 7)}))
    rpss = concat(maps, dim="bootstrap")
    assert rpss.dims == ("bootstrap", "Y", "X") and rpss.shape == (3, 5, 7)
    path = tmp_path / "unet_rpss_test_wk3-4.nc"
    rpss.to_netcdf(path)
    back = open_netcdf(path)
    assert back.dims == rpss.dims
    np.testing.assert_array_equal(back.values, rpss.values)
    np.testing.assert_allclose(back["Y"], rpss["Y"])
    # predictions (T, Y, X, category) with datetime starts and string categories, concatenated along T
    T = pd.date_range("2018-06-01", periods=4, freq="7D").values
    p1 = LabeledArray(rng.random((4, 2, 2, 3)), ("T", "Y", "X", "category"),
                      {"T": T, "Y": np.arange(2), "X": np.arange(2), "category": np.array(["below", "normal", "above"])})
    p2 = p1._like(p1.values + 1.0)
    p2.coords["T"] = T + np.timedelta64(28, "D")
    both = concat([p1, p2], dim="T")
    assert both.shape == (8, 2, 2, 3)
    both.to_netcdf(tmp_path / "p.nc", name="prediction")
    back = open_netcdf(tmp_path / "p.nc")
    np.testing.assert_array_equal(back["T"], both["T"].astype("datetime64[ns]"))
    assert list(back["category"]) == ["below", "normal", "above"]
    np.testing.assert_allclose(back.values, both.values)


def test_library_stamp_is_the_hash_of_the_sources_the_objects_were_compiled_from():
    """build.py stamps libs2s_unet.so with the unit hashes taken BEFORE nvcc ran (a source edited during the 6-minute compile must
    leave the library stale instead of stamped with text the compiler never saw)."""
    from s2s_ismr_unet_b200 import build as b
    units = [b._unit_hash(src) for src in b.sources()]
    assert len(units) >= 6 and all(len(u) == 64 for u in units)
    assert b._source_hash(units) == b._source_hash()                     # nothing edited in between: same stamp
    edited = list(units)
    edited[0] = "0" * 64
    assert b._source_hash(edited) != b._source_hash()
    # every header a unit includes is part of its hash (gconv.cuh is shared by three translation units)
    deps = {src.name: {d.name for d in b._deps(src)} for src in b.sources()}
    assert "gconv.cuh" in deps["gconv.cu"] and "gconv.cuh" in deps["unet.cu"] and "s2s_unet.h" in deps["unet.cu"]
