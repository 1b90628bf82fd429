"""Keras-3 `.keras` archive interop: the minimal HDF5 reader / writer and the archive layout
(layers/<snake_case(class)[_k]>/vars/<i>, optimizer/vars/<i>) that keras.models.load_model / model.save use
(reference: utils/training.py:114-115, 127-131).  h5py is not installable here, so the container is checked
structurally against the HDF5 specification and by round trip through the independent reader."""
import io
import json
import struct
import zipfile

import numpy as np
import pytest

from oracle import keras_unet as ko
from s2s_ismr_unet_b200.keras_api import hdf5_min as h5
from s2s_ismr_unet_b200.keras_api import keras_archive as ka


def _same(a, b, path=""):
    assert set(a) == set(b), (path, set(a) ^ set(b))
    for k in a:
        if isinstance(a[k], dict):
            _same(a[k], b[k], f"{path}/{k}")
        else:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), f"{path}/{k}"


def test_hdf5_round_trip_groups_scalars_and_many_links():
    rng = np.random.default_rng(0)
    tree = {"layers": {f"conv2d_{i}": {"vars": {"0": rng.normal(size=(3, 3, 4, 8)).astype(np.float32),
                                                "1": rng.normal(size=(8,)).astype(np.float32)}} for i in range(37)},
            "optimizer": {"vars": {str(i): rng.normal(size=(i % 5 + 1, 3)).astype(np.float32) for i in range(2, 160)}},
            "vars": {}, "f64": np.linspace(0, 1, 7), "u8": np.arange(5, dtype=np.uint8), "i32": np.arange(-3, 3, dtype=np.int32).reshape(2, 3)}
    tree["optimizer"]["vars"]["0"] = np.asarray(123456789012, np.int64)          # Adam's `iteration` is a 0-d int64
    tree["optimizer"]["vars"]["1"] = np.asarray(1e-3, np.float32)
    blob = h5.write_hdf5(tree)
    _same(tree, h5.read_hdf5(blob))


def test_hdf5_container_follows_the_version_0_format():
    blob = h5.write_hdf5({"g": {"d": np.arange(6, dtype=np.float32).reshape(2, 3)}})
    assert blob[:8] == b"\x89HDF\r\n\x1a\n"
    sb_ver, _, root_ver, _, _, off_size, len_size = struct.unpack_from("<7B", blob, 8)
    assert (sb_ver, root_ver, off_size, len_size) == (0, 0, 8, 8)
    leaf_k, internal_k = struct.unpack_from("<HH", blob, 16)
    assert (leaf_k, internal_k) == (4, 16)
    base, free, eof, drv = struct.unpack_from("<4Q", blob, 24)
    assert base == 0 and free == h5.UNDEF and drv == h5.UNDEF and eof == len(blob)
    name_off, ohdr, cache, _ = struct.unpack_from("<QQII", blob, 56)
    btree, heap = struct.unpack_from("<QQ", blob, 56 + 24)
    assert cache == 1 and blob[btree:btree + 4] == b"TREE" and blob[heap:heap + 4] == b"HEAP" and blob[ohdr] == 1
    # local heap: the free-list head points at a well-formed last free block (next = 1 = H5HL_FREE_NULL)
    seg_size, free_off, seg_addr = struct.unpack_from("<QQQ", blob, heap + 8)
    nxt, fsize = struct.unpack_from("<QQ", blob, seg_addr + free_off)
    assert nxt == 1 and free_off + fsize == seg_size
    # the raw data are little-endian IEEE and contiguous
    assert np.frombuffer(blob, "<f4").tobytes().find(np.arange(6, dtype="<f4").tobytes()) >= 0


def test_hdf5_reader_rejects_what_it_cannot_read():
    with pytest.raises(ValueError, match="signature"):
        h5.read_hdf5(b"not an hdf5 file at all........................")
    blob = bytearray(h5.write_hdf5({"d": np.zeros(3, np.float32)}))
    blob[8] = 2                                           # superblock version 2 (libver='latest')
    with pytest.raises(NotImplementedError, match="superblock"):
        h5.read_hdf5(bytes(blob))


@pytest.mark.parametrize("kw", [dict(), dict(n_blocks=5, filters=3, ct_kernel=5), dict(apool=False, bn=False, Cin=2), dict(head="deterministic")])
def test_keras_archive_round_trip(tmp_path, kw):
    ocfg = ko.UnetConfig(H=32, W=32, **({"Cin": 3} | kw))
    cfg = dict(input_shape=[32, 32, ocfg.Cin], filters=ocfg.filters, n_blocks=ocfg.n_blocks, ct_kernel=ocfg.ct_kernel, apool=ocfg.apool,
               bn=ocfg.bn, output=ocfg.head, activation="elu")
    w = ko.random_init(ocfg, 3)
    path = tmp_path / "best_model_unet_0.keras"
    ka.write_keras_archive(path, cfg, w)
    with zipfile.ZipFile(path) as z:
        assert {"config.json", "metadata.json", "model.weights.h5"} <= set(z.namelist())
        kc = json.loads(z.read("config.json"))
        tree = h5.read_hdf5(z.read("model.weights.h5"))
    assert kc["class_name"] == "Functional" and json.loads(zipfile.ZipFile(path).read("metadata.json"))["keras_version"].startswith("3.")
    # Keras' path layout: class-name-indexed layer groups in model.layers order, variables numbered per layer
    names = [L["config"]["name"] for L in kc["config"]["layers"]]
    assert names[0] == "input_layer" and "bottleneck" in names and names[1] == "down_conv1_1"
    assert "conv2d" in tree["layers"] and "conv2d_transpose" in tree["layers"] and set(tree["layers"]["conv2d"]["vars"]) == {"0", "1"}
    assert tree["layers"]["conv2d"]["vars"]["0"].shape == (3, 3, ocfg.Cin, ocfg.filters * 4)
    if ocfg.bn:
        assert set(tree["layers"]["batch_normalization"]["vars"]) == {"0", "1", "2", "3"}
    assert tree["layers"]["dropout"]["vars"] == {}
    ar = ka.read_keras_archive(path)
    assert ar["config"] == cfg
    assert list(ar["weights"]) == [n for n, _, _ in ko.param_specs(ocfg)]
    for k in w:
        np.testing.assert_array_equal(ar["weights"][k], w[k].astype(np.float32))


def test_keras_archive_carries_the_adam_state(tmp_path):
    from s2s_ismr_unet_b200.keras_api.optimizers import Adam
    from s2s_ismr_unet_b200.model import _trainable_layout
    ocfg = ko.UnetConfig(H=16, W=16, Cin=1)
    cfg = dict(input_shape=[16, 16, 1], filters=2, n_blocks=3, ct_kernel=3, apool=True, bn=True, output="proba", activation="elu")
    w = ko.random_init(ocfg, 4)
    lay = _trainable_layout(w)
    n = lay[-1]["offset"] + (lay[-1]["count"] + 3) // 4 * 4
    rng = np.random.default_rng(5)
    opt = dict(m=rng.normal(size=n).astype(np.float32), v=rng.random(n).astype(np.float32), step=17)
    path = tmp_path / "m.keras"
    ka.write_keras_archive(path, cfg, w, opt, Adam(learning_rate=1e-4), lay)
    ar = ka.read_keras_archive(path)
    assert ar["optimizer"]["learning_rate"] == pytest.approx(1e-4) and ar["opt_vars"]["step"] == 17
    for d in lay:       # Keras order: iteration, learning_rate, then (momentum, velocity) per trainable variable
        sl = slice(d["offset"], d["offset"] + d["count"])
        np.testing.assert_array_equal(ar["opt_vars"]["m"][d["name"]].ravel(), opt["m"][sl])
        np.testing.assert_array_equal(ar["opt_vars"]["v"][d["name"]].ravel(), opt["v"][sl])
    tree = h5.read_hdf5(zipfile.ZipFile(path).read("model.weights.h5"))
    assert tree["optimizer"]["vars"]["0"].dtype == np.int64 and tree["optimizer"]["vars"]["0"].shape == ()
    assert len(tree["optimizer"]["vars"]) == 2 + 2 * len(lay)


def test_foreign_archives_are_rejected_with_a_reason(tmp_path):
    path = tmp_path / "other.keras"
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("config.json", json.dumps({"class_name": "Sequential", "config": {"layers": []}}))
        z.writestr("model.weights.h5", h5.write_hdf5({"layers": {}}))
    with pytest.raises(ValueError, match="functional"):
        ka.read_keras_archive(path)
