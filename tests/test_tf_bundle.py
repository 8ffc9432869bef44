"""TensorBundle (SavedModel variables/) reader + mapping onto the engine's weight table. CPU only.
No TensorFlow exists here: the reader is checked against the in-repo writer, against a table block
with prefix-compressed keys built by hand, and against the published crc32c test vector."""
import os
import struct

import numpy as np
import pytest

from frame_interpolation_b200 import spec, tf_bundle, weights


def test_crc32c_known_answer():
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283      # CRC-32C (Castagnoli) check value
    assert tf_bundle.crc32c(b"") == 0


def test_varint_and_proto_wire():
    assert tf_bundle._put_varint(300) == b"\xac\x02"
    assert tf_bundle._varint(b"\xac\x02", 0) == (300, 2)
    msg = tf_bundle._proto_field(1, 0, 1) + tf_bundle._proto_field(5, 0, 150) + tf_bundle._proto_field(6, 5, 0xDEADBEEF)
    p = tf_bundle._parse_proto(msg)
    assert p == {1: [1], 5: [150], 6: [0xDEADBEEF]}


def test_block_with_shared_key_prefixes():
    # entries "abc"->"1", "abd"->"22" (shares "ab"), restart array [0], num_restarts 1
    body = bytes([0, 3, 1]) + b"abc" + b"1" + bytes([2, 1, 2]) + b"d" + b"22"
    blk = body + struct.pack("<I", 0) + struct.pack("<I", 1)
    data = blk + b"\x00" + b"\x00\x00\x00\x00"
    assert tf_bundle._read_block(data, 0, len(blk)) == [(b"abc", b"1"), (b"abd", b"22")]
    with pytest.raises(ValueError):
        tf_bundle._read_block(blk + b"\x01" + b"\x00\x00\x00\x00", 0, len(blk))   # snappy flag


def test_bundle_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    t = {"a/kernel": rng.standard_normal((3, 3, 2, 4)).astype(np.float32),
         "a/bias": rng.standard_normal(4).astype(np.float32),
         "step": np.array(7, np.int64)}
    prefix = str(tmp_path / "variables" / "variables")
    tf_bundle.write_bundle(prefix, t, with_crc=True)
    idx = tf_bundle.read_index(prefix + ".index")
    assert set(idx) == set(t) and idx["a/kernel"]["shape"] == (3, 3, 2, 4) and idx["a/kernel"]["dtype"] == 1
    back = tf_bundle.read_bundle(prefix)
    for k in t:
        np.testing.assert_array_equal(back[k], t[k])
    with open(prefix + ".index", "r+b") as f:
        f.seek(-8, os.SEEK_END)
        f.write(b"\x00" * 8)
    with pytest.raises(ValueError, match="magic"):
        tf_bundle.read_index(prefix + ".index")


def _object_graph_keys(w):
    """The reference's attribute paths (model.save object-graph checkpoint style)."""
    out = {}
    suf = "/.ATTRIBUTES/VARIABLE_VALUE"
    for name, _ in spec.weight_table():
        parts = name.split("/")
        kind = parts[-1]
        if parts[0] == "feat_net":
            k = int(parts[2].split("_")[-1])
            key = f"layer_with_weights-0/extract_sublevels/convs/{k}/{kind}"
        elif parts[0] == "predict_flow":
            p = spec.FLOW_PREDICTOR_NAMES.index(parts[1])
            key = f"layer_with_weights-1/_predictors/{p}/_convs/{int(parts[2].split('_')[-1])}/{kind}"
        elif parts[1] == "output_conv":
            key = f"layer_with_weights-2/output_conv/{kind}"
        else:
            key = f"layer_with_weights-2/convs/{int(parts[1].split('_')[-1])}/{int(parts[2].split('_')[-1])}/{kind}"
        out[key + suf] = w[name]
    out["optimizer/iter" + suf] = np.array(3, np.int64)
    out["_CHECKPOINTABLE_OBJECT_GRAPH"] = np.zeros(4, np.uint8)
    return out


def test_saved_model_conversion_object_graph_keys(tmp_path):
    w = weights.synthetic_weights(11)
    sm = tmp_path / "saved_model"
    tf_bundle.write_bundle(str(sm / "variables" / "variables"), _object_graph_keys(w))
    out = tf_bundle.convert_saved_model(str(sm), str(tmp_path / "style.filmw"))
    assert weights.digest(weights.load(out)) == weights.digest(w)


def test_saved_model_conversion_variable_name_keys(tmp_path):
    w = weights.synthetic_weights(12)
    named, k = {}, 0
    for name, _ in spec.weight_table():
        if not name.endswith("/kernel"):
            continue
        base = name[:-len("/kernel")]
        tf_name = base
        if base.startswith("fusion/"):
            tf_name = "fusion/conv2d" + ("" if k == 0 else f"_{k}")
            k += 1
        named[tf_name + "/kernel"] = w[base + "/kernel"]
        named[tf_name + "/bias"] = w[base + "/bias"]
    prefix = str(tmp_path / "v" / "variables")
    tf_bundle.write_bundle(prefix, named)
    got = tf_bundle.film_weights_from_bundle(tf_bundle.read_bundle(prefix))
    assert weights.digest(got) == weights.digest(w)
    del named["fusion/conv2d_3/kernel"]
    tf_bundle.write_bundle(prefix, named)
    with pytest.raises(ValueError, match="missing"):
        tf_bundle.film_weights_from_bundle(tf_bundle.read_bundle(prefix))
