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


def _fixture_keys():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "film_style_object_graph_keys.tsv")
    rows = []
    for line in open(path):
        if line.startswith("#") or not line.strip():
            continue
        key, dtype, shape = line.rstrip("\n").split("\t")
        shape = () if shape in ("()", "") else tuple(int(d) for d in shape.split(","))
        rows.append((key, dtype, shape))
    return rows


def test_object_graph_key_list_fixture_maps_onto_the_whole_weight_table(tmp_path):
    """The committed key list (object-graph style: `layer_with_weights-N/<attribute path>/.ATTRIBUTES/VARIABLE_VALUE`,
    plus optimizer slots, counters and the object-graph proto) is written as a real TensorBundle with seeded values and
    read back through the importer: every one of the 82 engine tensors must be found, with the right shape, from the
    right key, and nothing else may leak in."""
    rng = np.random.default_rng(5)
    tensors, by_key = {}, {}
    for key, dtype, shape in _fixture_keys():
        if dtype == "string":
            tensors[key] = np.zeros(8, np.uint8)          # stand-in for the serialized object graph
        elif dtype == "int64":
            tensors[key] = np.array(3, np.int64)
        else:
            tensors[key] = rng.standard_normal(shape).astype(np.float32) if shape else np.array(0.5, np.float32)
    prefix = str(tmp_path / "saved_model" / "variables" / "variables")
    tf_bundle.write_bundle(prefix, tensors, with_crc=True)
    got = tf_bundle.film_weights_from_bundle(tf_bundle.read_bundle(prefix))
    table = dict(spec.weight_table())
    assert set(got) == set(table) and len(got) == 82
    for name, shape in table.items():
        assert got[name].shape == tuple(shape)
    # spot checks of the attribute-path mapping, including the shared predictor (list index 3)
    suf = "/.ATTRIBUTES/VARIABLE_VALUE"
    np.testing.assert_array_equal(got["feat_net/sub_extractor/cfeat_conv_5/kernel"],
                                  tensors["layer_with_weights-0/extract_sublevels/convs/5/kernel" + suf])
    np.testing.assert_array_equal(got["predict_flow/flow_predictor_shared/conv_4/bias"],
                                  tensors["layer_with_weights-1/_predictors/3/_convs/4/bias" + suf])
    np.testing.assert_array_equal(got["fusion/level_2/conv_0/kernel"],
                                  tensors["layer_with_weights-2/convs/2/0/kernel" + suf])
    np.testing.assert_array_equal(got["fusion/output_conv/kernel"], tensors["layer_with_weights-2/output_conv/kernel" + suf])
    # the optimizer slot of cfeat_conv_0 must not have replaced the variable itself
    np.testing.assert_array_equal(got["feat_net/sub_extractor/cfeat_conv_0/kernel"],
                                  tensors["layer_with_weights-0/extract_sublevels/convs/0/kernel" + suf])


def _variable_name_keys(w, first_counter):
    """Plain variable names; the fusion convs carry Keras auto-names `conv2d_<n>` from a GLOBAL counter."""
    named, k = {}, first_counter
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
    return named


@pytest.mark.parametrize("first_counter", [0, 7, 98])
def test_keras_auto_names_with_a_counter_that_does_not_start_at_zero(tmp_path, first_counter):
    """`fusion/conv2d_7/kernel` ... `fusion/conv2d_19/kernel`: the counter is global to the Keras session (98 -> 110
    also crosses a digit boundary, so a lexicographic sort would scramble the order)."""
    w = weights.synthetic_weights(13)
    prefix = str(tmp_path / "v" / "variables")
    tf_bundle.write_bundle(prefix, _variable_name_keys(w, first_counter))
    got = tf_bundle.film_weights_from_bundle(tf_bundle.read_bundle(prefix))
    assert weights.digest(got) == weights.digest(w)


def test_mixed_key_styles_are_merged_not_all_or_nothing(tmp_path):
    """Fusion tensors under object-graph keys, feature/flow tensors under plain variable names: the importer must
    take each group from where it finds it (the fallback used to drop the named group)."""
    w = weights.synthetic_weights(14)
    mixed = {k: v for k, v in _object_graph_keys(w).items() if "layer_with_weights-2" in k}
    for name, _ in spec.weight_table():
        if not name.startswith("fusion/"):
            mixed["film_net/" + name + ":0"] = w[name]
    prefix = str(tmp_path / "m" / "variables")
    tf_bundle.write_bundle(prefix, mixed)
    got = tf_bundle.film_weights_from_bundle(tf_bundle.read_bundle(prefix))
    assert weights.digest(got) == weights.digest(w)
    # a reordered Keras counter is caught by the shape check instead of silently permuting layers
    named = _variable_name_keys(w, 0)
    named["fusion/conv2d_1/kernel"], named["fusion/conv2d_2/kernel"] = named["fusion/conv2d_2/kernel"], named["fusion/conv2d_1/kernel"]
    with pytest.raises(ValueError):
        weights.from_named_arrays(named)
