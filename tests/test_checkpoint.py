"""TensorFlow V2 checkpoint reader (tf_faster_rcnn_b200/checkpoint.py): what `saver.restore(sess, model)` of the
reference's tools (tools/demo.py:139-140, tools/test_net.py:111-113) resolves to here.  No TensorFlow in this image, so
the anchors are the published CRC-32C vectors, the LevelDB table magic, a byte-level hand-assembled index, and
round trips through this module's writer."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tf_faster_rcnn_b200 import checkpoint as ck, synth  # noqa: E402


def test_crc32c_known_answers():
    assert ck.crc32c(b"") == 0
    assert ck.crc32c(b"123456789") == 0xE3069283                    # CRC-32C check value
    assert ck.crc32c(b"\x00" * 32) == 0x8A9136AA                    # RFC 3720 B.4
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.unmask_crc(ck.mask_crc(0x12345678)) == 0x12345678
    assert ck.mask_crc(0) == 0xa282ead8


def test_crc32c_chunked_path_matches_bytewise():
    rng = np.random.default_rng(3)
    for n in (4096 * 8 - 1, 4096 * 8, 4096 * 8 + 1, 100003):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert ck.crc32c(data) == ck._crc_update_scalar(0xFFFFFFFF, data) ^ 0xFFFFFFFF


def _hand_built_index():
    """One float32 [2] tensor "w" at offset 0, assembled byte by byte from the format description (not via write_bundle)."""
    payload = struct.pack("<2f", 1.5, -2.0)
    entry = bytes([0x08, 0x01,                       # dtype = DT_FLOAT
                   0x12, 0x04, 0x12, 0x02, 0x08, 0x02,   # shape { dim { size: 2 } }
                   0x28, 0x08]) + b"\x35" + struct.pack("<I", ck.mask_crc(ck.crc32c(payload)))   # size = 8, crc32c fixed32
    header = bytes([0x08, 0x01, 0x1a, 0x02, 0x08, 0x01])
    block = (bytes([0, 0, len(header)]) + header +               # key "" (shared 0, non_shared 0)
             bytes([0, 1, len(entry)]) + b"w" + entry +
             struct.pack("<II", 0, 1))                          # one restart at 0
    out = bytearray()

    def emit(contents):
        off = len(out)
        out.extend(contents + b"\x00" + struct.pack("<I", ck.mask_crc(ck.crc32c(contents + b"\x00"))))
        return bytes([off, len(contents)])                      # both < 128: one-byte varints
    h_data = emit(block)
    h_meta = emit(struct.pack("<II", 0, 1))
    index_block = bytes([0, 1, len(h_data)]) + b"w" + h_data + struct.pack("<II", 0, 1)
    h_index = emit(index_block)
    footer = h_meta + h_index
    out.extend(footer + b"\x00" * (40 - len(footer)) + bytes.fromhex("57fb808b247547db"))
    return bytes(out), payload


def test_reads_hand_assembled_bundle(tmp_path):
    index, payload = _hand_built_index()
    prefix = str(tmp_path / "m.ckpt")
    open(prefix + ".index", "wb").write(index)
    open(prefix + ".data-00000-of-00001", "wb").write(payload)
    assert ck.list_variables(prefix) == [("w", 1, (2,))]
    out = ck.read_bundle(prefix)
    assert list(out) == ["w"] and out["w"].dtype == np.float32 and out["w"].tolist() == [1.5, -2.0]


def test_round_trip_and_multi_block_index(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {"vgg_16/conv1/conv1_1/weights": rng.standard_normal((3, 3, 3, 64)).astype(np.float32),
               "vgg_16/conv1/conv1_1/biases": rng.standard_normal(64).astype(np.float32),
               "global_step": np.int64(70000), "empty": np.zeros((0, 4), np.float32),
               "flags": np.array([True, False]), "half": rng.standard_normal(5).astype(np.float16)}
    for i in range(400):                                 # > one 4 KiB data block of index entries, shared key prefixes
        tensors["resnet_v1_101/block3/unit_%d/bottleneck_v1/conv2/weights" % i] = rng.standard_normal((2, i % 5 + 1)).astype(np.float32)
    prefix = str(tmp_path / "res.ckpt")
    ck.write_bundle(prefix, tensors)
    out = ck.read_bundle(prefix)
    assert sorted(out) == sorted(tensors)
    for k, v in tensors.items():
        v = np.asarray(v)
        assert out[k].dtype == v.dtype and out[k].shape == v.shape and np.array_equal(out[k], v), k
    only = ck.read_bundle(prefix, names=["global_step"])
    assert list(only) == ["global_step"] and int(only["global_step"]) == 70000


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "c.ckpt")
    ck.write_bundle(prefix, {"a": np.arange(100000, dtype=np.float32), "b": np.ones(3, np.float32)})
    data_path = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data_path, "rb").read())
    raw[12345] ^= 0x40
    open(data_path, "wb").write(bytes(raw))
    with pytest.raises(ck.CheckpointError, match="tensor checksum"):
        ck.read_bundle(prefix)
    assert ck.read_bundle(prefix, verify=False)["b"].tolist() == [1, 1, 1]
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[5] ^= 1
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ck.CheckpointError, match="block checksum"):
        ck.read_bundle(prefix)
    open(prefix + ".index", "wb").write(bytes(idx[:-1]) + b"\x00")
    with pytest.raises(ck.CheckpointError, match="magic"):
        ck.read_bundle(prefix)
    os.remove(data_path)
    open(prefix + ".index", "wb").write(_hand_built_index()[0])
    with pytest.raises(ck.CheckpointError, match="missing data shard"):
        ck.read_bundle(prefix)


def test_snappy_block():
    # literal "abcd" + copy(offset 4, len 8) + long literal (61 bytes, 1 extra length byte)
    tail = bytes(range(61))
    src = ck._put_varint(12 + 61) + bytes([3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) + bytes([60 << 2, 60]) + tail
    assert ck._snappy_uncompress(src) == b"abcdabcdabcd" + tail


def test_saver_restore_reads_bundle(tmp_path):
    """The shim's Saver.restore and the tools' --model both resolve <prefix> to the bundle; variables reach the network."""
    tensors = synth.make("res50", 5, 9)
    prefix = str(tmp_path / "res50_faster_rcnn_iter_1.ckpt")
    ck.write_bundle(prefix, tensors)
    got = ck.load_variables(prefix)
    assert set(got) == set(tensors) and all(np.array_equal(got[k], tensors[k]) for k in tensors)
    np.savez(str(tmp_path / "alt.ckpt.npz"), w=np.ones(2, np.float32))
    assert list(ck.load_variables(str(tmp_path / "alt.ckpt"))) == ["w"]
    with pytest.raises(IOError):
        ck.load_variables(str(tmp_path / "nothing.ckpt"))

    from tf_faster_rcnn_b200 import paths
    paths.add_lib_path(with_shims=True)
    import tensorflow as tf
    from nets import network
    from nets.resnet_v1 import resnetv1
    before = list(network._REGISTRY)
    net = resnetv1(num_layers=50)
    net.create_architecture("TEST", 5, tag="default", anchor_scales=[8, 16, 32])
    try:
        tf.train.Saver().restore(tf.Session(), prefix)
        assert net.weights is not None
    finally:
        network._REGISTRY[:] = before


def test_command_line_list_and_conversions(tmp_path, capsys):
    prefix = str(tmp_path / "a.ckpt")
    ck.write_bundle(prefix, {"x/w": np.arange(6, dtype=np.float32).reshape(2, 3), "step": np.int64(3)})
    ck._main(["list", prefix])
    text = capsys.readouterr().out
    assert "x/w" in text and "float32" in text and "[2, 3]" in text and "int64" in text
    ck._main(["to-npz", prefix, str(tmp_path / "a.npz")])
    ck._main(["to-bundle", str(tmp_path / "a.npz"), str(tmp_path / "b.ckpt")])
    assert open(prefix + ".index", "rb").read() == open(str(tmp_path / "b.ckpt.index"), "rb").read()


def test_spec_matches_generated_variables_and_restore_is_strict(tmp_path):
    for net, C, A in (("vgg16", 21, 9), ("res50", 5, 9), ("mobile", 81, 12)):
        w = synth.make(net, C, A)
        sp = synth.spec(net, C, A)
        assert list(w) == list(sp) and all(w[k].shape == tuple(sp[k]) for k in w)
        assert synth.check(net, w, C, A) == []
        assert synth.check(net, dict(w, **{"global_step": np.int64(1), "x/Momentum": np.zeros(3)}), C, A) == []   # extras ignored
    assert synth.spec("mobile", 21, 9, depth_multiplier=0.5)["MobilenetV1/Conv2d_0/weights"] == (3, 3, 3, 16)
    assert synth.spec("res50", 21, 9, rpn_channels=256)["resnet_v1_50/rpn_conv/3x3/weights"] == (3, 3, 1024, 256)

    from tf_faster_rcnn_b200 import paths
    paths.add_lib_path(with_shims=True)
    import tensorflow as tf
    from nets import network
    from nets.resnet_v1 import resnetv1
    before = list(network._REGISTRY)
    try:
        net = resnetv1(num_layers=50)
        net.create_architecture("TEST", 21, tag="default", anchor_scales=[8, 16, 32])
        assert net.arch_name() == "res50"
        w = synth.make("res50", 5, 9)                     # a 5-class checkpoint into a 21-class graph
        prefix = str(tmp_path / "five.ckpt")
        ck.write_bundle(prefix, w)
        with pytest.raises(ValueError, match="cls_score/weights: checkpoint has shape"):
            tf.train.Saver().restore(tf.Session(), prefix)
        del w["resnet_v1_50/block2/unit_1/bottleneck_v1/conv2/BatchNorm/moving_variance"]
        msgs = net.check_variables(w)
        assert any("moving_variance not found in checkpoint" in m for m in msgs)
        with pytest.raises(ValueError, match="not found in checkpoint"):
            net.load_weights(w, strict=True)
        net.load_weights(w)                               # the non-strict path (synthetic weights, tests) is unchanged
    finally:
        network._REGISTRY[:] = before
