"""Input step before the path (dataset.py): the hand-written TFRecord framing / protobuf wire format are pinned against the CRC32C
check value and the `protobuf` runtime; the parse function and the von Mises-Fisher target against closed forms.  CPU only."""
import os

import numpy as np
import pytest


def test_crc32c_check_value_and_masking(pkg):
    D = pkg.dataset
    assert D.crc32c(b"123456789") == 0xE3069283                       # the standard CRC-32C check value
    assert D.crc32c(b"") == 0
    c = D.crc32c(b"abc")
    assert D.masked_crc(b"abc") == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def test_example_wire_format_matches_protobuf_runtime(pkg):
    """Build tf.train.Example's messages (same names / field numbers as tensorflow/core/example/{example,feature}.proto) with the
    protobuf runtime and compare parse results in both directions."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="example_test.proto", package="tft", syntax="proto3")
    def msg(name):
        m = fd.message_type.add(); m.name = name; return m
    F = descriptor_pb2.FieldDescriptorProto
    m = msg("BytesList"); m.field.add(name="value", number=1, type=F.TYPE_BYTES, label=F.LABEL_REPEATED)
    m = msg("FloatList"); m.field.add(name="value", number=1, type=F.TYPE_FLOAT, label=F.LABEL_REPEATED)
    m = msg("Feature")
    m.oneof_decl.add(name="kind")
    m.field.add(name="bytes_list", number=1, type=F.TYPE_MESSAGE, type_name=".tft.BytesList", label=F.LABEL_OPTIONAL, oneof_index=0)
    m.field.add(name="float_list", number=2, type=F.TYPE_MESSAGE, type_name=".tft.FloatList", label=F.LABEL_OPTIONAL, oneof_index=0)
    m = msg("Features")
    e = m.nested_type.add(name="FeatureEntry"); e.options.map_entry = True
    e.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    e.field.add(name="value", number=2, type=F.TYPE_MESSAGE, type_name=".tft.Feature", label=F.LABEL_OPTIONAL)
    m.field.add(name="feature", number=1, type=F.TYPE_MESSAGE, type_name=".tft.Features.FeatureEntry", label=F.LABEL_REPEATED)
    m = msg("Example"); m.field.add(name="features", number=1, type=F.TYPE_MESSAGE, type_name=".tft.Features", label=F.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    Example = message_factory.GetMessageClass(pool.FindMessageTypeByName("tft.Example"))
    img = np.arange(24, dtype=np.float32).tobytes()
    ours = pkg.dataset.encode_example({"image": img, "azimuth": 63.0, "elevation": 17.25})
    ex = Example()
    ex.ParseFromString(ours)                                           # the runtime reads what we wrote
    assert ex.features.feature["image"].bytes_list.value[0] == img
    assert list(ex.features.feature["azimuth"].float_list.value) == [63.0]
    assert list(ex.features.feature["elevation"].float_list.value) == [17.25]
    theirs = Example()
    theirs.features.feature["image"].bytes_list.value.append(img)
    theirs.features.feature["azimuth"].float_list.value.append(63.0)
    theirs.features.feature["elevation"].float_list.value.append(17.25)
    back = pkg.dataset.decode_example(theirs.SerializeToString())      # we read what the runtime writes
    assert back == {"image": [img], "azimuth": [63.0], "elevation": [17.25]}


def test_tfrecord_round_trip_and_corruption(pkg, tmp_path):
    D = pkg.dataset
    rng = np.random.default_rng(0)
    imgs = [(rng.uniform(0, 1, D.IMSHAPE) ** 4 * 100).astype(np.float32) for _ in range(3)]
    recs = [D.serialize_ds(im, 63.0, 10.0 + i) for i, im in enumerate(imgs)]
    for compress in (True, False):
        path = str(tmp_path / f"a_{compress}.tfrecord")
        D.write_tfrecord(path, recs, compress=compress)
        assert list(D.read_tfrecord(path, compress=compress)) == recs
    raw = bytearray(open(str(tmp_path / "a_False.tfrecord"), "rb").read())
    raw[40] ^= 0xFF
    open(str(tmp_path / "bad.tfrecord"), "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        list(D.read_tfrecord(str(tmp_path / "bad.tfrecord"), compress=False))


def test_parse_function_and_vmf(pkg, tmp_path):
    D = pkg.dataset
    rng = np.random.default_rng(1)
    H, W, _ = D.IMSHAPE
    img = (rng.uniform(0, 1, D.IMSHAPE) ** 4 * 100).astype(np.float32)
    hdr, sun = D.parse_function(D.serialize_ds(img, 63.0, 8.0))
    assert hdr.shape == D.IMSHAPE and abs(float(hdr.mean()) - 0.5) < 1e-4                     # 0.5 / mean normalisation
    assert np.allclose(hdr / hdr.mean(), img[:, :, ::-1] / img.mean(), rtol=1e-5)            # channel flip (train.py:107)
    bins = D.sunpose_bins(H, W)
    assert bins.shape == (H * W, 3) and np.allclose(np.linalg.norm(bins, axis=1), 1, atol=1e-6)
    assert bins[:, 1].min() > 0 and np.all(np.diff(bins[::W, 1]) > 0)                        # sky dome: elevation grows with the row index
    assert sun.shape == (H * W,) and abs(float(sun.sum()) - 1) < 1e-5 and sun.min() >= 0
    peak = int(sun.argmax())
    want = D.sphere2world((W * 0.5 - 1, 8.0), H, W)                                           # AZIMUTH_gt, elevation
    assert float(bins[peak] @ want) == pytest.approx(float((bins @ want).max()))
    # kappa -> concentration: the mass within 10 degrees of the sun direction dominates
    assert float(sun[(bins @ want) > np.cos(np.radians(10))].sum()) > 0.5
    # batching helper
    for i in range(5):
        D.write_tfrecord(str(tmp_path / f"{i}.tfrecord"), [D.serialize_ds(img * (i + 1), 63.0, 5.0 + i)])
    batches = list(D.configure_dataset(str(tmp_path), batch_size=2))
    assert len(batches) == 2 and batches[0][0].shape == (2, H, W, 3) and batches[0][1].shape == (2, H * W)
