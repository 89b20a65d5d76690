"""Input step before the path, host side: the GZIP TFRecord files the reference trains from
(DataGeneration/makeTFRecord.py:24-31, 45-63: one tf.train.Example per file with `image` = raw float32 HDR bytes, `azimuth`,
`elevation`) and train._parse_function (train.py:96-117): decode, BGR->RGB flip, 0.5 / mean normalisation, and the von Mises-Fisher
sun-position target (train.py:42-52 over tf_utils.sunpose_init / sphere2world, tf_utils.py:95-129).

No TensorFlow: the TFRecord framing (length, masked CRC32C) and the three-message protobuf wire format are written out by hand —
tests/test_dataset.py pins them against the CRC32C check value and against the `protobuf` runtime building the same messages.
Everything here is numpy on the host; the arrays it yields are what `train.Step._preprocessing` / `train_sun.SunTrainer` take."""
from __future__ import annotations

import glob
import gzip
import os
import struct

import numpy as np

IMSHAPE = (32, 128, 3)                      # train.py:31
PI = np.float32(np.pi)


# ---- CRC32C (Castagnoli), masked as TFRecord does ----------------------------------------------------------------------------------
def _crc_table():
    tab = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_TAB = _crc_table()


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _TAB[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- protobuf wire format of tf.train.Example{features{feature{key -> Feature{bytes_list | float_list}}}} --------------------------
def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _read_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _ld(field: int, payload: bytes) -> bytes:           # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_example(features: dict) -> bytes:
    """features: name -> bytes (BytesList with one value) or float / sequence of floats (packed FloatList)."""
    entries = b""
    for key in sorted(features):                         # protobuf map entries; any order parses, sorted is deterministic
        val = features[key]
        if isinstance(val, (bytes, bytearray)):
            feature = _ld(1, _ld(1, bytes(val)))                                     # Feature.bytes_list = 1, BytesList.value = 1
        else:
            packed = np.asarray(val, "<f4").reshape(-1).tobytes()
            feature = _ld(2, _ld(1, packed))                                         # Feature.float_list = 2, FloatList.value = 1 (packed)
        entries += _ld(1, _ld(1, key.encode()) + _ld(2, feature))                    # Features.feature = 1 (map entry: key = 1, value = 2)
    return _ld(1, entries)                                                           # Example.features = 1


def _fields(buf):
    pos = 0
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 2:
            n, pos = _read_varint(buf, pos)
            yield field, wire, buf[pos:pos + n]
            pos += n
        elif wire == 5:
            yield field, wire, buf[pos:pos + 4]
            pos += 4
        elif wire == 0:
            v, pos = _read_varint(buf, pos)
            yield field, wire, v
        elif wire == 1:
            yield field, wire, buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError(f"unsupported protobuf wire type {wire}")


def decode_example(buf: bytes) -> dict:
    out = {}
    for f1, _, features in _fields(buf):
        if f1 != 1:
            continue
        for f2, _, entry in _fields(features):
            if f2 != 1:
                continue
            key, feature = None, b""
            for f3, _, v in _fields(entry):
                if f3 == 1:
                    key = bytes(v).decode()
                elif f3 == 2:
                    feature = v
            for kind, _, lst in _fields(feature):
                if kind == 1:                                                        # BytesList
                    out[key] = [bytes(v) for f, _, v in _fields(lst) if f == 1]
                elif kind == 2:                                                      # FloatList: packed or one fixed32 per element
                    vals = []
                    for f, wire, v in _fields(lst):
                        if f == 1:
                            vals.extend(np.frombuffer(bytes(v), "<f4").tolist())
                    out[key] = vals
    return out


# ---- TFRecord framing, GZIP-compressed files (TFRecordOptions(compression_type="GZIP"), makeTFRecord.py:11) -------------------------
def write_tfrecord(path, records, compress=True):
    opener = gzip.open if compress else open
    with opener(path, "wb") as f:
        for rec in records:
            head = struct.pack("<Q", len(rec))
            f.write(head + struct.pack("<I", masked_crc(head)) + rec + struct.pack("<I", masked_crc(rec)))


def read_tfrecord(path, compress=True, verify=True):
    opener = gzip.open if compress else open
    with opener(path, "rb") as f:
        while True:
            head = f.read(8)
            if not head:
                return
            if len(head) != 8:
                raise ValueError("truncated TFRecord length")
            (n,) = struct.unpack("<Q", head)
            (hcrc,) = struct.unpack("<I", f.read(4))
            rec = f.read(n)
            (dcrc,) = struct.unpack("<I", f.read(4))
            if verify and (hcrc != masked_crc(head) or len(rec) != n or dcrc != masked_crc(rec)):
                raise ValueError("corrupted TFRecord (CRC mismatch)")
            yield rec


def serialize_ds(image: np.ndarray, azimuth: float, elevation: float) -> bytes:
    """DataGeneration/makeTFRecord.py:24-31 (image.tostring() of the float32 HDR panorama, :45)."""
    return encode_example({"image": np.asarray(image, "<f4").tobytes(), "azimuth": float(azimuth), "elevation": float(elevation)})


# ---- sun-position target (train.py:40-52, tf_utils.py:95-129) -------------------------------------------------------------------------
def sunpose_bins(h, w):
    """tf_utils.sunpose_init(i, h, w) for i in range(h*w) -> [h*w, 3] unit vectors (fp32 ops in the reference's order)."""
    i = np.arange(h * w, dtype=np.float32)
    row = np.floor(i / np.float32(w))
    x = ((i + np.float32(1.0)) - row * np.float32(w) - np.float32(1.0)) * np.float32(360.0 / w) + np.float32(360.0 / (w * 2.0))
    y = row * np.float32(90.0 / h) + np.float32(90.0 / (2.0 * h))
    phi = y * np.float32(PI / np.float32(180.0))
    theta = (x - np.float32(180.0)) * np.float32(PI / np.float32(180.0))
    return np.stack([np.cos(phi) * np.cos(theta), np.sin(phi), np.cos(phi) * np.sin(theta)], axis=1).astype(np.float32)


def sphere2world(sunpose, h, w, skydome=True):
    """tf_utils.sphere2world (:95-110)."""
    x, y = (np.float32(v) for v in sunpose)
    unit_w = np.float32(2) * PI / np.float32(w)
    unit_h = PI / np.float32(h * 2 if skydome else h)
    theta = (x - np.float32(0.5 * w)) * unit_w
    phi = (np.float32(h) - y) * unit_h if skydome else (np.float32(h * 0.5) - y) * unit_h
    return np.array([np.cos(phi) * np.cos(theta), np.sin(phi), np.cos(phi) * np.sin(theta)], np.float32)


def vMF(x, y, h, w, kappa=80.0, bins=None):
    """train.vMF (train.py:42-52): exp(kappa * <bin, sun direction>) normalised over the h*w sky bins."""
    bins = sunpose_bins(h, w) if bins is None else bins
    dot = bins @ sphere2world((x, y), h, w, skydome=True)
    pdf = np.exp(np.float32(kappa) * dot.astype(np.float32)).astype(np.float32)
    return pdf / pdf.sum(dtype=np.float32)


def parse_function(record: bytes, imshape=IMSHAPE, bins=None):
    """train._parse_function (train.py:96-117) -> (hdr [H, W, 3] float32, sun_pose [H*W] float32)."""
    ex = decode_example(record)
    hdr = np.frombuffer(ex["image"][0], "<f4").reshape(imshape)[:, :, ::-1]          # :105-107
    hdr = (np.float32(0.5) * hdr / (hdr.mean(dtype=np.float32) + np.float32(1e-6))).astype(np.float32)   # :109-110
    azimuth = imshape[1] * 0.5 - 1                                                   # AZIMUTH_gt, train.py:32
    elevation = ex["elevation"][0]
    return hdr, vMF(azimuth, elevation, imshape[0], imshape[1], bins=bins)           # :112-115


def configure_dataset(dirpath, batch_size=32, imshape=IMSHAPE, shuffle_seed=None, drop_remainder=True):
    """train.configureDataset (:119-133): every *.tfrecord under dirpath, parsed, optionally shuffled, batched."""
    files = sorted(glob.glob(os.path.join(dirpath, "*.tfrecord")))
    bins = sunpose_bins(imshape[0], imshape[1])
    items = [parse_function(rec, imshape, bins) for path in files for rec in read_tfrecord(path)]
    if shuffle_seed is not None:
        np.random.default_rng(shuffle_seed).shuffle(items)
    for lo in range(0, len(items), batch_size):
        chunk = items[lo:lo + batch_size]
        if len(chunk) < batch_size and drop_remainder:
            return
        yield np.stack([c[0] for c in chunk]), np.stack([c[1] for c in chunk])
