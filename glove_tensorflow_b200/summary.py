"""TensorBoard event files without TensorFlow: the summaries the reference's model_fn and tf.estimator write.

The reference's ``add_summary`` [ref src/models/model_utils.py:113-118] registers, under the name scope ``mf``, one scalar
(``mf/global_bias``) and two histograms (``mf/row_biases``, ``mf/col_biases``); tf.estimator adds ``loss`` and
``global_step/sec`` every ``save_summary_steps`` (100) TRAIN steps to ``<job_dir>/events.out.tfevents.*`` and the EVAL
metrics of the head [ref src/models/estimator.py:87-92] to ``<job_dir>/eval/``.  This module writes the same tags in the
same on-disk format (TFRecord framing with masked CRC-32C, ``Event`` / ``Summary`` / ``HistogramProto`` protobufs encoded
by hand), so ``tensorboard --logdir <job_dir>`` shows the same dashboards.  ``read_events`` is the inverse, used by the
tests (and handy for checking a reference run's files against ours)."""
import os
import socket
import struct
import time

import numpy as np

# ---- CRC-32C (Castagnoli), table-driven ----------------------------------------------------------------------------
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- protobuf wire format (just what Event / Summary / HistogramProto need) ---------------------------------------------
def _varint(n: int) -> bytes:
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _key(field: int, wire: int) -> bytes:
    return _varint((field << 3) | wire)


def _f_double(field, v): return _key(field, 1) + struct.pack("<d", float(v))
def _f_float(field, v): return _key(field, 5) + struct.pack("<f", float(v))
def _f_varint(field, v): return _key(field, 0) + _varint(int(v))
def _f_bytes(field, b): return _key(field, 2) + _varint(len(b)) + b
def _f_packed_doubles(field, a): return _f_bytes(field, np.asarray(a, "<f8").tobytes())


def _default_bucket_limits():
    """TensorFlow's default histogram buckets (tensorflow/core/lib/histogram/histogram.cc): +-1e-12 * 1.1^k up to 1e20."""
    pos = []
    v = 1e-12
    while v < 1e20:
        pos.append(v)
        v *= 1.1
    return np.array([-x for x in reversed(pos)] + [0.0] + pos + [np.finfo(np.float64).max])


_LIMITS = _default_bucket_limits()


def histogram_proto(values) -> bytes:
    """HistogramProto of a tensor, bucketed like tf.summary.histogram (v1): only non-empty buckets are kept, each with its
    upper limit (an empty bucket is kept where it separates two non-empty ones, as TF does)."""
    v = np.asarray(values, np.float64).reshape(-1)
    if v.size == 0:
        v = np.zeros(1)
    idx = np.searchsorted(_LIMITS, v, side="right")          # bucket i holds limits[i-1] <= x < limits[i]
    idx = np.minimum(idx, len(_LIMITS) - 1)
    counts = np.bincount(idx, minlength=len(_LIMITS)).astype(np.float64)
    nz = np.flatnonzero(counts)
    lo, hi = nz[0], nz[-1]
    keep = np.arange(lo, hi + 1)
    msg = (_f_double(1, v.min()) + _f_double(2, v.max()) + _f_double(3, v.size) + _f_double(4, v.sum())
           + _f_double(5, np.square(v).sum()) + _f_packed_doubles(6, _LIMITS[keep]) + _f_packed_doubles(7, counts[keep]))
    return msg


class EventWriter:
    """Appends TFRecord-framed Event protos to <logdir>/events.out.tfevents.<time>.<host> (what tf.summary.FileWriter does)."""

    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, "events.out.tfevents.%010d.%s" % (int(time.time()), socket.gethostname()))
        self._f = open(self.path, "ab")
        self._write(_f_double(1, time.time()) + _f_bytes(3, b"brain.Event:2"))     # file_version record

    def _write(self, event: bytes):
        head = struct.pack("<Q", len(event))
        self._f.write(head + struct.pack("<I", masked_crc(head)) + event + struct.pack("<I", masked_crc(event)))

    def add(self, step, scalars=None, histograms=None):
        values = b""
        for tag, v in (scalars or {}).items():
            values += _f_bytes(1, _f_bytes(1, tag.encode()) + _f_float(2, v))
        for tag, arr in (histograms or {}).items():
            values += _f_bytes(1, _f_bytes(1, tag.encode()) + _f_bytes(5, histogram_proto(arr)))
        self._write(_f_double(1, time.time()) + _f_varint(2, step) + _f_bytes(5, values))
        self._f.flush()

    def close(self):
        self._f.close()


# ---- reader (tests / cross-checks) -------------------------------------------------------------------------------------
def _parse(msg: bytes):
    """[(field, wire, value)] of one protobuf message (value: int, 8 / 4 raw bytes, or bytes)."""
    out, i = [], 0
    while i < len(msg):
        k = shift = 0
        while True:
            b = msg[i]; i += 1
            k |= (b & 0x7F) << shift; shift += 7
            if not b & 0x80:
                break
        field, wire = k >> 3, k & 7
        if wire == 0:
            v = shift = 0
            while True:
                b = msg[i]; i += 1
                v |= (b & 0x7F) << shift; shift += 7
                if not b & 0x80:
                    break
        elif wire == 1:
            v = msg[i:i + 8]; i += 8
        elif wire == 5:
            v = msg[i:i + 4]; i += 4
        elif wire == 2:
            n = shift = 0
            while True:
                b = msg[i]; i += 1
                n |= (b & 0x7F) << shift; shift += 7
                if not b & 0x80:
                    break
            v = msg[i:i + n]; i += n
        else:
            raise ValueError("unsupported wire type %d" % wire)
        out.append((field, wire, v))
    return out


def read_events(path):
    """[{'step': n, 'scalars': {tag: float}, 'histograms': {tag: {min, max, num, sum, sum_squares, bucket_limit, bucket}}}];
    verifies both CRCs of every record."""
    data = open(path, "rb").read()
    events, i = [], 0
    while i < len(data):
        head = data[i:i + 8]
        n = struct.unpack("<Q", head)[0]
        assert struct.unpack("<I", data[i + 8:i + 12])[0] == masked_crc(head), "length CRC mismatch"
        body = data[i + 12:i + 12 + n]
        assert struct.unpack("<I", data[i + 12 + n:i + 16 + n])[0] == masked_crc(body), "data CRC mismatch"
        i += 16 + n
        ev = {"step": 0, "scalars": {}, "histograms": {}}
        for field, wire, v in _parse(body):
            if field == 2:
                ev["step"] = v
            elif field == 3:
                ev["file_version"] = v.decode()
            elif field == 5:
                for f2, _, val in _parse(v):
                    if f2 != 1:
                        continue
                    tag, simple, histo = None, None, None
                    for f3, w3, x in _parse(val):
                        if f3 == 1:
                            tag = x.decode()
                        elif f3 == 2:
                            simple = struct.unpack("<f", x)[0]
                        elif f3 == 5:
                            h = {}
                            for f4, w4, y in _parse(x):
                                name = {1: "min", 2: "max", 3: "num", 4: "sum", 5: "sum_squares", 6: "bucket_limit", 7: "bucket"}[f4]
                                h[name] = struct.unpack("<d", y)[0] if w4 == 1 else np.frombuffer(y, "<f8")
                            histo = h
                    if simple is not None:
                        ev["scalars"][tag] = simple
                    if histo is not None:
                        ev["histograms"][tag] = histo
        events.append(ev)
    return events
