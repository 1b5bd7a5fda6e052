"""TFRecord shards and tf.train.Example messages without TensorFlow.

The reference stores every dataset as TFRecord shards of tf.train.Example protos with the features
`name` (bytes), `xyz_pose` (float list), `png16` (the PNG file's bytes) and, for the NYU test set, `bbx`
(5 floats): data/icvl.py:118-128, data/nyu.py:159-176, data/msra.py:186-196, written by
data/dataset_base.py:52-66 (tf.python_io.TFRecordWriter) and read back by tf.TFRecordReader +
tf.parse_single_example (dataset_base.py:182-199, icvl.py:131-143).  TensorFlow is not installable here, so this
module restates the two public on-disk formats:

  TFRecord framing   uint64 length | uint32 masked_crc32c(length) | data | uint32 masked_crc32c(data)   (little endian)
                     masked(c) = ((c >> 15 | c << 17) + 0xa282ead8) mod 2^32, crc32c = CRC-32/Castagnoli
  Example proto      Example{1: Features{1: map<string, Feature>}}, Feature{1: BytesList{1: bytes*},
                     2: FloatList{1: packed float*}, 3: Int64List{1: packed varint*}}

Host-side byte plumbing only; no arithmetic of the hot path lives here.
"""
import struct

import numpy as np

# ---- CRC-32C (Castagnoli, reflected polynomial 0x82F63B78) -----------------------------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.arange(256, dtype=np.uint32)
        for _ in range(8):
            t = np.where(t & 1, (t >> 1) ^ np.uint32(0x82F63B78), t >> 1).astype(np.uint32)
        _CRC_TABLE = [int(x) for x in t]
    return _CRC_TABLE


def _crc_scalar(buf, c):
    tab = _crc_table()
    for b in buf:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c


_LANE = 1024          # bytes per lane of the vectorised path
_SHIFT_TABS = None    # 4 x 256 tables of the linear map "advance the CRC register over _LANE zero bytes"


def _shift_tables():
    """The CRC register update is linear over GF(2): crc(A||B) = shift_{|B|}(crc(A)) xor crc0(B).  Build shift_{_LANE} by
    squaring the one-byte step (as columns = images of the 32 unit vectors), then expand it to byte-indexed tables."""
    global _SHIFT_TABS
    if _SHIFT_TABS is None:
        tab = _crc_table()
        cols = [tab[(1 << i) & 0xFF] ^ ((1 << i) >> 8) for i in range(32)]       # one zero byte
        apply = lambda cs, v: _xor_all(cs[i] for i in range(32) if (v >> i) & 1)
        n = _LANE
        assert n & (n - 1) == 0
        while n > 1:
            cols = [apply(cols, c) for c in cols]
            n >>= 1
        _SHIFT_TABS = [[apply(cols, b << (8 * k)) for b in range(256)] for k in range(4)]
    return _SHIFT_TABS


def _xor_all(it):
    r = 0
    for v in it:
        r ^= v
    return r


def crc32c(data):
    """CRC-32C of a bytes-like object.  Long inputs are cut into _LANE-byte lanes whose CRCs advance together as one NumPy
    vector (table look-ups over all lanes per byte position) and are then chained with the linear shift map."""
    buf = memoryview(bytes(data))
    n = len(buf)
    lanes = n // _LANE
    c = 0xFFFFFFFF
    if lanes >= 8:
        tab = np.array(_crc_table(), dtype=np.uint32)
        a = np.frombuffer(buf[:lanes * _LANE], dtype=np.uint8).reshape(lanes, _LANE)
        v = np.zeros(lanes, dtype=np.uint32)
        for j in range(_LANE):
            v = tab[(v ^ a[:, j]) & 0xFF] ^ (v >> 8)
        t0, t1, t2, t3 = _shift_tables()
        for x in v.tolist():
            c = t0[c & 0xFF] ^ t1[(c >> 8) & 0xFF] ^ t2[(c >> 16) & 0xFF] ^ t3[c >> 24] ^ x
        buf = buf[lanes * _LANE:]
    return _crc_scalar(buf, c) ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- TFRecord framing --------------------------------------------------------------------------------------------------
class TFRecordError(IOError):
    pass


def read_records(path, verify="length"):
    """Yield the payload of every record of one shard.  verify: 'none' | 'length' (header crc only, cheap) | 'all'."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise TFRecordError("%s: truncated record header" % path)
            (length,), (lcrc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify != "none" and masked_crc32c(head[:8]) != lcrc:
                raise TFRecordError("%s: corrupted record length" % path)
            data = f.read(length)
            tail = f.read(4)
            if len(data) < length or len(tail) < 4:
                raise TFRecordError("%s: truncated record" % path)
            if verify == "all" and masked_crc32c(data) != struct.unpack("<I", tail)[0]:
                raise TFRecordError("%s: corrupted record data" % path)
            yield data


class TFRecordWriter:
    """tf.python_io.TFRecordWriter (data/dataset_base.py:59-63)."""

    def __init__(self, path):
        self._f = open(path, "wb")

    def write(self, data):
        head = struct.pack("<Q", len(data))
        self._f.write(head + struct.pack("<I", masked_crc32c(head)) + data + struct.pack("<I", masked_crc32c(data)))

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ---- protobuf wire format (the subset tf.train.Example uses) ------------------------------------------------------------
def _varint(buf, pos):
    r, shift = 0, 0
    while True:
        b = buf[pos]; pos += 1
        r |= (b & 0x7F) << shift
        if not b & 0x80:
            return r, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _fields(buf):
    """Yield (field_number, wire_type, value) of one message; value is int (varint/fixed) or a memoryview (length-delimited)."""
    buf = memoryview(buf)
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fn, wt, v


def _ld(fn, payload):
    return _put_varint((fn << 3) | 2) + _put_varint(len(payload)) + payload


def parse_example(data):
    """tf.parse_single_example without a feature_map: -> {name: bytes list | float32 array | int64 array}."""
    out = {}
    for fn, _, features in _fields(data):
        if fn != 1:
            continue
        for fn2, _, entry in _fields(features):                 # map<string, Feature> entries
            if fn2 != 1:
                continue
            key, feat = None, None
            for fn3, _, v in _fields(entry):
                if fn3 == 1:
                    key = bytes(v).decode("utf-8")
                elif fn3 == 2:
                    feat = v
            val = None
            for kind, _, lst in _fields(feat if feat is not None else b""):
                if kind == 1:                                   # BytesList
                    val = [bytes(v) for f4, _, v in _fields(lst) if f4 == 1]
                elif kind == 2:                                 # FloatList (packed, or repeated fixed32)
                    parts = []
                    for f4, wt, v in _fields(lst):
                        if f4 == 1:
                            parts.append(bytes(v))
                    val = np.frombuffer(b"".join(parts), dtype="<f4").copy()
                elif kind == 3:                                 # Int64List
                    ints = []
                    for f4, wt, v in _fields(lst):
                        if f4 != 1:
                            continue
                        if wt == 0:
                            ints.append(v)
                        else:
                            p, mv = 0, v
                            while p < len(mv):
                                x, p = _varint(mv, p)
                                ints.append(x)
                    val = np.array([x - (1 << 64) if x >= (1 << 63) else x for x in ints], dtype=np.int64)
            if key is not None:
                out[key] = val
    return out


def make_example(features):
    """Inverse of parse_example: {name: bytes | [bytes] | float sequence} -> serialized tf.train.Example
    (_bytes_feature / _float_feature of data/dataset_base.py:18-26)."""
    entries = b""
    for key in sorted(features):
        v = features[key]
        if isinstance(v, (bytes, bytearray)):
            v = [bytes(v)]
        if isinstance(v, list) and v and isinstance(v[0], (bytes, bytearray)):
            feat = _ld(1, b"".join(_ld(1, bytes(x)) for x in v))
        else:
            feat = _ld(2, _ld(1, np.asarray(v, dtype="<f4").tobytes()))
        entries += _ld(1, _ld(1, key.encode("utf-8")) + _ld(2, feat))
    return _ld(1, entries)
