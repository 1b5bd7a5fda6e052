"""PNG decode/encode for the two pixel formats the reference's datasets use, without TensorFlow or OpenCV.

  tf.image.decode_png(png16, channels=1, dtype=tf.uint16)   data/icvl.py:138, data/msra.py:205   -> 16-bit greyscale
  tf.image.decode_png(img, channels=3, dtype=tf.uint8)      data/nyu.py:148-156                  -> 8-bit RGB (depth = G*256 + B)
  cv2.imwrite(path, dm.astype('uint16'))                    data/msra.py:146                     -> encode_png

Supported: colour types 0 (grey), 2 (RGB), 4 (grey+alpha), 6 (RGBA), bit depths 8 and 16, non-interlaced, all five
row filters.  Host-side byte plumbing (zlib + NumPy); no arithmetic of the hot path lives here.
"""
import struct
import zlib

import numpy as np

_SIG = b"\x89PNG\r\n\x1a\n"
_CHANNELS = {0: 1, 2: 3, 4: 2, 6: 4}


class PngError(ValueError):
    pass


def _chunks(data):
    pos = 8
    while pos + 8 <= len(data):
        (ln,), typ = struct.unpack(">I", data[pos:pos + 4]), data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + ln]
        if len(body) < ln:
            raise PngError("truncated %s chunk" % typ.decode("latin1"))
        yield typ, body
        pos += 12 + ln


def _unfilter(raw, h, stride, bpp):
    """Undo the per-row filters.  Sub/Up are vectorised (uint8 cumsum wraps mod 256); Average and Paeth carry a
    dependency on the pixel to the left and run as a scalar loop over the row."""
    out = np.zeros((h + 1, stride), dtype=np.uint8)              # row 0 = the all-zero "previous row" of the first row
    rows = np.frombuffer(raw, dtype=np.uint8, count=h * (stride + 1)).reshape(h, stride + 1)
    for y in range(h):
        ft, line, prev = int(rows[y, 0]), rows[y, 1:], out[y]
        cur = out[y + 1]
        if ft == 0:
            cur[:] = line
        elif ft == 1:
            for c in range(bpp):
                cur[c::bpp] = np.cumsum(line[c::bpp], dtype=np.uint8)
        elif ft == 2:
            cur[:] = line + prev
        elif ft == 3:
            ln, pv, cu = line.tolist(), prev.tolist(), [0] * stride
            for i in range(stride):
                left = cu[i - bpp] if i >= bpp else 0
                cu[i] = (ln[i] + ((left + pv[i]) >> 1)) & 0xFF
            cur[:] = cu
        elif ft == 4:
            ln, pv, cu = line.tolist(), prev.tolist(), [0] * stride
            for i in range(stride):
                a = cu[i - bpp] if i >= bpp else 0
                b = pv[i]
                c = pv[i - bpp] if i >= bpp else 0
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cu[i] = (ln[i] + pred) & 0xFF
            cur[:] = cu
        else:
            raise PngError("unknown row filter %d" % ft)
    return out[1:]


def decode_png(data):
    """PNG bytes -> (H,W) or (H,W,C) array, uint8 or uint16 (native byte order)."""
    data = bytes(data)
    if data[:8] != _SIG:
        raise PngError("not a PNG stream")
    ihdr, idat = None, []
    for typ, body in _chunks(data):
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
    if ihdr is None or not idat:
        raise PngError("missing IHDR/IDAT")
    w, h, depth, ctype, _, _, interlace = ihdr
    if interlace or depth not in (8, 16) or ctype not in _CHANNELS:
        raise PngError("unsupported PNG (depth %d, colour type %d, interlace %d)" % (depth, ctype, interlace))
    ch = _CHANNELS[ctype]
    bpp = ch * depth // 8
    stride = w * bpp
    raw = zlib.decompress(b"".join(idat))
    if len(raw) < h * (stride + 1):
        raise PngError("short pixel data")
    px = _unfilter(raw, h, stride, bpp)
    if depth == 16:
        img = px.reshape(h, w * ch, 2).astype(np.uint16)
        img = (img[..., 0] << 8) | img[..., 1]                   # big-endian samples
    else:
        img = px
    img = img.reshape(h, w, ch)
    return np.ascontiguousarray(img[..., 0] if ch == 1 else img)


def _chunk(typ, body):
    return struct.pack(">I", len(body)) + typ + body + struct.pack(">I", zlib.crc32(typ + body) & 0xFFFFFFFF)


def encode_png(img, filter_type=0, level=6):
    """(H,W) uint16/uint8 grey or (H,W,3) uint8 RGB -> PNG bytes.  filter_type 0..4 is applied to every row (tests use the
    non-trivial ones to exercise the decoder)."""
    img = np.asarray(img)
    if img.dtype not in (np.uint8, np.uint16):
        raise PngError("encode_png takes uint8 or uint16")
    ch = 1 if img.ndim == 2 else img.shape[2]
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[ch]
    h, w = img.shape[:2]
    depth = 16 if img.dtype == np.uint16 else 8
    px = img.astype(">u2").view(np.uint8) if depth == 16 else img
    px = np.ascontiguousarray(px).reshape(h, -1).astype(np.int32)
    bpp = ch * depth // 8
    stride = px.shape[1]
    prev = np.zeros(stride, np.int32)
    lines = []
    for y in range(h):
        cur = px[y]
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        ul = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if filter_type == 0:
            f = cur
        elif filter_type == 1:
            f = cur - left
        elif filter_type == 2:
            f = cur - prev
        elif filter_type == 3:
            f = cur - ((left + prev) >> 1)
        elif filter_type == 4:
            p = left + prev - ul
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
            f = cur - pred
        else:
            raise PngError("filter_type must be 0..4")
        lines.append(bytes([filter_type]) + (f & 0xFF).astype(np.uint8).tobytes())
        prev = cur
    ihdr = struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)
    return _SIG + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zlib.compress(b"".join(lines), level)) + _chunk(b"IEND", b"")
