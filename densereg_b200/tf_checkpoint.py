"""TensorFlow V2 checkpoint ("tensor bundle") reader / writer and the um_v1 variable-name map, without TensorFlow
(SURVEY.md 8f-4).

The reference saves and restores with `tf.train.Saver(tf.global_variables())` (model/train_single_gpu.py:113,172-175,
model/test_model.py:31-35): `model.ckpt-<step>.index` + `model.ckpt-<step>.data-00000-of-00001`, the format the authors'
pretrained models (exp/scripts/fetch_*_model.sh) ship in.  This module restates the public on-disk format
(tensorflow/core/util/tensor_bundle + tensorflow/core/lib/io/table = the LevelDB table format):

  .index   sorted string table: data blocks of prefix-compressed (key, value) entries + restart array, each block followed by
           a 5-byte trailer {compression type (0 raw, 1 snappy), masked crc32c}; index block; 48-byte footer with the
           metaindex/index block handles and the magic 0xdb4775248b80fb57.  key "" -> BundleHeaderProto, key <variable name>
           -> BundleEntryProto {1 dtype, 2 shape, 3 shard_id, 4 offset, 5 size, 6 crc32c (fixed32, masked)}.
  .data-*  raw little-endian tensor bytes at [offset, offset+size).

and maps variable names to the flat parameter / BRN-state / Adam buffers of the engine:

  network/slim/ops.py:266,270-295   variable_scope(scope,'Conv') -> 'Conv', 'Conv_1', ... (uniquified per enclosing scope;
                                    the stem lives under 'hg_imgproc', network/um_v1.py:84), 'weights', 'biases'
  network/slim/ops.py:81-128        '<conv>/BatchReNorm/{beta,gamma,moving_mean,moving_variance,r_max,d_max,curr_t}'
  network/slim/ops.py:134-138       assign_moving_average(zero_debias=True, TF 1.3) -> '.../moving_mean/biased', '.../moving_mean/local_step'
                                    created under variable_scope(<full op name>) INSIDE the BatchReNorm scope, which prefixes the scope twice
  model/hourglass_um_crop_tiny.py:436-439  AdamOptimizer slots '<var>/Adam', '<var>/Adam_1', 'beta1_power', 'beta2_power'
  model/train_single_gpu.py:41-43   'global_step'

PARITY UNPINNED for the name map: no TF-written checkpoint exists on the box; the container format is pinned by its
published constants (magic, masked CRC-32C, varint coding) and by a writer/reader round trip, the snappy path against
pyarrow's snappy codec.  Host-side byte plumbing only.
"""
import os
import re
import struct

import numpy as np

from .tfrecord import _fields, _ld, _put_varint, _varint, crc32c

TABLE_MAGIC = 0xDB4775248B80FB57
DT_FLOAT, DT_INT32, DT_INT64 = 1, 3, 9
_DTYPES = {DT_FLOAT: "<f4", DT_INT32: "<i4", DT_INT64: "<i8", 2: "<f8"}
_DT_OF = {np.dtype("float32"): DT_FLOAT, np.dtype("int32"): DT_INT32, np.dtype("int64"): DT_INT64, np.dtype("float64"): 2}


class CheckpointError(IOError):
    pass


def _mask(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- snappy (raw block format) -------------------------------------------------------------------------------------------
def snappy_decompress(buf):
    buf = bytes(buf)
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]; pos += 1
        kind = tag & 3
        if kind == 0:                                            # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little"); pos += nb
            ln += 1
            out += buf[pos:pos + ln]; pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]; pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8); pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little"); pos += 4
        if off == 0 or off > len(out):
            raise CheckpointError("corrupt snappy stream")
        for _ in range(ln):                                      # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise CheckpointError("snappy length mismatch")
    return bytes(out)


# ---- table (LevelDB sstable) ---------------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify=True):
    body, trailer = data[offset:offset + size], data[offset + size:offset + size + 5]
    if len(body) < size or len(trailer) < 5:
        raise CheckpointError("truncated table block")
    if verify and _mask(crc32c(body + trailer[:1])) != struct.unpack("<I", trailer[1:])[0]:
        raise CheckpointError("table block checksum mismatch")
    if trailer[0] == 1:
        body = snappy_decompress(body)
    elif trailer[0] != 0:
        raise CheckpointError("unknown block compression %d" % trailer[0])
    return body


def _block_entries(block):
    (num_restarts,) = struct.unpack("<I", block[-4:])
    end = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + unshared]; pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path, verify=True):
    """-> ordered list of (key bytes, value bytes) of one .index file."""
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise CheckpointError("%s is not a TensorFlow checkpoint index (bad magic)" % path)
    footer = data[-48:]
    _, p = _varint(footer, 0); _, p = _varint(footer, p)         # metaindex handle (unused)
    ioff, p = _varint(footer, p); isz, p = _varint(footer, p)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isz, verify)):
        boff, q = _varint(handle, 0); bsz, q = _varint(handle, q)
        out.extend(_block_entries(_read_block(data, boff, bsz, verify)))
    return out


def _build_block(entries, restart_interval=16):
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(prev), len(k))
            while shared < m and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    out += b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path, entries, block_size=4096):
    """entries: iterable of (key bytes, value bytes), keys strictly increasing (bytewise)."""
    entries = list(entries)
    for a, b in zip(entries, entries[1:]):
        if not a[0] < b[0]:
            raise CheckpointError("table keys must be strictly increasing")
    out = bytearray()

    def emit(block):
        off = len(out)
        out.extend(block + b"\x00" + struct.pack("<I", _mask(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    index, cur, cur_bytes = [], [], 0
    for k, v in entries:
        cur.append((k, v)); cur_bytes += len(k) + len(v) + 3
        if cur_bytes >= block_size:
            index.append((cur[-1][0], emit(_build_block(cur)))); cur, cur_bytes = [], 0
    if cur or not index:
        index.append((cur[-1][0] if cur else b"", emit(_build_block(cur))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    footer = meta + idx
    out.extend(footer + bytes(40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    with open(path, "wb") as f:
        f.write(out)


# ---- bundle ------------------------------------------------------------------------------------------------------------------
def _parse_shape(buf):
    dims = []
    for fn, _, dim in _fields(buf):
        if fn == 2:
            size = 0
            for f2, _, v in _fields(dim):
                if f2 == 1:
                    size = v
            dims.append(size)
    return tuple(dims)


def _parse_entry(buf):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for fn, wt, v in _fields(buf):
        if fn == 1: e["dtype"] = v
        elif fn == 2: e["shape"] = _parse_shape(v)
        elif fn == 3: e["shard_id"] = v
        elif fn == 4: e["offset"] = v
        elif fn == 5: e["size"] = v
        elif fn == 6: e["crc32c"] = struct.unpack("<I", v)[0]
        elif fn == 7: e["sliced"] = True
    return e


def read_bundle(prefix, verify=True, names=None):
    """-> {variable name: ndarray} of `prefix`.index / `prefix`.data-*-of-*.  `names`: optional subset filter."""
    entries = read_table(prefix + ".index", verify)
    num_shards = 1
    if entries and entries[0][0] == b"":
        for fn, _, v in _fields(entries[0][1]):
            if fn == 1: num_shards = v
            elif fn == 2 and v != 0:
                raise CheckpointError("big-endian checkpoints are not supported")
    shards = {}
    out = {}
    for key, val in entries:
        if key == b"":
            continue
        name = key.decode("utf-8")
        if names is not None and name not in names:
            continue
        e = _parse_entry(val)
        if e["sliced"]:
            raise CheckpointError("partitioned variable %s is not supported" % name)
        if e["dtype"] not in _DTYPES:
            continue                                             # strings / resources: nothing of the model lives there
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), "rb")
        f = shards[sid]; f.seek(e["offset"]); raw = f.read(e["size"])
        if len(raw) != e["size"]:
            raise CheckpointError("tensor %s is truncated" % name)
        if verify and e["crc32c"] is not None and _mask(crc32c(raw)) != e["crc32c"]:
            raise CheckpointError("tensor %s fails its checksum" % name)
        arr = np.frombuffer(raw, dtype=_DTYPES[e["dtype"]])
        if int(np.prod(e["shape"], dtype=np.int64)) != arr.size:
            raise CheckpointError("tensor %s: shape %s does not match %d bytes" % (name, e["shape"], e["size"]))
        out[name] = arr.reshape(e["shape"]).copy()
    for f in shards.values():
        f.close()
    return out


def write_bundle(prefix, tensors, checksum=True):
    """{name: ndarray (float32 / int32 / int64 / float64)} -> `prefix`.index + `prefix`.data-00000-of-00001.
    checksum=False writes crc32c=0 entries skipped by readers only when they do not verify -- keep True outside tests."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    header = b"\x08\x01" + b"\x10\x00" + _ld(3, b"\x08\x01")      # num_shards=1, LITTLE endian, VersionDef{producer=1}
    entries = [(b"", header)]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            a = np.asarray(tensors[name])
            if a.dtype not in _DT_OF:
                raise CheckpointError("unsupported dtype %s for %s" % (a.dtype, name))
            raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
            shape = b"".join(_ld(2, b"\x08" + _put_varint(int(d))) for d in a.shape)
            ent = b"\x08" + _put_varint(_DT_OF[a.dtype]) + _ld(2, shape)
            if offset:
                ent += b"\x20" + _put_varint(offset)
            ent += b"\x28" + _put_varint(len(raw))
            ent += b"\x35" + struct.pack("<I", _mask(crc32c(raw)) if checksum else 0)
            entries.append((name.encode("utf-8"), ent))
            f.write(raw); offset += len(raw)
    write_table(prefix + ".index", entries)


# ---- um_v1 variable names --------------------------------------------------------------------------------------------------------
def tf_scopes(layers):
    """Engine layer table (creation order, names 'stem/...' | 's<k>/...') -> TF conv scope per layer:
    'hg_imgproc/Conv', 'hg_imgproc/Conv_1', ... for the stem, then 'Conv', 'Conv_1', ... at the root scope."""
    out, n_stem, n_root = [], 0, 0
    for L in layers:
        if L["name"].startswith("stem/"):
            out.append("hg_imgproc/Conv" + ("_%d" % n_stem if n_stem else "")); n_stem += 1
        else:
            out.append("Conv" + ("_%d" % n_root if n_root else "")); n_root += 1
    return out


def variable_map(layers):
    """-> list of (tf name, buffer 'params'|'state', flat offset, shape) for every model variable of the checkpoint.
    Flat layouts: params = weights HWIO | beta,gamma or biases; state per BRN conv = moving_mean[C],
    moving_variance[C], biased_mean[C], biased_variance[C], r_max, d_max, curr_t, local_step."""
    m = []
    for L, sc in zip(layers, tf_scopes(layers)):
        k, cin, cout = L["k"], L["cin"], L["cout"]
        m.append((sc + "/weights", "params", L["w_off"], (k, k, cin, cout)))
        if L["brn"]:
            b = sc + "/BatchReNorm"
            m.append((b + "/beta", "params", L["p_off"], (cout,)))
            m.append((b + "/gamma", "params", L["p_off"] + cout, (cout,)))
            s = L["s_off"]
            m.append((b + "/moving_mean", "state", s, (cout,)))
            m.append((b + "/moving_variance", "state", s + cout, (cout,)))
            m.append((b + "/" + b + "/moving_mean/biased", "state", s + 2 * cout, (cout,)))
            m.append((b + "/" + b + "/moving_variance/biased", "state", s + 3 * cout, (cout,)))
            m.append((b + "/r_max", "state", s + 4 * cout, (1,)))
            m.append((b + "/d_max", "state", s + 4 * cout + 1, (1,)))
            m.append((b + "/curr_t", "state", s + 4 * cout + 2, (1,)))
            m.append((b + "/" + b + "/moving_mean/local_step", "state", s + 4 * cout + 3, ()))
            m.append((b + "/" + b + "/moving_variance/local_step", "state", s + 4 * cout + 3, ()))   # same counter, kept once
        else:
            m.append((sc + "/biases", "params", L["p_off"], (cout,)))
    return m


_OPTIONAL_SUFFIXES = ("/biased", "/local_step", "/r_max", "/d_max", "/curr_t")


def _lookup(tensors, name):
    """Exact name, else the same variable without the doubled zero-debias scope ('<conv>/BatchReNorm/moving_mean/biased'), else either
    form under stock slim's scope name 'BatchNorm' (the reference's fork opens 'BatchReNorm', network/slim/ops.py:81)."""
    cands = [name]
    parts = name.split("/BatchReNorm/")
    if len(parts) == 3:
        cands.append(parts[0] + "/BatchReNorm/" + parts[2])
    cands += [c.replace("BatchReNorm", "BatchNorm") for c in list(cands)]
    for c in cands:
        if c in tensors:
            return tensors[c]
    return None


def _scope_index(scope):
    """'hg_imgproc/Conv_12' -> ('hg_imgproc', 12); 'Conv' -> ('', 0): creation order of slim's auto-numbered conv scopes."""
    head, _, leaf = scope.rpartition("/")
    m = re.match(r"^(.*?)(?:_(\d+))?$", leaf)
    return head, int(m.group(2) or 0)


def remap_conv_scopes(tensors, layers):
    """Tolerant matching for checkpoints whose conv scopes are not the ones tf_scopes() predicts (another enclosing scope, or one flat
    numbering): order the checkpoint's '<scope>/weights' tensors by creation order (scope prefix in order of first appearance of the
    stem, then the auto-number) and accept the renaming only if the whole sequence of HWIO shapes equals this network's.  Returns
    {predicted scope: checkpoint scope} or None."""
    keys = [k[:-len("/weights")] for k, v in tensors.items() if k.endswith("/weights") and np.asarray(v).ndim == 4]
    if len(keys) != len(layers):
        return None
    def order(scope):
        head, idx = _scope_index(scope)
        return (0 if "imgproc" in head else 1, head, idx)       # the stem scope ('hg_imgproc', um_v1.py:84) is created first
    keys.sort(key=order)
    pred = tf_scopes(layers)
    for L, k in zip(layers, keys):
        if tuple(np.asarray(tensors[k + "/weights"]).shape) != (L["k"], L["k"], L["cin"], L["cout"]):
            return None
    return dict(zip(pred, keys))


def load_into_flat(tensors, layers, n_params, n_state, strict=True):
    """{tf name: ndarray} -> (params, state, adam_m, adam_v, global_step).  adam_* are None when the checkpoint carries no
    Adam slots.  Missing weights / beta / gamma / biases / moving statistics raise (strict) -- the zero-debias and r/d schedule
    variables are optional (inference does not read them) and default to the reference's initial values."""
    params = np.zeros(n_params, np.float32); state = np.zeros(n_state, np.float32)
    adam_m = np.zeros(n_params, np.float32); adam_v = np.zeros(n_params, np.float32)
    have_adam, missing = True, []
    vmap = variable_map(layers)
    if any(_lookup(tensors, name) is None for name, _, _, _ in vmap if name.endswith("/weights")):
        ren = remap_conv_scopes(tensors, layers)                  # predicted scope names absent: try creation order + shapes
        if ren:
            def renamed(name):
                for a, b in ren.items():
                    if name == a or name.startswith(a + "/"):
                        return (b + name[len(a):]).replace("/" + a + "/", "/" + b + "/")
                return name
            # longest scope first so that 'Conv_1' is not taken for a prefix of 'Conv_12'
            ren = dict(sorted(ren.items(), key=lambda kv: -len(kv[0])))
            vmap = [(renamed(name), buf, off, shape) for name, buf, off, shape in vmap]
    for name, buf, off, shape in vmap:
        n = int(np.prod(shape, dtype=np.int64))
        t = _lookup(tensors, name)
        if t is None:
            if name.endswith(_OPTIONAL_SUFFIXES):
                if name.endswith("/r_max"):
                    state[off] = 1.0                             # ops.py:110-114 initialisers
                continue
            missing.append(name)
            continue
        t = np.asarray(t, np.float32)
        if t.size != n or (t.ndim > 1 and tuple(t.shape) != tuple(shape)):
            raise CheckpointError("%s has shape %s, this network expects %s" % (name, tuple(t.shape), tuple(shape)))
        dst = params if buf == "params" else state
        dst[off:off + n] = t.reshape(-1)
        if buf == "params":
            m, v = tensors.get(name + "/Adam"), tensors.get(name + "/Adam_1")
            if m is None or v is None:
                have_adam = False
            else:
                adam_m[off:off + n] = np.asarray(m, np.float32).reshape(-1)
                adam_v[off:off + n] = np.asarray(v, np.float32).reshape(-1)
    if missing and strict:
        have = sorted(k for k in tensors if "/Adam" not in k)
        raise CheckpointError("checkpoint lacks %d variable(s) of this network, first: %s\ncheckpoint holds %d variables, e.g.:\n  %s"
                              % (len(missing), missing[0], len(have), "\n  ".join(have[:40])))
    step = int(np.asarray(tensors["global_step"]).reshape(-1)[0]) if "global_step" in tensors else 0
    return params, state, (adam_m if have_adam else None), (adam_v if have_adam else None), step


def flat_to_tensors(layers, params, state, adam_m=None, adam_v=None, global_step=0, beta1=0.5, beta2=0.999):
    """Inverse of load_into_flat: the variable set tf.train.Saver(tf.global_variables()) writes for this graph."""
    out = {"global_step": np.array(global_step, np.float32)}     # tf.get_variable('global_step', []) defaults to float32 (train_single_gpu.py:41)
    for name, buf, off, shape in variable_map(layers):
        n = int(np.prod(shape, dtype=np.int64))
        src = params if buf == "params" else state
        out[name] = np.asarray(src[off:off + n], np.float32).reshape(shape)
        if buf == "params" and adam_m is not None:
            out[name + "/Adam"] = np.asarray(adam_m[off:off + n], np.float32).reshape(shape)
            out[name + "/Adam_1"] = np.asarray(adam_v[off:off + n], np.float32).reshape(shape)
    if adam_m is not None:
        out["beta1_power"] = np.array(beta1 ** (global_step + 1), np.float32)
        out["beta2_power"] = np.array(beta2 ** (global_step + 1), np.float32)
    return out


def import_checkpoint(engine, prefix, strict=True):
    """saver.restore(sess, '<train_dir>/model.ckpt-<step>') (test_model.py:31-35, train_single_gpu.py:125-128) into a DenseRegEngine."""
    import torch
    tensors = read_bundle(prefix)
    params, state, m, v, step = load_into_flat(tensors, engine.layers(), engine.n_params, engine.n_state, strict)
    engine.load_flat(torch.from_numpy(params), torch.from_numpy(state))
    if m is not None and engine.adam_m is not None:
        engine.adam_m.copy_(torch.from_numpy(m)); engine.adam_v.copy_(torch.from_numpy(v))
    return step


def export_checkpoint(engine, prefix, global_step=0):
    """saver.save(sess, '<train_dir>/model.ckpt', global_step) (train_single_gpu.py:172-175): a bundle the reference can restore."""
    g = lambda t: None if t is None else t.detach().cpu().numpy()
    write_bundle(prefix, flat_to_tensors(engine.layers(), g(engine.params), g(engine.state), g(engine.adam_m), g(engine.adam_v), global_step))
    return prefix
