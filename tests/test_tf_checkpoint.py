"""TF V2 checkpoint bundle reader/writer + um_v1 variable-name map, CPU only.  The container format is pinned
by its published constants and a writer/reader round trip; the snappy path against pyarrow's snappy codec; the name map against
the layer table of the oracle (which tests/test_oracle_net.py ties to the CUDA library's table)."""
import os
import struct

import numpy as np
import pytest

from densereg_b200 import tf_checkpoint as T
from oracle import um_v1_torch as O


def _layers(S, F, J):
    specs, n_params, n_state = O.build_specs(S, F, J)
    return [dict(name=c.name, k=c.k, cin=c.cin, cout=c.cout, brn=int(c.brn), w_off=c.w_off, p_off=c.p_off, s_off=c.s_off) for c in specs], n_params, n_state


def test_table_roundtrip_multiblock_and_footer(tmp_path):
    p = str(tmp_path / "t.index")
    ents = [(b"", b"header")] + [(("scope/Conv_%03d/weights" % i).encode(), bytes([i % 251]) * (i % 37 + 1)) for i in range(400)]
    T.write_table(p, ents, block_size=512)
    raw = open(p, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw) > 48
    assert T.read_table(p) == ents
    bad = bytearray(raw); bad[10] ^= 0xFF
    open(p, "wb").write(bad)
    with pytest.raises(T.CheckpointError):
        T.read_table(p)
    bad = bytearray(raw); bad[-48 - 2] ^= 0xFF                   # a flipped checksum byte of the index block itself
    open(p, "wb").write(bad)
    with pytest.raises(T.CheckpointError):
        T.read_table(p)
    assert T.read_table(p, verify=False) == ents               # contents are intact; only the stored checksum differs
    open(p, "wb").write(raw[:-1] + b"\x00")
    with pytest.raises(T.CheckpointError):
        T.read_table(p)
    with pytest.raises(T.CheckpointError):
        T.write_table(p, [(b"b", b""), (b"a", b"")])


def test_snappy_matches_pyarrow(tmp_path):
    pa = pytest.importorskip("pyarrow")
    rng = np.random.RandomState(0)
    for data in (b"", b"a", b"hello hello hello hello world" * 40, rng.bytes(5000), bytes(70000), b"ab" * 40000 + rng.bytes(100)):
        comp = pa.compress(data, codec="snappy", asbytes=True)
        assert T.snappy_decompress(comp) == data
    # a table whose data block is stored snappy-compressed (type byte 1) reads the same
    ents = [(b"k%03d" % i, b"value-%d" % i * 5) for i in range(50)]
    block = T._build_block(ents)
    comp = pa.compress(block, codec="snappy", asbytes=True)
    out = bytearray(comp + b"\x01" + struct.pack("<I", T._mask(T.crc32c(comp + b"\x01"))))
    h_data = T._put_varint(0) + T._put_varint(len(comp))
    def emit(b):
        off = len(out); out.extend(b + b"\x00" + struct.pack("<I", T._mask(T.crc32c(b + b"\x00")))); return T._put_varint(off) + T._put_varint(len(b))
    meta = emit(T._build_block([])); idx = emit(T._build_block([(ents[-1][0], h_data)], 1))
    footer = meta + idx
    out.extend(footer + bytes(40 - len(footer)) + struct.pack("<Q", T.TABLE_MAGIC))
    p = str(tmp_path / "s.index"); open(p, "wb").write(out)
    assert T.read_table(p) == ents


def test_bundle_roundtrip_and_checksum(tmp_path):
    rng = np.random.RandomState(1)
    tensors = {"global_step": np.array(1200, np.float32), "Conv/weights": rng.randn(3, 3, 8, 16).astype(np.float32),
               "Conv/BatchReNorm/beta": rng.randn(16).astype(np.float32), "ids": np.arange(5, dtype=np.int64),
               "i32": np.array([[1, -2], [3, 4]], np.int32), "empty": np.zeros((0, 4), np.float32)}
    prefix = str(tmp_path / "model.ckpt-1200")
    T.write_bundle(prefix, tensors)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    back = T.read_bundle(prefix)
    assert set(back) == set(tensors)
    for k in tensors:
        assert back[k].dtype == tensors[k].dtype and back[k].shape == tensors[k].shape and np.array_equal(back[k], tensors[k])
    assert set(T.read_bundle(prefix, names={"ids"})) == {"ids"}
    d = bytearray(open(prefix + ".data-00000-of-00001", "rb").read()); d[7] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(d)
    with pytest.raises(T.CheckpointError):
        T.read_bundle(prefix)
    open(prefix + ".data-00000-of-00001", "wb").write(d[:20])
    with pytest.raises(T.CheckpointError):
        T.read_bundle(prefix, verify=False)


@pytest.mark.parametrize("S,F,J", [(1, 64, 16), (2, 128, 14)])
def test_variable_names_follow_tf_scopes(S, F, J):
    layers, n_params, n_state = _layers(S, F, J)
    sc = T.tf_scopes(layers)
    assert sc[0] == "hg_imgproc/Conv" and sc[1] == "hg_imgproc/Conv_1"
    n_stem = sum(1 for L in layers if L["name"].startswith("stem/"))
    assert n_stem == 1 + 4 + 3 + (3 if F == 64 else 4) and sc[n_stem - 1] == "hg_imgproc/Conv_%d" % (n_stem - 1)      # conv_1, Res(32,64), Res(64), Res(64,F)
    assert sc[n_stem] == "Conv" and sc[-1] == "Conv_%d" % (len(layers) - n_stem - 1)
    vm = T.variable_map(layers)
    names = [v[0] for v in vm]
    assert len(set(names)) == len(names)
    assert "hg_imgproc/Conv/BatchReNorm/hg_imgproc/Conv/BatchReNorm/moving_mean/biased" in names           # doubled zero-debias scope
    assert "Conv/BatchReNorm/gamma" in names
    # every flat parameter / state element is covered exactly once (local_step is stored once for both moving averages)
    cov_p = np.zeros(n_params, np.int32); cov_s = np.zeros(n_state, np.int32)
    for name, buf, off, shape in vm:
        n = int(np.prod(shape, dtype=np.int64))
        (cov_p if buf == "params" else cov_s)[off:off + n] += 0 if name.endswith("moving_variance/local_step") else 1
    assert cov_p.min() == 1 and cov_p.max() == 1 and cov_s.min() == 1 and cov_s.max() == 1
    bias_layers = [n for n in names if n.endswith("/biases")]
    assert len(bias_layers) == sum(1 for L in layers if not L["brn"])


def test_export_import_roundtrip(tmp_path):
    layers, n_params, n_state = _layers(1, 64, 16)
    rng = np.random.RandomState(2)
    params = rng.randn(n_params).astype(np.float32); state = rng.rand(n_state).astype(np.float32)
    m = rng.randn(n_params).astype(np.float32); v = rng.rand(n_params).astype(np.float32)
    prefix = str(tmp_path / "um" / "model.ckpt-300")
    T.write_bundle(prefix, T.flat_to_tensors(layers, params, state, m, v, global_step=300))
    tensors = T.read_bundle(prefix)
    assert tensors["hg_imgproc/Conv/weights"].shape == (7, 7, 1, 32) and tensors["Conv_2/weights"].shape[:2] == (1, 1)
    assert abs(float(tensors["beta1_power"]) - 0.5 ** 301) < 1e-30 or float(tensors["beta1_power"]) == np.float32(0.5 ** 301)
    p2, s2, m2, v2, step = T.load_into_flat(tensors, layers, n_params, n_state)
    assert step == 300 and np.array_equal(p2, params) and np.array_equal(s2, state) and np.array_equal(m2, m) and np.array_equal(v2, v)
    # inference-only checkpoint: no Adam slots, no zero-debias / schedule variables
    slim = {k: a for k, a in tensors.items() if not (k.endswith(("/Adam", "/Adam_1", "/biased", "/local_step", "/r_max", "/d_max", "/curr_t")) or "_power" in k)}
    p3, s3, m3, v3, _ = T.load_into_flat(slim, layers, n_params, n_state)
    assert m3 is None and v3 is None and np.array_equal(p3, params)
    L0 = layers[0]; c = L0["cout"]; s = L0["s_off"]
    assert np.array_equal(s3[s:s + 2 * c], state[s:s + 2 * c]) and s3[s + 4 * c] == 1.0 and s3[s + 4 * c + 1] == 0.0   # r_max=1, d_max=0 defaults
    # un-doubled zero-debias names are accepted too
    alt = dict(tensors)
    k = "Conv/BatchReNorm/Conv/BatchReNorm/moving_mean/biased"
    alt["Conv/BatchReNorm/moving_mean/biased"] = alt.pop(k)
    assert np.array_equal(T.load_into_flat(alt, layers, n_params, n_state)[1], state)
    # wrong architecture -> loud failure
    other, np2, ns2 = _layers(1, 64, 14)
    with pytest.raises(T.CheckpointError):
        T.load_into_flat(tensors, other, np2, ns2)
    del slim["Conv_5/weights"]
    with pytest.raises(T.CheckpointError) as ei:
        T.load_into_flat(slim, layers, n_params, n_state)
    assert "checkpoint holds" in str(ei.value) and "Conv_15/weights" in str(ei.value)    # the failure lists what the file does contain


def test_tolerant_name_matching(tmp_path):
    """pretrained bundles may use stock slim's 'BatchNorm' scope, or conv scopes under another enclosing scope / one flat numbering.
    Both are matched (the second by creation order + the full sequence of HWIO shapes); anything else still fails loudly."""
    layers, n_params, n_state = _layers(1, 64, 16)
    rng = np.random.RandomState(5)
    params = rng.randn(n_params).astype(np.float32); state = rng.rand(n_state).astype(np.float32)
    tensors = T.flat_to_tensors(layers, params, state)
    bn = {k.replace("BatchReNorm", "BatchNorm"): v for k, v in tensors.items()}
    p2, s2, *_ = T.load_into_flat(bn, layers, n_params, n_state)
    assert np.array_equal(p2, params) and np.array_equal(s2, state)
    # every conv under one enclosing scope with one flat numbering: tower/Conv, tower/Conv_1, ... (stem first)
    scopes = T.tf_scopes(layers)
    flat = {}
    for k, v in tensors.items():
        for i, sc in sorted(enumerate(scopes), key=lambda t: -len(t[1])):
            if k == sc or k.startswith(sc + "/"):
                new = "tower/Conv" + ("_%d" % i if i else "")
                flat[(new + k[len(sc):]).replace("/" + sc + "/", "/" + new + "/")] = v
                break
        else:
            flat[k] = v
    assert "tower/Conv_3/weights" in flat and "Conv_3/weights" not in flat
    p3, s3, *_ = T.load_into_flat(flat, layers, n_params, n_state)
    assert np.array_equal(p3, params) and np.array_equal(s3, state)
    flat["tower/Conv_7/weights"] = flat["tower/Conv_7/weights"][..., :-1]                  # one shape off -> no renaming, loud failure
    with pytest.raises(T.CheckpointError):
        T.load_into_flat(flat, layers, n_params, n_state)


@pytest.mark.gpu
def test_gpu_engine_export_import_same_xyz(built_lib, tmp_path):
    import torch
    from densereg_b200 import synth
    from densereg_b200.engine import DenseRegEngine
    a = DenseRegEngine(1, 64, 16, max_batch=2, training=True)
    a.init_params(seed=3, stddev=0.05)
    a.adam_m.normal_(); a.adam_v.uniform_()
    layers = a.layers()
    ref, n_params, n_state = _layers(1, 64, 16)
    key = lambda Ls: [(L["name"], L["w_off"], L["p_off"], L["s_off"] if L["brn"] else -1) for L in Ls]     # s_off is meaningless without BRN
    assert key(layers) == key(ref)
    prefix = T.export_checkpoint(a, str(tmp_path / "model.ckpt-7"), global_step=7)
    b = DenseRegEngine(1, 64, 16, max_batch=2, training=True)
    assert T.import_checkpoint(b, prefix) == 7
    assert torch.equal(a.params, b.params) and torch.equal(a.state, b.state) and torch.equal(a.adam_m, b.adam_m) and torch.equal(a.adam_v, b.adam_v)
    dms, poses, cfgs, coms = synth.make_batch(2, 16, seed=0)
    cu = lambda x: torch.from_numpy(x).cuda()
    assert torch.equal(a.infer(cu(dms), cu(cfgs), cu(coms)), b.infer(cu(dms), cu(cfgs), cu(coms)))


def test_golden_format_fixtures_are_stable(tmp_path):
    """The committed byte fixtures (tests/golden/make_format_golden.py) still parse and the writers still reproduce them byte for byte."""
    import importlib.util
    from densereg_b200 import png, tfrecord
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_format_golden", os.path.join(gdir, "make_format_golden.py"))
    G = importlib.util.module_from_spec(spec); spec.loader.exec_module(G)
    recs = list(tfrecord.read_records(os.path.join(gdir, "tfrecord_2examples.bin"), verify="all"))
    assert len(recs) == 2
    for rec, (img, feat) in zip(recs, G.examples()):
        got = tfrecord.parse_example(rec)
        assert rec == tfrecord.make_example(dict(feat, png16=got["png16"][0]))     # writer reproduces the record (PNG bytes taken as stored: zlib output may vary)
        assert got["name"] == [feat["name"]] and np.array_equal(got["xyz_pose"], feat["xyz_pose"])
        assert np.array_equal(png.decode_png(got["png16"][0]), img)
    assert recs and tfrecord.parse_example(recs[1])["bbx"].tolist() == [10.0, 20.0, 110.0, 140.0, 650.5]
    back = T.read_bundle(os.path.join(gdir, "ckpt_small"))
    want = G.tensors()
    assert set(back) == set(want) and all(np.array_equal(back[k], want[k]) and back[k].shape == want[k].shape for k in want)
    T.write_bundle(str(tmp_path / "ckpt_small"), want)
    for ext in (".index", ".data-00000-of-00001"):
        assert open(str(tmp_path / "ckpt_small") + ext, "rb").read() == open(os.path.join(gdir, "ckpt_small") + ext, "rb").read()
