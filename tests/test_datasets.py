"""Dataset readers without TensorFlow: TFRecord framing, tf.train.Example wire format, PNG codec, and the
ICVL / NYU / MSRA dataset objects end to end on tiny datasets written into tmp_path in the reference's directory layout.
The protobuf wire codec is pinned against the official protobuf runtime (dynamic descriptors of tensorflow/core/example/*.proto),
CRC-32C against its published check value, the PNG decoder against OpenCV."""
import os
import pickle
import struct

import numpy as np
import pytest

from densereg_b200 import datasets, png, tfrecord


# ---- TFRecord framing ------------------------------------------------------------------------------------------------------
def test_crc32c_known_answers():
    assert tfrecord.crc32c(b"123456789") == 0xE3069283                      # CRC-32C (Castagnoli) check value
    assert tfrecord.crc32c(b"") == 0
    assert tfrecord.crc32c(bytes(32)) == 0x8A9136AA                         # RFC 3720 B.4: 32 bytes of zeros
    assert tfrecord.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43                # RFC 3720 B.4: 32 bytes of 0xFF
    assert tfrecord.crc32c(bytes(range(32))) == 0x46DD794E                  # RFC 3720 B.4: 0x00..0x1F
    rng = np.random.RandomState(0)                                          # the lane-parallel path == the byte loop
    for n in (8191, 8192, 8193, 70001):
        d = rng.bytes(n)
        assert tfrecord.crc32c(d) == tfrecord._crc_scalar(d, 0xFFFFFFFF) ^ 0xFFFFFFFF
    assert tfrecord.crc32c(bytes(range(32)) * 1024) == tfrecord._crc_scalar(bytes(range(32)) * 1024, 0xFFFFFFFF) ^ 0xFFFFFFFF
    c = tfrecord.crc32c(b"abc")
    assert tfrecord.masked_crc32c(b"abc") == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def test_tfrecord_roundtrip_and_corruption(tmp_path):
    p = str(tmp_path / "shard")
    payloads = [b"", b"x", bytes(range(256)) * 7, b"last"]
    with tfrecord.TFRecordWriter(p) as w:
        for d in payloads:
            w.write(d)
    assert list(tfrecord.read_records(p, verify="all")) == payloads
    raw = bytearray(open(p, "rb").read())
    assert len(raw) == sum(16 + len(d) for d in payloads)                   # 8 length + 4 crc + data + 4 crc
    bad = bytearray(raw); bad[12 + 16 + 1 + 12 + 5] ^= 0x40                 # flip a bit inside the third record's data
    open(p, "wb").write(bad)
    assert len(list(tfrecord.read_records(p, verify="length"))) == 4
    with pytest.raises(tfrecord.TFRecordError):
        list(tfrecord.read_records(p, verify="all"))
    bad = bytearray(raw); bad[0] ^= 1                                       # corrupt the first length field
    open(p, "wb").write(bad)
    with pytest.raises(tfrecord.TFRecordError):
        list(tfrecord.read_records(p))
    open(p, "wb").write(raw[:-3])                                           # truncated tail
    with pytest.raises(tfrecord.TFRecordError):
        list(tfrecord.read_records(p))


# ---- tf.train.Example against the official protobuf runtime -------------------------------------------------------------------
def _example_classes():
    """Build tensorflow/core/example/{feature,example}.proto message classes with the protobuf runtime (no TensorFlow)."""
    pb = pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="densereg_test_example.proto", package="dr_test", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = fd.message_type.add(); m.name = name; return m

    def field(m, name, num, typ, label=T.LABEL_OPTIONAL, type_name=None, oneof=None, packed=None):
        f = m.field.add(); f.name, f.number, f.type, f.label = name, num, typ, label
        if type_name: f.type_name = ".dr_test." + type_name
        if oneof is not None: f.oneof_index = oneof
        if packed is not None: f.options.packed = packed
        return f

    field(msg("BytesList"), "value", 1, T.TYPE_BYTES, T.LABEL_REPEATED)
    field(msg("FloatList"), "value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, packed=True)
    field(msg("Int64List"), "value", 1, T.TYPE_INT64, T.LABEL_REPEATED, packed=True)
    f = msg("Feature"); f.oneof_decl.add().name = "kind"
    field(f, "bytes_list", 1, T.TYPE_MESSAGE, type_name="BytesList", oneof=0)
    field(f, "float_list", 2, T.TYPE_MESSAGE, type_name="FloatList", oneof=0)
    field(f, "int64_list", 3, T.TYPE_MESSAGE, type_name="Int64List", oneof=0)
    fs = msg("Features")
    entry = fs.nested_type.add(); entry.name = "FeatureEntry"; entry.options.map_entry = True
    field(entry, "key", 1, T.TYPE_STRING); field(entry, "value", 2, T.TYPE_MESSAGE, type_name="Feature")
    field(fs, "feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, type_name="Features.FeatureEntry")
    field(msg("Example"), "features", 1, T.TYPE_MESSAGE, type_name="Features")
    pool = descriptor_pool.DescriptorPool(); pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("dr_test.Example"))


def test_example_codec_matches_protobuf_runtime():
    Example = _example_classes()
    pose = (np.arange(48, dtype=np.float32) * 1.25 - 7).astype(np.float32)
    img = bytes(range(256)) * 3
    ex = Example()
    ex.features.feature["name"].bytes_list.value.append(b"201403121135/image_0000.png")
    ex.features.feature["xyz_pose"].float_list.value.extend(pose.tolist())
    ex.features.feature["png16"].bytes_list.value.append(img)
    ex.features.feature["ids"].int64_list.value.extend([0, 1, -1, 2 ** 40, -2 ** 62])
    got = tfrecord.parse_example(ex.SerializeToString())                    # official writer -> our reader
    assert got["name"] == [b"201403121135/image_0000.png"] and got["png16"] == [img]
    assert np.array_equal(got["xyz_pose"], pose) and got["xyz_pose"].dtype == np.float32
    assert got["ids"].tolist() == [0, 1, -1, 2 ** 40, -2 ** 62]
    ours = tfrecord.make_example({"name": b"a/b.png", "xyz_pose": pose, "png16": img, "bbx": [1.0, 2.0, 3.0, 4.0, 5.5]})
    back = Example(); back.ParseFromString(ours)                            # our writer -> official reader
    assert list(back.features.feature["name"].bytes_list.value) == [b"a/b.png"]
    assert np.array_equal(np.array(back.features.feature["xyz_pose"].float_list.value, np.float32), pose)
    assert list(back.features.feature["bbx"].float_list.value) == [1.0, 2.0, 3.0, 4.0, 5.5]
    assert back.features.feature["png16"].bytes_list.value[0] == img


def test_example_unpacked_float_list():
    # proto2 writers may emit repeated fixed32 un-packed: FloatList{1: f, 1: f}
    fl = b"".join(b"\x0d" + struct.pack("<f", v) for v in (1.5, -2.0))
    feat = tfrecord._ld(2, fl)
    ex = tfrecord._ld(1, tfrecord._ld(1, tfrecord._ld(1, b"p") + tfrecord._ld(2, feat)))
    assert tfrecord.parse_example(ex)["p"].tolist() == [1.5, -2.0]


# ---- PNG ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("filter_type", [0, 1, 2, 3, 4])
def test_png_roundtrip_all_filters(filter_type):
    rng = np.random.RandomState(filter_type)
    g16 = rng.randint(0, 65536, size=(13, 17)).astype(np.uint16)
    rgb = rng.randint(0, 256, size=(9, 11, 3)).astype(np.uint8)
    g8 = rng.randint(0, 256, size=(5, 7)).astype(np.uint8)
    for img in (g16, rgb, g8):
        out = png.decode_png(png.encode_png(img, filter_type))
        assert out.dtype == img.dtype and np.array_equal(out, img)


def test_png_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(3)
    yy, xx = np.mgrid[0:48, 0:64]
    depth = (600 + 40 * np.sin(xx / 7.0) + 30 * np.cos(yy / 5.0) + rng.randint(0, 3, (48, 64))).astype(np.uint16)
    ok, buf = cv2.imencode(".png", depth)                                   # libpng picks row filters adaptively
    assert ok and np.array_equal(png.decode_png(buf.tobytes()), depth)
    rgb = np.stack([np.zeros_like(depth), depth >> 8, depth & 0xFF], -1).astype(np.uint8)
    ok, buf = cv2.imencode(".png", rgb[..., ::-1])
    assert ok and np.array_equal(png.decode_png(buf.tobytes()), rgb)
    for ft in range(5):                                                     # our encoder -> OpenCV's decoder
        dec = cv2.imdecode(np.frombuffer(png.encode_png(depth, ft), np.uint8), cv2.IMREAD_UNCHANGED)
        assert dec.dtype == np.uint16 and np.array_equal(dec, depth)


def test_png_rejects_garbage():
    with pytest.raises(png.PngError):
        png.decode_png(b"not a png at all")
    good = png.encode_png(np.zeros((4, 4), np.uint16))
    with pytest.raises(png.PngError):
        png.decode_png(good[:40])


# ---- dataset objects ---------------------------------------------------------------------------------------------------------
def _depth_frame(rng, h, w, z=400):
    f = np.zeros((h, w), np.uint16)
    y0, x0 = rng.randint(20, h - 80), rng.randint(20, w - 80)
    f[y0:y0 + 60, x0:x0 + 50] = z + rng.randint(0, 40, (60, 50))
    return f


def _make_icvl(root, subset_dir, n, rng):
    cfg = datasets.IcvlDataset.cfg
    d = os.path.join(root, subset_dir, "Depth", "201403121135")
    os.makedirs(d)
    frames, uvds, lines = [], [], []
    for i in range(n):
        f = _depth_frame(rng, cfg.h, cfg.w)
        open(os.path.join(d, "image_%04d.png" % i), "wb").write(png.encode_png(f, filter_type=i % 5))
        uvd = np.stack([rng.uniform(60, 260, 16), rng.uniform(40, 200, 16), rng.uniform(300, 450, 16)], 1)
        frames.append(f); uvds.append(uvd)
        lines.append("201403121135/image_%04d.png " % i + " ".join("%.4f" % v for v in uvd.reshape(-1)))
    lines.insert(1, "test_seq_1/image_0000.png " + " ".join(["1.0"] * 48))       # dropped: does not start with '2014'
    open(os.path.join(root, subset_dir, "labels.txt"), "w").write("\n".join(lines) + "\n")
    return frames, uvds


def test_icvl_write_and_read_shards(tmp_path):
    rng = np.random.RandomState(0)
    root = str(tmp_path / "icvl")
    frames, uvds = _make_icvl(root, "Testing", 10, rng)
    ds = datasets.IcvlDataset("testing", directory=root)
    assert not ds.available()
    assert len(ds.loadAnnotation()) == 10                                   # the non-2014 row is skipped (icvl.py:104-105)
    a0 = ds.annotations[0]
    cfg = ds.cfg
    exp = np.stack([(uvds[0][:, 0] - cfg.cx) * uvds[0][:, 2] / cfg.fx, (uvds[0][:, 1] - cfg.cy) * uvds[0][:, 2] / cfg.fy, uvds[0][:, 2]], 1)
    assert a0.name == "201403121135/image_0000.png" and np.allclose(np.array(a0.pose).reshape(-1, 3), exp, atol=2e-3)
    assert np.allclose(datasets.xyz2uvd(a0.pose, cfg), uvds[0], atol=2e-3)       # uvd -> xyz -> uvd
    written = ds.write_TFRecord_multi_thread(num_threads=2, num_shards=4)   # icvl.py:156-157
    assert [os.path.basename(p) for p in written] == ["testing-%d-of-4" % i for i in range(4)]
    assert ds.available() and ds.exact_num == 1596 and len(ds.filenames) == 5
    # shard boundaries follow np.linspace(..).astype(int): threads [0,5),[5,10); shards [0,2),[2,5),[5,7),[7,10)
    counts = [len(list(tfrecord.read_records(p, verify="all"))) for p in written]
    assert counts == [2, 3, 2, 3]
    ex = list(ds.examples())
    assert len(ex) == 13                                                    # the last shard is listed twice (icvl.py:74)
    for i, (image, pose, name, bbx) in enumerate(ex[:10]):
        assert image.dtype == np.float32 and np.array_equal(image, frames[i].astype(np.float32))
        assert name == "201403121135/image_%04d.png" % i and bbx is None
        assert pose.shape == (48,) and np.allclose(pose, np.array(ds.annotations[i].pose, np.float32))
    it = ds.examples()
    f, p, names, bb = ds.frame_batch(4, it)
    assert f.shape == (4, 240, 320) and p.shape == (4, 48) and len(names) == 4 and bb is None
    ds.frame_batch(4, it); ds.frame_batch(4, it)
    tail = ds.frame_batch(4, it, allow_partial=True)
    assert len(tail[2]) == 1
    with pytest.raises(datasets.EndOfData):
        ds.frame_batch(4, it)


def test_icvl_shuffled_training_stream(tmp_path):
    rng = np.random.RandomState(1)
    root = str(tmp_path / "icvl")
    _make_icvl(root, "Training", 6, rng)
    ds = datasets.IcvlDataset("training_small", directory=root)
    ds.loadAnnotation()
    os.makedirs(ds.tf_dir)
    ds.saveSampleToRecord(range(6), ds.filenames[0])                        # training_small reads training-0-of-100 only
    assert ds.available() and ds.approximate_num == 220
    names_a = [e[2] for e in ds.examples(shuffle=True, seed=5, epochs=2)]
    names_b = [e[2] for e in ds.examples(shuffle=True, seed=5, epochs=2)]
    assert names_a == names_b and len(names_a) == 12 and sorted(set(names_a)) == sorted(a.name for a in ds.annotations)
    assert names_a != [e[2] for e in ds.examples(shuffle=True, seed=6, epochs=2)]
    it = ds.examples(shuffle=True, seed=0, epochs=None)                     # endless
    assert len([next(it) for _ in range(40)]) == 40


def test_nyu_depth_decode_joint_selection_and_boxes(tmp_path):
    sio = pytest.importorskip("scipy.io")
    rng = np.random.RandomState(2)
    root = str(tmp_path / "nyu")
    cfg = datasets.NyuDataset.cfg
    n = 3
    for sub, cams in (("dataset/train", 3), ("dataset/test", 1)):
        d = os.path.join(root, sub); os.makedirs(d)
        joints = rng.uniform(-200, 200, size=(cams, n, 36, 3)); joints[..., 2] = rng.uniform(600, 900, size=(cams, n, 36))
        sio.savemat(os.path.join(d, "joint_data.mat"), {"joint_xyz": joints})
        for c in range(cams):
            for i in range(n):
                depth = _depth_frame(rng, cfg.h, cfg.w, z=700)
                rgb = np.stack([rng.randint(0, 256, depth.shape), depth >> 8, depth & 0xFF], -1).astype(np.uint8)
                open(os.path.join(d, "depth_%d_%07d.png" % (c + 1, i + 1)), "wb").write(png.encode_png(rgb, filter_type=(i + c) % 5))
                if sub.endswith("train") and c == 0 and i == 0:
                    first_depth, first_joints = depth, joints[0, 0].copy()
    bbx = [np.array([[100.0 + i], [120.0], [300.0], [330.0], [950.0]], np.float32) for i in range(n)]     # nyu_bbx.pkl layout (5,1)
    bbx_path = str(tmp_path / "nyu_bbx.pkl")
    pickle.dump(bbx, open(bbx_path, "wb"), protocol=2)

    tr = datasets.NyuDataset("training", directory=root)
    assert tr.jnt_num == 14 and tr.pose_dim == 42 and len(tr.filenames) == 101
    assert len(tr.loadAnnotation()) == 9 and tr.annotations[4].name == "depth_2_0000002.png"
    os.makedirs(tr.tf_dir)
    tr.saveSampleToRecord(range(9), os.path.join(tr.tf_dir, "training-0-of-300"))
    val = datasets.NyuDataset("training_small", directory=root)            # files 0,10,20 of 300
    for k in (10, 20):
        tr.saveSampleToRecord([], os.path.join(tr.tf_dir, "training-%d-of-300" % k))
    image, pose, name, b = next(val.examples())
    assert name == "depth_1_0000001.png" and b is None
    assert np.array_equal(image, first_depth.astype(np.float32))            # depth = G*256 | B (nyu.py:151-155)
    flipped = first_joints * np.array([1.0, -1.0, 1.0])                     # y negated on load (nyu.py:118)
    keep = [0, 3, 6, 9, 12, 15, 18, 21, 24, 25, 27, 30, 31, 32]
    assert np.allclose(pose.reshape(14, 3), flipped[keep].astype(np.float32))

    te = datasets.NyuDataset("testing", directory=root, bbx_path=bbx_path)
    assert len(te.loadAnnotation()) == 3 and te.exact_num == 8252
    te.write_TFRecord_multi_thread(num_threads=1, num_shards=16)
    assert te.available()
    ex = list(te.examples())
    assert len(ex) == 4 and ex[1][3].tolist() == [101.0, 120.0, 300.0, 330.0, 950.0]   # last shard listed twice (nyu.py:80)
    assert ex[3][2] == ex[2][2] == "depth_1_0000003.png"
    frames, poses, names, bb = te.frame_batch(3, iter(ex))
    assert bb.shape == (3, 5) and bb.dtype == np.float32 and frames.shape == (3, 480, 640)


def test_msra_bin_to_png_and_shards(tmp_path):
    rng = np.random.RandomState(4)
    root = str(tmp_path / "msra15")
    cfg = datasets.MsraDataset.cfg
    full = {}
    for g in datasets.MsraDataset.pose_list:
        d = os.path.join(root, "P3", g); os.makedirs(d)
        rows = []
        for i in range(2):
            left, top, right, bottom = 100 + i, 60, 180, 150
            crop = rng.uniform(300, 500, size=(bottom - top, right - left)).astype(np.float32)
            if g == "2" and i == 1:
                crop[:] = 0                                                 # empty frame -> repeats the previous one (msra.py:139-143)
            with open(os.path.join(d, "%06d_depth.bin" % i), "wb") as f:
                f.write(struct.pack("<6i", cfg.w, cfg.h, left, top, right, bottom)); f.write(crop.tobytes())
            fr = np.zeros((cfg.h, cfg.w), np.float32); fr[top:bottom, left:right] = crop
            full[(g, i)] = fr
            rows.append(" ".join("%.3f" % v for v in rng.uniform(-100, 100, 63)))
        open(os.path.join(d, "joint.txt"), "w").write("2\n" + "\n".join(rows) + "\n")
    ds = datasets.MsraDataset("testing", 3, directory=root)
    assert ds.name == "msra_P3" and ds.exact_num == 8488
    ann = ds.loadAnnotation()
    assert len(ann) == 34 and ann[0].name == os.path.join("1", "000000_depth")
    raw = np.array([float(v) for v in open(os.path.join(root, "P3", "1", "joint.txt")).read().split("\n")[1].split()]).reshape(-1, 3)
    assert np.allclose(np.array(ann[0].pose).reshape(-1, 3), raw * np.array([1.0, -1.0, -1.0]))   # msra.py:104-110
    ds.cvtBin2Png()
    ds.write_TFRecord_multi_thread(num_threads=20, num_shards=100)          # msra.py:214
    assert ds.available() and os.path.basename(ds.filenames[0]) == "P3-0-of-100"
    ex = list(ds.examples())
    assert len(ex) == 34 + len(list(tfrecord.read_records(ds.filenames[-1])))
    by_name = {e[2]: e[0] for e in ex}
    assert np.array_equal(by_name[os.path.join("1", "000001_depth")], full[("1", 1)].astype(np.uint16).astype(np.float32))
    assert np.array_equal(by_name[os.path.join("2", "000001_depth")], full[("2", 0)].astype(np.uint16).astype(np.float32))
    tr = datasets.MsraDataset("training", 3, directory=root)
    assert len(tr.filenames) == 801 and not any("P3-" in os.path.basename(p) for p in tr.filenames)


def test_open_dataset_switch():
    assert isinstance(datasets.open_dataset("icvl", "training", directory="/nonexistent"), datasets.IcvlDataset)
    assert datasets.open_dataset("msra", "testing", pid=2, directory="/nonexistent").pid == 2
    assert not datasets.open_dataset("nyu", "testing", directory="/nonexistent").available()
    with pytest.raises(ValueError):
        datasets.open_dataset("bighand", "training")
    with pytest.raises(ValueError):
        datasets.IcvlDataset("bogus")


# ---- frames -> crops -> xyz on the GPU ---------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_icvl_shards_to_crops_match_oracle(built_lib, tmp_path):
    import torch
    from oracle import crop_numpy as C
    from densereg_b200 import synth
    from densereg_b200.engine import DenseRegEngine
    frames, poses, cfg = synth.make_frames(6, 16, "icvl", seed=3)
    root = str(tmp_path / "icvl")
    d = os.path.join(root, "Testing", "Depth", "2014_seq"); os.makedirs(d)
    ds = datasets.IcvlDataset("testing", directory=root)
    f16 = np.clip(np.rint(frames), 0, 65535).astype(np.uint16)
    ds._annotations = []
    for i in range(6):
        open(os.path.join(d, "image_%04d.png" % i), "wb").write(png.encode_png(f16[i], filter_type=4))
        ds._annotations.append(datasets.Annotation("2014_seq/image_%04d.png" % i, poses[i].tolist()))
    ds.write_TFRecord_multi_thread(num_threads=2, num_shards=4)
    eng = DenseRegEngine(1, 64, 16, max_batch=4, training=False)
    eng.init_params(seed=0)
    dms, p_d, cfgs, coms, names = ds.batch_device(eng, 4)
    torch.cuda.synchronize()
    assert names == ["2014_seq/image_%04d.png" % i for i in range(4)]
    for b in range(4):
        crop, ncfg = C.crop_from_xyz_pose(f16[b].astype(np.float32), poses[b], cfg, icvl=True)
        assert np.array_equal(dms[b, :, :, 0].cpu().numpy(), crop), "crop from the decoded shard is not bit-exact"
        np.testing.assert_allclose(cfgs[b].cpu().numpy(), ncfg, rtol=1e-6)
    xyz = eng.infer(dms, cfgs, coms)
    assert xyz.shape == (4, 48) and bool(torch.isfinite(xyz).all())
    lo_hi = ds.batch_device(eng, 4, lo=1, hi=2)                             # a rank's shard of the next (partial) batch
    assert lo_hi[0].shape[0] == 1 and len(lo_hi[4]) == 1
