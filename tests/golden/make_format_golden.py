"""Golden byte fixtures for the on-disk formats restated without TensorFlow (regression guard; the formats are pinned against independent
implementations in tests/test_datasets.py / tests/test_tf_checkpoint.py: RFC 3720 CRC-32C vectors, the protobuf runtime, OpenCV, pyarrow snappy).
    python tests/golden/make_format_golden.py
Fixtures: tfrecord_2examples.bin (one TFRecord shard: two tf.train.Example records with name / xyz_pose / png16 [/ bbx]),
          ckpt_small.index + ckpt_small.data-00000-of-00001 (a tf.train.Saver V2 bundle with three variables)."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from densereg_b200 import png, tfrecord, tf_checkpoint  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def examples():
    out = []
    for i in range(2):
        img = (np.arange(4 * 6, dtype=np.uint16).reshape(4, 6) * 37 + 400 + i).astype(np.uint16)
        feat = {"name": b"2014_seq/image_%04d.png" % i, "xyz_pose": np.linspace(-50, 350, 48).astype(np.float32) + i,
                "png16": png.encode_png(img, filter_type=i + 1, level=9)}
        if i == 1:
            feat["bbx"] = [10.0, 20.0, 110.0, 140.0, 650.5]
        out.append((img, feat))
    return out


def tensors():
    return {"global_step": np.array(3, np.float32), "hg_imgproc/Conv/weights": (np.arange(7 * 7 * 1 * 2, dtype=np.float32) / 8).reshape(7, 7, 1, 2),
            "hg_imgproc/Conv/BatchReNorm/moving_mean": np.array([0.25, -1.5], np.float32)}


if __name__ == "__main__":
    with tfrecord.TFRecordWriter(os.path.join(OUT, "tfrecord_2examples.bin")) as w:
        for _, feat in examples():
            w.write(tfrecord.make_example(feat))
    tf_checkpoint.write_bundle(os.path.join(OUT, "ckpt_small"), tensors())
    print("format fixtures written to", OUT)
