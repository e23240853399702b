"""Input pipeline (K15): file list -> threaded cv2 decode / crop / resize -> uint8 NHWC batches -> normalise."""
import os

import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

from edl_b200.utils import image_pipeline as ip  # noqa: E402


def _dataset(tmp_path, n=20):
    rng = np.random.RandomState(0)
    lines = []
    for i in range(n):
        h, w = int(rng.randint(40, 90)), int(rng.randint(40, 90))
        img = np.full((h, w, 3), (i * 10) % 255, dtype=np.uint8)
        img[:, : w // 2, 2] = 255                     # BGR: left half red -> detects channel order and flips
        name = "img_%03d.jpg" % i
        cv2.imwrite(str(tmp_path / name), img)
        lines.append("%s %d" % (name, i % 5))
    (tmp_path / "train_list.txt").write_text("\n".join(lines) + "\n")
    return ip.read_file_list(str(tmp_path / "train_list.txt"))


def test_loader_shards_epochs_and_shapes(tmp_path):
    samples = _dataset(tmp_path)
    assert len(samples) == 20 and samples[3][1] == 3 and os.path.isabs(samples[0][0])
    seen = []
    for rank in range(2):
        ld = ip.ImageBatchLoader(samples, 4, size=32, train=True, rank=rank, world=2, seed=7, threads=3, pin=False)
        assert len(ld) == 2
        batches = list(ld)
        assert len(batches) == 2
        for img, lab, flip in batches:
            assert img.shape == (4, 32, 32, 3) and img.dtype == torch.uint8
            assert lab.dtype == torch.int64 and flip.shape == (4,)
            seen += lab.tolist()
    assert len(seen) == 16
    ld0 = ip.ImageBatchLoader(samples, 4, size=32, train=True, seed=7, threads=2, pin=False)
    a = [lab.tolist() for _, lab, _ in ld0]
    ld0.set_epoch(1)
    b = [lab.tolist() for _, lab, _ in ld0]
    assert a != b                                      # reshuffled per epoch


def test_eval_decode_is_rgb_center_crop_and_normalise(tmp_path):
    samples = _dataset(tmp_path, 4)
    ld = ip.ImageBatchLoader(samples, 4, size=32, train=False, threads=2, pin=False, drop_last=False)
    (img, lab, flip), = list(ld)
    assert lab.tolist() == [0, 1, 2, 3] and int(flip.sum()) == 0
    assert img[0, 16, 2, 0] > 200 and img[0, 16, 29, 0] < 60          # red on the left, in channel 0 (RGB)
    x, y = ip.to_device_batch((img, lab, flip), "cpu", torch.float32)
    assert x.shape == (4, 3, 32, 32) and abs(float(x[0, 0, 16, 2]) - (1.0 - 0.485) / 0.229) < 0.1


def test_missing_file_raises(tmp_path):
    ld = ip.ImageBatchLoader([(str(tmp_path / "nope.jpg"), 0)] * 4, 4, size=16, train=False, threads=1, pin=False)
    with pytest.raises(IOError):
        list(ld)
