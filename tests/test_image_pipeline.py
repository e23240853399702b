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


# ------------------------------------------------------------------ DALI-style path: host-side pieces (no GPU needed)
def _photo(h, w, seed):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    return cv2.GaussianBlur(img, (5, 5), 0)


@pytest.mark.parametrize("box", [(10, 20, 200, 300), (0, 0, 375, 500), (100, 100, 50, 40), (7, 3, 301, 17), (0, 0, 64, 64)])
@pytest.mark.parametrize("flip", [False, True])
def test_augment_reference_has_cv2_bilinear_geometry(box, flip):
    """The NumPy model of csrc/augment.cu = crop -> cv2.resize(INTER_LINEAR) -> mirror, up to cv2's integer rounding."""
    img = _photo(375, 500, 1)
    y, x, ch, cw = box
    want = cv2.resize(img[y:y + ch, x:x + cw], (64, 64), interpolation=cv2.INTER_LINEAR).astype(np.float32)
    want = want[:, ::-1] if flip else want
    got = ip.augment_reference(img, box, flip, 64, mean=(0, 0, 0), std=(1, 1, 1)) * 255.0
    assert np.abs(got - want).max() < 1.0
    m = ip.augment_reference(img, box, flip, 64)                      # ImageNet statistics
    assert abs(float(m[5, 6, 1]) - (got[5, 6, 1] / 255.0 - 0.456) / 0.224) < 1e-4


def test_eval_crop_box_matches_resize_then_center_crop():
    img = _photo(300, 420, 2)
    box = ip.eval_crop_box(300, 420, 224, 256)
    assert box[2] == box[3] == round(300 * 224 / 256) and box[0] == (300 - box[2]) // 2
    s = 256 / 300.0
    big = cv2.resize(img, (int(round(420 * s)), 256), interpolation=cv2.INTER_LINEAR)
    yy, xx = (big.shape[0] - 224) // 2, (big.shape[1] - 224) // 2
    want = big[yy:yy + 224, xx:xx + 224].astype(np.float32)
    got = ip.augment_reference(img, box, False, 224, mean=(0, 0, 0), std=(1, 1, 1)) * 255.0
    assert np.abs(got - want).mean() < 4.0                            # same pixels up to a sub-pixel phase


def test_augment_plan_packs_items_for_the_kernel():
    import random

    aug = ip.GpuJpegAugmenter.__new__(ip.GpuJpegAugmenter)            # plan() is pure host code
    aug.size = 224
    dims = [(375, 500), (64, 48), (1200, 1600)]
    items, total = aug.plan(dims, random.Random(0), train=True)
    assert items.shape == (3, 8) and items.dtype == np.int32
    offs = items.view(np.int64)[:, 0]
    assert offs[0] == 0 and all(o % 256 == 0 for o in offs) and total >= offs[2] + 1200 * 1600 * 3
    assert offs[1] >= 375 * 500 * 3 and offs[2] >= offs[1] + 64 * 48 * 3
    for (h, w), row in zip(dims, items):
        pitch, y, x, ch, cw, flip = row[2:]
        assert pitch == 3 * w and 0 <= y and y + ch <= h and 0 <= x and x + cw <= w and flip in (0, 1)
    ev, _ = aug.plan(dims, random.Random(0), train=False)
    assert list(ev[0, 3:]) == [*ip.eval_crop_box(375, 500, 224), 0]


def test_nvjpeg_mode_ships_file_bytes(tmp_path):
    samples = _dataset(tmp_path, 8)
    ld = ip.ImageBatchLoader(samples, 4, size=32, train=True, seed=3, threads=2, pin=False, decode="nvjpeg")
    batches = list(ld)
    assert len(batches) == 2 and all(isinstance(b, ip.JpegBatch) and len(b) == 4 for b in batches)
    assert all(blob[:2] == b"\xff\xd8" for b in batches for blob in b.blobs)          # JPEG SOI marker
    assert batches[0].seed != batches[1].seed and batches[0].labels.dtype == torch.int64
    with pytest.raises(RuntimeError):
        ip.to_device_batch(batches[0], "cpu")                         # this path needs a CUDA device
