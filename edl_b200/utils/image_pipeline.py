"""ImageNet-style input pipeline.

The reference preprocesses on CPU threads with OpenCV (decode, random-resized crop, flip, normalise to fp32
NCHW; example/distill/resnet/utils/{img_tool.py:106-157, reader_cv2.py:27-105}) or on the GPU with DALI
(example/distill/resnet/dali.py:37-106).  Here the CPU part stops at **uint8 NHWC**: worker threads decode,
crop and resize straight into a pinned staging batch (a quarter of the fp32 bytes over PCIe/C2C), the flip
mask travels as one byte per image, and the fused ``normalize_u8`` kernel produces the bf16 channels_last
tensor on the device (``ops/misc.py``).

File list format = the reference's ``train_list.txt`` / ``val_list.txt``: one ``relative/path.jpg label`` per line.
Elastic sharding: ``rank`` / ``world`` select every world-th line after a per-epoch shuffle with a shared seed.
"""
from __future__ import annotations

import math
import os
import queue
import random
import threading
from typing import List, Optional, Tuple

import numpy as np
import torch


def read_file_list(list_path: str, root: Optional[str] = None) -> List[Tuple[str, int]]:
    root = root if root is not None else os.path.dirname(os.path.abspath(list_path))
    out = []
    with open(list_path) as f:
        for line in f:
            line = line.strip()
            if line:
                path, label = line.rsplit(None, 1)
                out.append((os.path.join(root, path), int(label)))
    return out


def _random_resized_crop_box(h, w, rng, lower_scale=0.08, lower_ratio=3.0 / 4.0, upper_ratio=4.0 / 3.0):
    """Inception-style crop (the reference's random_crop, img_tool.py:60-90): area in [lower_scale, 1] of the
    image, log-uniform aspect ratio; falls back to the centre crop when no box fits after 10 draws."""
    area = h * w
    for _ in range(10):
        target = rng.uniform(lower_scale, 1.0) * area
        ratio = math.exp(rng.uniform(math.log(lower_ratio), math.log(upper_ratio)))
        cw, ch = int(round(math.sqrt(target * ratio))), int(round(math.sqrt(target / ratio)))
        if 0 < cw <= w and 0 < ch <= h:
            return rng.randint(0, h - ch), rng.randint(0, w - cw), ch, cw
    s = min(h, w)
    return (h - s) // 2, (w - s) // 2, s, s


def decode_train(path: str, size: int, rng: random.Random) -> np.ndarray:
    import cv2

    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise IOError("cannot decode %s" % path)
    y, x, ch, cw = _random_resized_crop_box(img.shape[0], img.shape[1], rng)
    img = cv2.resize(img[y:y + ch, x:x + cw], (size, size), interpolation=cv2.INTER_LINEAR)
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def decode_eval(path: str, size: int, resize_short: int = 256) -> np.ndarray:
    import cv2

    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise IOError("cannot decode %s" % path)
    h, w = img.shape[:2]
    s = resize_short / min(h, w)
    img = cv2.resize(img, (max(size, int(round(w * s))), max(size, int(round(h * s)))), interpolation=cv2.INTER_LINEAR)
    h, w = img.shape[:2]
    y, x = (h - size) // 2, (w - size) // 2
    return cv2.cvtColor(img[y:y + size, x:x + size], cv2.COLOR_BGR2RGB)


class ImageBatchLoader:
    """Iterable over (uint8 NHWC pinned batch, int64 labels, uint8 flip mask).  ``threads`` decoder threads fill
    ``prefetch`` staging batches ahead of the consumer; an epoch covers this rank's shard once."""

    def __init__(self, samples: List[Tuple[str, int]], batch_size: int, size: int = 224, train: bool = True,
                 rank: int = 0, world: int = 1, seed: int = 0, threads: int = 8, prefetch: int = 4,
                 drop_last: bool = True, pin: Optional[bool] = None):
        self.samples, self.bs, self.size, self.train = samples, batch_size, size, train
        self.rank, self.world, self.seed = rank, world, seed
        self.threads, self.prefetch, self.drop_last = max(1, threads), max(1, prefetch), drop_last
        self.pin = torch.cuda.is_available() if pin is None else pin
        self.epoch = 0

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def _shard(self):
        idx = list(range(len(self.samples)))
        if self.train:
            random.Random(self.seed + self.epoch).shuffle(idx)      # same permutation on every rank
        return idx[self.rank::self.world]

    def __len__(self):
        n = len(self._shard())
        return n // self.bs if self.drop_last else (n + self.bs - 1) // self.bs

    def __iter__(self):
        idx = self._shard()
        nb = len(self)
        out_q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        work_q: "queue.Queue" = queue.Queue()
        results = {}
        lock = threading.Condition()
        stop = threading.Event()

        def alloc(n):
            t = torch.empty((n, self.size, self.size, 3), dtype=torch.uint8)
            return t.pin_memory() if self.pin else t

        def worker(wid):
            rng = random.Random(self.seed * 1000003 + self.epoch * 1009 + self.rank * 131 + wid)
            while not stop.is_set():
                item = work_q.get()
                if item is None:
                    return
                b, j, si, buf = item
                path, _ = self.samples[si]
                try:
                    img = decode_train(path, self.size, rng) if self.train else decode_eval(path, self.size)
                    buf[j].copy_(torch.from_numpy(np.ascontiguousarray(img)))
                    err = None
                except Exception as e:  # noqa: BLE001
                    err = e
                with lock:
                    cnt, first_err = results[b]
                    results[b] = (cnt + 1, first_err or err)
                    lock.notify_all()

        def feeder():
            try:
                for b in range(nb):
                    ids = idx[b * self.bs:(b + 1) * self.bs]
                    buf = alloc(len(ids))
                    with lock:
                        results[b] = (0, None)
                    for j, si in enumerate(ids):
                        work_q.put((b, j, si, buf))
                    with lock:
                        while results[b][0] < len(ids) and not stop.is_set():
                            lock.wait(0.5)
                        _, err = results.pop(b)
                    if err is not None:
                        out_q.put(err)
                        return
                    labels = torch.tensor([self.samples[si][1] for si in ids], dtype=torch.int64)
                    flips = ((torch.rand(len(ids)) < 0.5).to(torch.uint8) if self.train
                             else torch.zeros(len(ids), dtype=torch.uint8))
                    out_q.put((buf, labels, flips))
                out_q.put(None)
            except Exception as e:  # noqa: BLE001
                out_q.put(e)

        ws = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(self.threads)]
        for t in ws:
            t.start()
        fd = threading.Thread(target=feeder, daemon=True)
        fd.start()
        try:
            while True:
                item = out_q.get()
                if item is None:
                    return
                if isinstance(item, Exception):
                    raise item
                yield item
        finally:
            stop.set()
            for _ in ws:
                work_q.put(None)


def to_device_batch(batch, device, dtype=torch.bfloat16):
    """(uint8 NHWC pinned, labels, flips) -> (normalised channels_last images on ``device``, labels on ``device``)."""
    from .. import ops

    img, labels, flips = batch
    x = ops.normalize_u8(img.to(device, non_blocking=True), flip=flips.to(device, non_blocking=True))
    return x.to(dtype) if x.dtype != dtype else x, labels.to(device, non_blocking=True)
