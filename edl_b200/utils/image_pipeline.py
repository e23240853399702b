"""ImageNet-style input pipeline.

The reference preprocesses on CPU threads with OpenCV (decode, random-resized crop, flip, normalise to fp32
NCHW; example/distill/resnet/utils/{img_tool.py:106-157, reader_cv2.py:27-105}) or on the GPU with DALI
(example/distill/resnet/dali.py:37-106).  Here the CPU part stops at **uint8 NHWC**: worker threads decode,
crop and resize straight into a pinned staging batch (a quarter of the fp32 bytes over PCIe/C2C), the flip
mask travels as one byte per image, and the fused ``normalize_u8`` kernel produces the bf16 channels_last
tensor on the device (``ops/misc.py``).

``ImageBatchLoader(decode="nvjpeg")`` is the DALI-style path (example/distill/resnet/dali.py:60-106): worker threads
only read file bytes, ``GpuJpegAugmenter`` decodes the batch on the GPU (nvJPEG, ``csrc/jpeg_decode.cpp``) and ONE kernel
does random-resized crop + flip + normalisation (``csrc/augment.cu``); pixels never visit the host.

File list format = the reference's ``train_list.txt`` / ``val_list.txt``: one ``relative/path.jpg label`` per line.
Elastic sharding: ``rank`` / ``world`` select every world-th line after a per-epoch shuffle with a shared seed.
"""
from __future__ import annotations

import math
import os
import queue
import random
import threading
import time
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
                 drop_last: bool = True, pin: Optional[bool] = None, decode: str = "cpu"):
        assert decode in ("cpu", "nvjpeg"), decode
        self.decode = decode
        self._augmenters = {}
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
                    if self.decode == "nvjpeg":                      # bytes only: the GPU decodes and augments
                        with open(path, "rb") as fh:
                            buf[j] = fh.read()
                    else:
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
                    if stop.is_set():
                        return
                    ids = idx[b * self.bs:(b + 1) * self.bs]
                    buf = [None] * len(ids) if self.decode == "nvjpeg" else alloc(len(ids))
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
                    if self.decode == "nvjpeg":
                        out_q.put(JpegBatch(self, buf, labels, self.seed * 7919 + self.epoch * 104729 + self.rank * 31 + b))
                    else:
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
            # Join before returning: a daemon thread that is still inside an OpenCV / torch call when the interpreter
            # finalises takes the process down with "terminate called without an active exception" (observed once in
            # three runs of the JPEG file-list example).  The feeder may be blocked on a full out_q: drain it.
            deadline = time.time() + 5.0
            while fd.is_alive() and time.time() < deadline:
                try:
                    out_q.get_nowait()
                except queue.Empty:
                    pass
                fd.join(timeout=0.05)
            for t in ws:
                t.join(timeout=max(0.0, deadline - time.time()))


    def augmenter(self, device) -> "GpuJpegAugmenter":
        key = str(device)
        if key not in self._augmenters:
            self._augmenters[key] = GpuJpegAugmenter(device, self.size)
        return self._augmenters[key]


class JpegBatch:
    """What ``ImageBatchLoader(decode="nvjpeg")`` yields: undecoded files + labels; ``to_device_batch`` turns it into
    the device batch.  ``seed`` makes the crops / flips of a batch reproducible whatever thread consumes it."""

    def __init__(self, loader: ImageBatchLoader, blobs: List[bytes], labels: torch.Tensor, seed: int):
        self.loader, self.blobs, self.labels, self.seed = loader, blobs, labels, seed

    def __len__(self):
        return len(self.blobs)


def eval_crop_box(h: int, w: int, size: int = 224, resize_short: int = 256):
    """The centre crop of ``decode_eval`` expressed in SOURCE pixels: resizing the short side to ``resize_short`` and
    cutting ``size`` x ``size`` out of the middle = resampling the centred square of side size/resize_short * short."""
    side = max(1, min(min(h, w), int(round(min(h, w) * size / float(resize_short)))))
    return (h - side) // 2, (w - side) // 2, side, side


def augment_reference(img: np.ndarray, box, flip: bool, size: int, mean=None, std=None) -> np.ndarray:
    """NumPy model of ``csrc/augment.cu`` for ONE image (uint8 [H, W, 3] -> float32 [size, size, 3]): bilinear
    resample of the crop box with cv2.INTER_LINEAR's geometry (half-pixel centres, edge clamp), mirror, normalise."""
    from ..ops.misc import IMAGENET_MEAN, IMAGENET_STD

    mean = IMAGENET_MEAN if mean is None else mean
    std = IMAGENET_STD if std is None else std
    y, x, ch, cw = box
    crop = img[y:y + ch, x:x + cw].astype(np.float32)

    def taps(n_src, n_dst, mirror):
        o = np.arange(n_dst, dtype=np.float32)
        if mirror:
            o = n_dst - 1 - o
        f = (o + 0.5) * (np.float32(n_src) / np.float32(n_dst)) - 0.5
        i0 = np.floor(f).astype(np.int64)
        a = (f - i0).astype(np.float32)
        a[i0 < 0] = 0.0
        i0 = np.clip(i0, 0, n_src - 1)
        i1 = np.clip(i0 + 1, 0, n_src - 1)
        return i0, i1, a

    y0, y1, ay = taps(ch, size, False)
    x0, x1, ax = taps(cw, size, flip)
    ay, ax = ay[:, None, None], ax[None, :, None]
    out = ((1 - ax) * (1 - ay) * crop[y0][:, x0] + ax * (1 - ay) * crop[y0][:, x1]
           + (1 - ax) * ay * crop[y1][:, x0] + ax * ay * crop[y1][:, x1])
    return ((out / 255.0 - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)).astype(np.float32)


class GpuJpegAugmenter:
    """JPEG files (bytes) -> normalised bf16 channels_last batch on ``device`` without a host round trip:
    nvJPEG batched decode into one pooled uint8 buffer (images keep their own sizes), then the fused
    crop / resize / flip / normalise kernel.  Images nvJPEG cannot take (CMYK, corrupt headers) are decoded with OpenCV
    and uploaded into the same pool.  All GPU work is queued on the caller's current stream."""

    ALIGN = 256

    def __init__(self, device, size: int = 224, backend: str = "default", cpu_threads: int = 4, mean=None, std=None):
        from .. import ops
        from ..ops.misc import IMAGENET_MEAN, IMAGENET_STD

        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GpuJpegAugmenter needs a CUDA device (use decode='cpu' on hosts without one)")
        self.size = size
        self.mean, self.std = list(mean or IMAGENET_MEAN), list(std or IMAGENET_STD)
        self._native = ops.native()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._dec = self._native.JpegDecoder(index, backend, cpu_threads)
        self._pool = torch.empty(0, dtype=torch.uint8, device=self.device)
        self._items_host = None

    def plan(self, dims, rng: random.Random, train: bool):
        """Crop boxes, flips and pool offsets for images of the given (h, w): int32 [N, 8] rows of ``AugmentItem``."""
        items = np.zeros((len(dims), 8), dtype=np.int32)
        offs = items.view(np.int64)                            # [N, 4] view (little endian)
        off = 0
        for i, (h, w) in enumerate(dims):
            y, x, ch, cw = _random_resized_crop_box(h, w, rng) if train else eval_crop_box(h, w, self.size)
            flip = 1 if (train and rng.random() < 0.5) else 0
            items[i, 2:] = (3 * w, y, x, ch, cw, flip)
            offs[i, 0] = off                                   # int64 offset = the first two int32 words of the row
            off += (h * 3 * w + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        return items, off

    def __call__(self, blobs: List[bytes], rng: Optional[random.Random] = None, train: bool = True) -> torch.Tensor:
        from .. import ops

        rng = rng or random.Random()
        info = self._dec.image_info(blobs)
        dims = [(h, w) for h, w, _ in info]
        items, total = self.plan(dims, rng, train)
        if self._pool.numel() < total:
            self._pool = torch.empty(int(total * 1.25), dtype=torch.uint8, device=self.device)
        offsets = [int(v) for v in items.view(np.int64)[:, 0]]
        gpu = [i for i, (_, _, c) in enumerate(info) if c in (1, 3)]
        with torch.cuda.device(self.device):
            if gpu:
                self._dec.decode([blobs[i] for i in gpu], self._pool, [offsets[i] for i in gpu],
                                 [3 * dims[i][1] for i in gpu], [dims[i][0] for i in gpu])
            for i in set(range(len(blobs))) - set(gpu):              # CMYK & co: CPU decode, same pool
                import cv2

                img = cv2.cvtColor(cv2.imdecode(np.frombuffer(blobs[i], np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
                flat = torch.from_numpy(np.ascontiguousarray(img)).reshape(-1)
                self._pool[offsets[i]:offsets[i] + flat.numel()].copy_(flat)
            y = torch.empty((len(blobs), self.size, self.size, 3), dtype=torch.bfloat16, device=self.device)
            self._items_host = torch.from_numpy(items).pin_memory()   # kept until the next call (async H2D source)
            self._native.crop_resize_normalize(self._pool, self._items_host.to(self.device, non_blocking=True), y,
                                               self.mean, self.std)
            ops.count_launch()
        return y.permute(0, 3, 1, 2)


def to_device_batch(batch, device, dtype=torch.bfloat16):
    """(uint8 NHWC pinned, labels, flips) -> (normalised channels_last images on ``device``, labels on ``device``);
    a ``JpegBatch`` (``decode="nvjpeg"``) is decoded and augmented on the device instead."""
    from .. import ops

    if isinstance(batch, JpegBatch):
        x = batch.loader.augmenter(device)(batch.blobs, random.Random(batch.seed), batch.loader.train)
        return (x.to(dtype) if x.dtype != dtype else x), batch.labels.to(device, non_blocking=True)
    img, labels, flips = batch
    x = ops.normalize_u8(img.to(device, non_blocking=True), flip=flips.to(device, non_blocking=True))
    return x.to(dtype) if x.dtype != dtype else x, labels.to(device, non_blocking=True)
