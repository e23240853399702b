"""Input-pipeline throughput: OpenCV worker threads vs nvJPEG + fused augmentation kernel (DALI parity).

    python tools/bench_loader.py --images 2048 --batch 32 --threads 8 [--modes cpu,nvjpeg,nvjpeg_hw]

Writes ``--images`` synthetic JPEGs (ImageNet-like 500x375, quality 90) to a temp directory once, then times a full
pass of ``ImageBatchLoader`` + ``to_device_batch`` per mode (device batches are produced and dropped; with a GPU the
pass ends with a synchronize).  One JSON line per mode.  The number to beat is what one training GPU consumes:
6.7 k img/s (profiles/README.md).  Reference: example/distill/resnet/dali.py vs utils/reader_cv2.py.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edl_b200.utils import image_pipeline as ip  # noqa: E402


def make_dataset(root, n, h=375, w=500, quality=90):
    import cv2

    rng = np.random.RandomState(0)
    lines = []
    base = [cv2.GaussianBlur(rng.randint(0, 256, (h, w, 3)).astype(np.uint8), (0, 0), 3) for _ in range(8)]
    for i in range(n):
        img = np.roll(base[i % 8], (i * 7) % w, axis=1)                    # distinct files, photo-like spectrum
        name = "im%06d.jpg" % i
        cv2.imwrite(os.path.join(root, name), img, [cv2.IMWRITE_JPEG_QUALITY, quality])
        lines.append("%s %d" % (name, i % 1000))
    with open(os.path.join(root, "train_list.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return ip.read_file_list(os.path.join(root, "train_list.txt"))


def run_mode(samples, mode, args, dev):
    decode = "cpu" if mode == "cpu" else "nvjpeg"
    ld = ip.ImageBatchLoader(samples, args.batch, size=224, train=True, threads=args.threads, prefetch=8, decode=decode)
    if mode == "nvjpeg_hw":
        ld._augmenters[str(dev)] = ip.GpuJpegAugmenter(dev, 224, backend="hardware")
    n = 0
    for epoch in range(2):                                                  # epoch 0 = warm-up (page cache, pools)
        ld.set_epoch(epoch)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.time()
        n = 0
        for batch in ld:
            x, _ = ip.to_device_batch(batch, dev)
            n += x.shape[0]
        if dev.type == "cuda":
            torch.cuda.synchronize()
        dt = time.time() - t0
    return {"mode": mode, "images": n, "seconds": round(dt, 3), "img_per_s": round(n / dt, 1), "threads": args.threads,
            "batch": args.batch, "device": str(dev)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--modes", default="")
    args = ap.parse_args()
    cuda = torch.cuda.is_available()
    dev = torch.device("cuda", 0) if cuda else torch.device("cpu")
    modes = args.modes.split(",") if args.modes else (["cpu", "nvjpeg", "nvjpeg_hw"] if cuda else ["cpu"])
    with tempfile.TemporaryDirectory() as root:
        samples = make_dataset(root, args.images)
        for mode in modes:
            try:
                print(json.dumps(run_mode(samples, mode, args, dev)), flush=True)
            except Exception as e:  # noqa: BLE001  (e.g. no NVJPG engine backend on this GPU / driver)
                print(json.dumps({"mode": mode, "error": "%s: %s" % (type(e).__name__, str(e)[:200])}), flush=True)


if __name__ == "__main__":
    main()
