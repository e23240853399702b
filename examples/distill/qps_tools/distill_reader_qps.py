#!/usr/bin/env python
"""QPS of the DistillReader alone for a sweep of teacher batch sizes (reference:
example/distill/qps_tools/distill_reader_qps.py + run.sh:23-28).

    python examples/distill/qps_tools/distill_reader_qps.py --teachers ip:port[,ip:port] [--nop]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from paddle_edl.distill import distill_worker  # noqa: E402
from paddle_edl.distill.distill_reader import DistillReader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--teachers", default="127.0.0.1:9292")
    ap.add_argument("--nop", action="store_true", help="use the NOP teacher (measures the pipeline only)")
    ap.add_argument("--samples", type=int, default=2048)
    ap.add_argument("--image_size", type=int, default=224)
    args = ap.parse_args()
    distill_worker._NOP_PREDICT_TEST = args.nop
    img = np.random.rand(3, args.image_size, args.image_size).astype("float32")

    def gen():
        for i in range(0, args.samples, 32):
            yield [(img, np.array([j], dtype="int64")) for j in range(32)]

    for tbs in (1, 2, 4, 8, 16, 32):
        dr = DistillReader(ins=["image", "label"], predicts=["score"])
        dr.set_teacher_batch_size(tbs)
        dr.set_fixed_teacher(args.teachers)
        r = dr.set_sample_list_generator(gen)
        t0, n = time.time(), 0
        for batch in r():
            n += len(batch)
        print("teacher_batch_size %2d: %8.1f samples/s" % (tbs, n / (time.time() - t0)), flush=True)
        dr.stop()


if __name__ == "__main__":
    main()
