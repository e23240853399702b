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
    ap.add_argument("--teacher_bs", type=int, default=0, help="one teacher batch size (run.sh sweeps it); 0 = sweep 1..32 here")
    ap.add_argument("--conf_file", default="", help="serving conf (feeds / fetches): parse_config.get_ins_predicts")
    ap.add_argument("--out", default="", help="append one JSON line per measurement")
    ap.add_argument("--image_size", type=int, default=224)
    ap.add_argument("--reader_process", action="store_true",
                    help="DistillReader.set_reader_process(): reader in a forked process")
    ap.add_argument("--reader_cost_us", type=int, default=0,
                    help="pure-Python work per sample INSIDE the reader (emulates decode / augmentation that holds the GIL)")
    ap.add_argument("--consumer_cost_us", type=int, default=0,
                    help="pure-Python work per sample in the consumer (emulates the training loop's host side)")
    args = ap.parse_args()
    distill_worker._NOP_PREDICT_TEST = args.nop
    img = np.random.rand(3, args.image_size, args.image_size).astype("float32")

    def burn(us):                       # holds the GIL, like Python-level decode / augmentation code does
        t_end = time.perf_counter() + us * 1e-6
        while time.perf_counter() < t_end:
            pass

    def gen():
        for i in range(0, args.samples, 32):
            if args.reader_cost_us:
                burn(32 * args.reader_cost_us)
            yield [(img, np.array([j], dtype="int64")) for j in range(32)]

    ins, predicts = ["image", "label"], ["score"]
    if args.conf_file:
        from parse_config import get_ins_predicts
        feeds, _, _, predicts = get_ins_predicts(args.conf_file)
        ins = feeds + [n for n in ("label",) if n not in feeds]
    for tbs in ((args.teacher_bs,) if args.teacher_bs else (1, 2, 4, 8, 16, 32)):
        dr = DistillReader(ins=ins, predicts=predicts)
        dr.set_teacher_batch_size(tbs)
        dr.set_fixed_teacher(args.teachers)
        dr.set_reader_process(args.reader_process)
        r = dr.set_sample_list_generator(gen)
        t0, n = time.time(), 0
        for batch in r():
            n += len(batch)
            if args.consumer_cost_us:
                burn(len(batch) * args.consumer_cost_us)
        qps = n / (time.time() - t0)
        print("teacher_batch_size %2d: %8.1f samples/s" % (tbs, qps), flush=True)
        if args.out:
            import json
            with open(args.out, "a") as fh:
                fh.write(json.dumps({"teacher_batch_size": tbs, "samples_per_s": qps, "nop": bool(args.nop),
                                     "teachers": args.teachers, "samples": n, "image_size": args.image_size,
                                     "reader_process": bool(args.reader_process), "reader_cost_us": args.reader_cost_us,
                                     "consumer_cost_us": args.consumer_cost_us}) + "\n")
        dr.stop()


if __name__ == "__main__":
    main()
