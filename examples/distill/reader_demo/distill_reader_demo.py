#!/usr/bin/env python
"""DistillReader API tour for the three reader formats (reference:
example/distill/reader_demo/distill_reader_demo.py).  Starts an in-process teacher so it runs anywhere:

    python examples/distill/reader_demo/distill_reader_demo.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from paddle_edl.distill.distill_reader import DistillReader  # noqa: E402
from paddle_edl.distill.teacher_server import TeacherServer  # noqa: E402


def sample_reader():
    for i in range(24):
        yield np.random.rand(1, 28, 28).astype("float32"), np.array([i % 10], dtype="int64")


def sample_list_reader():
    batch = []
    for s in sample_reader():
        batch.append(s)
        if len(batch) == 8:
            yield batch
            batch = []


def batch_reader():
    for b in sample_list_reader():
        yield np.stack([s[0] for s in b]), np.stack([s[1] for s in b])


def main():
    teacher = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(784, 10), torch.nn.Softmax(-1))
    with TeacherServer(teacher, ["img"], ["prediction"], {"img": [1, 28, 28]}) as srv:
        for name, setter, gen in (("sample", "set_sample_generator", sample_reader),
                                  ("sample_list", "set_sample_list_generator", sample_list_reader),
                                  ("batch", "set_batch_generator", batch_reader)):
            dr = DistillReader(ins=["img", None], predicts=["prediction"])   # the label slot is not sent to the teacher
            dr.set_teacher_batch_size(4)
            dr.set_fixed_teacher(srv.endpoint)
            reader = getattr(dr, setter)(gen)
            dr.print_config()
            n = 0
            for item in reader():
                n += 1
            first = item if name == "sample" else (item[0] if name == "sample_list" else tuple(a[0] for a in item))
            print("%-12s -> %d items, slots per sample: %d, prediction shape %s" % (name, n, len(first), np.shape(first[-1])))
            dr.stop()


if __name__ == "__main__":
    main()
