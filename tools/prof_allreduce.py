#!/usr/bin/env python
"""ONE process, TWO (or more) GPUs, no process group: every GPU gets a thread that builds its side of a symmetric pool
through an in-process store and runs the fused all-reduce kernels.  Exists so that Nsight Compute can capture a
kernel that synchronises with peers: kernel replay cannot (the peers do not replay), application replay can --

    ncu --replay-mode application --set full --clock-control none -k regex:allreduce_twoshot -s 4 -c 1 \\
        -o gpurun_out/prof_allreduce python tools/prof_allreduce.py --mb 16 --algo multimem

and as a bandwidth microbenchmark without ncu (prints achieved GB/s against the NVLink 5 roofline).
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edl_b200 import ops  # noqa: E402
from edl_b200.parallel.symm import Fabric, SymmetricPool  # noqa: E402

NVLINK_DIR_GBPS = 900.0


class DictStore:
    """The smallest possible rendezvous store: a dict behind a condition variable."""

    def __init__(self):
        self._d, self._cv = {}, threading.Condition()

    def set(self, k, v):
        with self._cv:
            self._d[k] = v
            self._cv.notify_all()

    def get(self, k):
        with self._cv:
            if not self._cv.wait_for(lambda: k in self._d, timeout=60):
                raise RuntimeError("timed out waiting for " + k)
            return self._d[k]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=min(2, torch.cuda.device_count()))
    ap.add_argument("--mb", type=float, default=16.0)
    ap.add_argument("--algo", default="multimem", choices=["multimem", "twoshot", "fused"])
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--blocks", type=int, default=32)
    args = ap.parse_args()
    W = args.gpus
    assert W >= 2, "needs >= 2 GPUs"
    C = ops.native()
    store = DictStore()
    n = int(args.mb * (1 << 20)) // 2 // 8 * 8
    res, errs = {}, []
    go = threading.Barrier(W)

    def rank_main(r):
        try:
            dev = torch.device("cuda", r)
            torch.cuda.set_device(dev)
            pool = SymmetricPool(2 * n * 2 + (8 << 20), device=dev, fabric=Fabric(store, r, W, "prof"))
            g, p = pool.alloc(n, torch.bfloat16), pool.alloc(n, torch.bfloat16)
            g.tensor.normal_()
            master = torch.randn(n, device=dev)
            mom = torch.zeros(n, device=dev)
            lr = torch.tensor([0.1], device=dev)
            mm = args.algo != "twoshot" and pool.has_multicast

            def launch():
                if args.algo == "fused":
                    C.allreduce_sgd(g.data_ptrs, g.sig_ptrs, g.mc_ptr, p.data_ptrs, p.mc_ptr, r, master, mom, None, lr,
                                    1.0 / W, None, None, None, 0.9, 1e-4, False, mm, args.blocks, 30.0)
                else:
                    C.allreduce_twoshot(g.data_ptrs, g.sig_ptrs, g.mc_ptr, r, g.tensor, n, 1.0 / W, None, None, mm,
                                        args.blocks, 30.0)

            for _ in range(3):
                launch()
            torch.cuda.synchronize(dev)
            go.wait()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                launch()
            e1.record()
            torch.cuda.synchronize(dev)
            res[r] = (e0.elapsed_time(e1) / args.iters, pool.has_multicast, pool.check_error())
            go.wait()
        except Exception as e:  # noqa: BLE001
            errs.append("rank %d: %r" % (r, e))
            go.abort()

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(W)]
    t0 = time.time()
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise SystemExit("; ".join(errs))
    ms = max(v[0] for v in res.values())
    nbytes = n * 2
    bus = nbytes * 2 * (W - 1) / W / (ms * 1e-3) / 1e9
    print(json.dumps({"algo": args.algo, "world": W, "bytes": nbytes, "ms": ms, "algbw_GBps": nbytes / (ms * 1e-3) / 1e9,
                      "busbw_GBps": bus, "frac_of_nvlink_dir": bus / NVLINK_DIR_GBPS, "multicast": res[0][1],
                      "comm_error": max(v[2] for v in res.values()), "blocks": args.blocks,
                      "wall_s": round(time.time() - t0, 2)}))


if __name__ == "__main__":
    main()
