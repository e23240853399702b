"""Trainer stand-in for the elastic data plane under IN-PLACE rescale: every pod reads its share of a file list through
``collective.distribute_reader.Reader`` (leader-balanced batches, work stealing), records what it "trained" on, and
survives membership changes without restarting -- the consumed ranges of all pods are merged after every stage
rendezvous and the re-created reader skips them.  The test asserts that every record was consumed exactly once."""
import json
import os
import sys
import time

from edl_b200 import elastic
from edl_b200.collective.dataset import TxtFileSplitter
from edl_b200.collective.distribute_reader import Reader
from edl_b200.utils.state import DataCheckpoint


def main():
    files = sys.argv[1].split(",")
    out_dir = os.environ["READER_DEMO_OUT"]
    step_s = float(os.environ.get("READER_DEMO_STEP", "0.03"))
    ctx = elastic.ElasticContext("gloo")
    info = ctx.start()
    job, kv = ctx.env.job_id, ctx.kv
    dc = DataCheckpoint("demo", files)
    consumed = []
    log = open(os.path.join(out_dir, "consumed_%d.jsonl" % os.getpid()), "a")
    while True:
        # everybody of this stage exchanges what has been trained on so far (joiners bring nothing)
        for other in ctx.allgather_object(dc.processed_data):
            dc.merge(other)
        pods = ctx.pod_ids
        name = "demo-%s" % info.stage            # one reader per stage, the same name on every pod
        # a small cache: batches are handed out at the pace they are trained on, so stealing balances the PODS' work
        # (with the default of 100 a fast accesser prefetches most of the epoch and the other pod idles at the end)
        reader = Reader(files, TxtFileSplitter, 4, cache_capcity=2, name=name, pod_id=ctx.env.pod_id, pod_ids=pods,
                        is_leader=(ctx.env.pod_id == pods[0]), etcd=ctx.etcd, data_checkpoint=dc)
        finished = True
        print("rank %d: reader %s over pods %s (leader %s)" % (info.rank, name[:13], [x[:6] for x in pods], pods[0][:6]), flush=True)
        for batch in reader:
            time.sleep(step_s)                                  # the "training step"
            meta = batch["meta"]
            dc.mark(meta["file_idx"], meta["begin"], meta["end"])
            for rec in batch["data"]:
                consumed.append((meta["file_idx"], rec[0]))
                log.write(json.dumps({"file": meta["file_idx"], "rec": rec[0], "stage": info.stage, "pid": os.getpid(),
                                      "world": info.size}) + "\n")
            log.flush()
            if ctx.should_switch():
                finished = False
                break
        reader.stop()
        done_prefix = "/%s/reader_demo/done/%s/" % (job, info.stage)
        if finished:
            kv.put(done_prefix + str(info.rank), b"1")
            while True:                                          # epoch-end barrier through the store
                if len(kv.get_prefix(done_prefix)[0]) >= info.size:
                    print("rank %d: epoch complete, %d records here" % (info.rank, len(consumed)), flush=True)
                    ctx.close()
                    return 0
                if ctx.should_switch():
                    finished = False
                    break
                time.sleep(0.05)
        try:
            old = info
            info = ctx.rescale()
            print("reader demo rescaled in place: %d -> %d, pid %d" % (old.size, info.size, os.getpid()), flush=True)
        except elastic.EdlEvicted:
            ctx.close()
            return 0


if __name__ == "__main__":
    sys.exit(main())
