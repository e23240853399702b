"""Elastic data plane: file-list slicing, batch-id balancing / stealing on the leader DataServer,
and two Readers (two logical pods) that together consume every record exactly once
(reference: tests/unittests/test_data_server.py:61-95 and the stale test_data_reader.py)."""
import threading

import pytest

from edl_b200.collective.dataset import TxtFileSplitter
from edl_b200.collective.distribute_reader import Reader
from edl_b200.utils import data_server, data_server_client, exceptions
from edl_b200.utils.state import DataCheckpoint


def _files(tmp_path, sizes):
    out = []
    for i, n in enumerate(sizes):
        p = tmp_path / ("f%d.txt" % i)
        p.write_text("".join("file%d-line%d\n" % (i, j) for j in range(n)))
        out.append(str(p))
    return out


def test_leader_balancer_rpc(tmp_path):
    files = _files(tmp_path, [3, 3, 3])
    with data_server.DataServer("podA").start() as srv:
        srv.servicer.create_reader("r", files, ["podA", "podB"])
        c = data_server_client.Client()
        assert c.get_file_list(srv.endpoint, "r", "podA", files) == [(0, files[0]), (2, files[2])]
        assert c.get_file_list(srv.endpoint, "r", "podB", files) == [(1, files[1])]
        with pytest.raises(exceptions.EdlFileListNotMatchError):
            c.get_file_list(srv.endpoint, "r", "podB", files[:2], timeout=0.1)
        c.report_batch_data_meta(srv.endpoint, "r", "podA", "epA", ["a0", "a1", "a2", "a3", "a4", "a5"])
        metas = c.get_batch_data_meta(srv.endpoint, "r", "podB")       # podB has nothing: steals from podA
        assert metas[0].producer_pod_id == "podA" and metas[0].data_server_endpoint == "epA"
        stolen = list(metas[0].batch_data_ids)
        assert 1 <= len(stolen) <= 3
        own = [i for m in c.get_batch_data_meta(srv.endpoint, "r", "podA") for i in m.batch_data_ids]
        assert own and not set(own) & set(stolen)
        c.reach_data_end(srv.endpoint, "r", "podA")
        c.reach_data_end(srv.endpoint, "r", "podB")
        seen = set(stolen) | set(own)
        with pytest.raises(exceptions.EdlDataEndError):
            while True:
                for m in c.get_batch_data_meta(srv.endpoint, "r", "podB"):
                    seen |= set(m.batch_data_ids)
        assert seen == {"a0", "a1", "a2", "a3", "a4", "a5"}
        with pytest.raises(exceptions.EdlReaderNameError):
            c.get_batch_data_meta(srv.endpoint, "nope", "podA")
        c.close()


def test_two_pods_consume_everything_once(tmp_path):
    files = _files(tmp_path, [50, 3, 40, 7])      # uneven on purpose: pod A's slice is much bigger
    # small prefetch caches + a consumer that takes time per batch, like a training step does
    ra = Reader(files, TxtFileSplitter, batch_size=4, name="r", pod_id="podA", pod_ids=["podA", "podB"], is_leader=True,
                cache_capcity=2)
    rb = Reader(files, TxtFileSplitter, batch_size=4, name="r", pod_id="podB", pod_ids=["podA", "podB"], is_leader=False,
                leader_endpoint=ra.endpoint, cache_capcity=2)
    got = {"a": [], "b": []}

    def run(r, key):
        import time
        for item in r:
            got[key].extend(rec[1] for rec in item["data"])
            assert item["meta"]["end"] >= item["meta"]["begin"]
            time.sleep(0.01)

    ts = [threading.Thread(target=run, args=(ra, "a")), threading.Thread(target=run, args=(rb, "b"))]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    allrec = got["a"] + got["b"]
    assert len(allrec) == 100 and len(set(allrec)) == 100
    assert min(len(got["a"]), len(got["b"])) >= 20     # work stealing evened out the 90/10 file split
    ra.stop()
    rb.stop()


def test_resume_skips_processed_ranges(tmp_path):
    files = _files(tmp_path, [20])
    dc = DataCheckpoint("r", files)
    dc.mark(0, 0, 11)
    r = Reader(files, TxtFileSplitter, batch_size=4, name="r2", pod_id="p", pod_ids=["p"], is_leader=True,
               data_checkpoint=dc)
    lines = [rec[0] for item in r for rec in item["data"]]
    assert lines == list(range(12, 20))
    r.stop()
