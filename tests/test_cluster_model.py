"""Pod / Cluster JSON round trips, rank assignment, leader election, generator + barrier, watcher
(reference tests: test_pod.py, test_cluster.py, test_resource_pods.py, test_leader_pod.py,
test_cluster_generator.py, test_cluster_watcher.py)."""
import time

import pytest

from conftest import FakeJobEnv
from edl_b200.utils import (cluster as edl_cluster, cluster_generator, cluster_watcher, constants, exceptions,
                            leader_pod, pod_server, pod_server_client, resource_pods, status as edl_status,
                            train_status)
from edl_b200.utils.pod import Pod

TTL = constants.ETCD_TTL


def _pod(job_env):
    return Pod().from_env(job_env)


def test_pod_and_cluster_json_roundtrip(kv_server):
    je = FakeJobEnv(kv_server.endpoint, "j", nproc=3)
    je.gpus = ["0", "1", "2", "3", "4"]
    p = _pod(je)
    assert [len(t.gpus) for t in p.trainers] == [2, 2, 1]
    p2 = Pod().from_json(p.to_json())
    assert p == p2 and p2.trainers[1].endpoint == p.trainers[1].endpoint
    c = edl_cluster.Cluster()
    c._pods = [p, _pod(je)]
    c.new_stage()
    assert c.assign_ranks() == 6
    assert [t.global_rank for t in c.pods[1].trainers] == [3, 4, 5]  # no collisions across pods
    c2 = edl_cluster.Cluster().from_json(c.to_json())
    assert c == c2 and len(c2.get_trainers_endpoints()) == 6
    bad = c.to_dict()
    bad["pods"]["5"] = bad["pods"].pop("1")
    with pytest.raises(exceptions.EdlRankError):
        edl_cluster.Cluster().from_dict(bad)


def test_resource_register_ttl(etcd, kv_server):
    je = FakeJobEnv(kv_server.endpoint, "j")
    pods = [_pod(je), _pod(je)]
    regs = [resource_pods.Register(je, p.id, p.to_json(), etcd=etcd) for p in pods]
    time.sleep(TTL + 0.5)
    assert set(resource_pods.load_from_etcd(etcd)) == {p.id for p in pods}
    regs[1].stop()
    assert set(resource_pods.load_from_etcd(etcd)) == {pods[0].id}
    regs[0].stop()
    assert resource_pods.load_from_etcd(etcd) == {}


def test_leader_election_and_failover(etcd, kv_server):
    je = FakeJobEnv(kv_server.endpoint, "j", min_nodes=1, max_nodes=2)
    p0, p1 = _pod(je), _pod(je)
    r0 = resource_pods.Register(je, p0.id, p0.to_json(), etcd=etcd)
    r1 = resource_pods.Register(je, p1.id, p1.to_json(), etcd=etcd)
    l0 = leader_pod.Register(je, p0.id, etcd=etcd)
    l1 = leader_pod.Register(je, p1.id, etcd=etcd)
    assert l0.is_leader() and not l1.is_leader()
    assert leader_pod.get_pod_leader_id(etcd) == p0.id
    assert leader_pod.load_from_etcd(etcd).id == p0.id
    l0.stop()
    deadline = time.time() + TTL * 2 + 3
    while time.time() < deadline and not l1.is_leader():
        time.sleep(0.1)
    assert l1.is_leader() and leader_pod.get_pod_leader_id(etcd) == p1.id
    for x in (l1, r0, r1):
        x.stop()


def test_generator_barrier_scale_out_and_in(etcd, kv_server):
    je = FakeJobEnv(kv_server.endpoint, "j", min_nodes=2, max_nodes=3)
    pods = [_pod(je) for _ in range(3)]
    servers = [pod_server.PodServer(je, p.id, etcd=etcd).start() for p in pods]
    for p, s in zip(pods, servers):
        p.port = s.port
        edl_status.save_pod_status_to_etcd(etcd, p.id, edl_status.Status.INITIAL)
    regs = [resource_pods.Register(je, p.id, p.to_json(), etcd=etcd) for p in pods[:2]]
    leader = leader_pod.Register(je, pods[0].id, etcd=etcd)
    # barrier with only 1 of 2 pods arrived must time out
    cli = pod_server_client.Client(pods[0].endpoint)
    with pytest.raises(exceptions.EdlBarrierError):
        cli.barrier("j", pods[0].id, timeout=1.0)
    import threading
    out = {}
    t = threading.Thread(target=lambda: out.setdefault("c1", pod_server_client.Client(pods[0].endpoint).barrier(
        "j", pods[1].id, timeout=10)))
    t.start()
    c0 = cli.barrier("j", pods[0].id, timeout=10)
    t.join()
    assert c0.get_pods_ids_list() == [pods[0].id, pods[1].id] == out["c1"].get_pods_ids_list()
    assert c0.pods[0].id == pods[0].id  # leader is rank 0
    stage1 = c0.stage
    for p in pods[:2]:
        edl_status.save_pod_status_to_etcd(etcd, p.id, edl_status.Status.RUNNING)
    # ---- scale out: a third pod registers
    regs.append(resource_pods.Register(je, pods[2].id, pods[2].to_json(), etcd=etcd))
    deadline = time.time() + 10
    c = None
    while time.time() < deadline:
        c = edl_cluster.load_from_etcd(etcd)
        if c is not None and len(c.pods) == 3:
            break
        time.sleep(0.1)
    assert len(c.pods) == 3 and c.stage != stage1 and c.get_pods_ids_list()[:2] == [pods[0].id, pods[1].id]
    stage2 = c.stage
    # ---- near-the-end suppresses further scale-out (nothing to add here, but the flag must parse)
    train_status.save_to_etcd(etcd, pods[0].id, train_status.TrainStatus.NEARTHEEND)
    assert train_status.any_near_the_end(etcd, [pods[0].id])
    # ---- scale in: pod 1 dies (its lease expires)
    edl_status.save_pod_status_to_etcd(etcd, pods[2].id, edl_status.Status.RUNNING)
    regs[1].stop()
    deadline = time.time() + TTL * 2 + 10
    while time.time() < deadline:
        c = edl_cluster.load_from_etcd(etcd)
        if len(c.pods) == 2 and c.stage != stage2:
            break
        time.sleep(0.1)
    assert c.get_pods_ids_list() == [pods[0].id, pods[2].id]
    assert [t.global_rank for p in c.pods for t in p.trainers] == [0, 1]
    # ---- scheduler-driven resize through the leader's ScaleIn / ScaleOut RPCs
    je.min_nodes = 1                                     # (shared job env) wider range for this part
    regs[1] = resource_pods.Register(je, pods[1].id, pods[1].to_json(), etcd=etcd)       # pod 1 comes back
    train_status.save_to_etcd(etcd, pods[0].id, train_status.TrainStatus.RUNNING)
    edl_status.save_pod_status_to_etcd(etcd, pods[1].id, edl_status.Status.INITIAL)

    def wait_pods(n, not_stage):
        deadline = time.time() + 15
        while time.time() < deadline:
            cc = edl_cluster.load_from_etcd(etcd)
            if len(cc.pods) == n and cc.stage != not_stage:
                return cc
            time.sleep(0.1)
        raise AssertionError("cluster did not reach %d pods: %s" % (n, edl_cluster.load_from_etcd(etcd).get_pods_ids_list()))

    c3 = wait_pods(3, c.stage)
    lcli = pod_server_client.Client(pods[0].endpoint)
    with pytest.raises(exceptions.EdlLeaderError):
        pod_server_client.Client(pods[2].endpoint).scale_in(1)        # only the leader takes resize requests
    lcli.scale_in(1)
    c2 = wait_pods(2, c3.stage)
    assert c2.pods[0].id == pods[0].id and len(c2.get_pods_ids_set() & {pods[1].id, pods[2].id}) == 1
    for p in c2.pods:
        edl_status.save_pod_status_to_etcd(etcd, p.id, edl_status.Status.RUNNING)
    dropped = ({p.id for p in pods} - c2.get_pods_ids_set()).pop()
    edl_status.save_pod_status_to_etcd(etcd, dropped, edl_status.Status.INITIAL)   # an evicted pod re-enters as a joiner
    lcli.scale_out()
    c3b = wait_pods(3, c2.stage)
    assert c3b.get_pods_ids_set() == {p.id for p in pods}
    leader.stop()
    for r in (regs[0], regs[2]):
        r.stop()
    for s in servers:
        s.stop()


def test_non_leader_cannot_write_cluster(etcd, kv_server):
    je = FakeJobEnv(kv_server.endpoint, "j", min_nodes=1, max_nodes=1)
    p0, p1 = _pod(je), _pod(je)
    r0 = resource_pods.Register(je, p0.id, p0.to_json(), etcd=etcd)
    l0 = leader_pod.Register(je, p0.id, etcd=etcd)
    g = cluster_generator.Generator(je, p1.id, etcd=etcd)  # p1 is NOT the leader
    r1 = resource_pods.Register(je, p1.id, p1.to_json(), etcd=etcd)
    etcd.remove_server(constants.ETCD_CLUSTER, constants.ETCD_CLUSTER)
    with pytest.raises((exceptions.EdlTableError, exceptions.EdlGenerateClusterError)):
        g._generate_cluster_once()
    for x in (l0, r0, r1):
        x.stop()


def test_cluster_watcher(etcd, kv_server):
    je = FakeJobEnv(kv_server.endpoint, "j")
    c = edl_cluster.Cluster()
    c._pods = [_pod(je), _pod(je)]
    c.new_stage()
    c.assign_ranks()
    etcd.set_server_permanent(constants.ETCD_CLUSTER, constants.ETCD_CLUSTER, c.to_json())
    w = cluster_watcher.Watcher(je, c, etcd=etcd)
    time.sleep(0.6)
    assert not w.changed
    c.new_stage()
    etcd.set_server_permanent(constants.ETCD_CLUSTER, constants.ETCD_CLUSTER, c.to_json())
    deadline = time.time() + 5
    while time.time() < deadline and not w.changed:
        time.sleep(0.05)
    assert w.changed and w.get_new_cluster().stage == c.stage
    w.stop()


def test_proto_renderings_match_runtime_schema():
    import os
    from edl_b200.protos import schema

    here = os.path.dirname(schema.__file__)
    for name in schema.PROTO_FILES:
        text = schema.render_proto(name)
        assert open(os.path.join(here, os.path.basename(name))).read() == text, name + " is stale: python -m edl_b200.protos.schema"
    assert "rpc Barrier(BarrierRequest) returns (BarrierResponse)" in schema.render_proto("edl/pod_server.proto")


def test_step_meter_benchmark_log_and_metrics_endpoint(tmp_path):
    import json
    import time
    import urllib.request
    from edl_b200.utils.metrics import MetricsExporter, StepMeter, write_benchmark_log

    m = StepMeter(batch_per_trainer=32, world=8, window=5)
    for _ in range(12):
        time.sleep(0.002)
        m.step()
    s = m.summary()
    assert s["steps"] == 12 and 0 < s["avg_step_time_s"] < 0.1 and s["best_img_per_s"] > 0
    p = write_benchmark_log(3, s, str(tmp_path / "benchmark_logs"))
    assert json.load(open(p))["world"] == 8 and p.endswith("log_3")
    exp = MetricsExporter(port=0, host="127.0.0.1").start()
    try:
        exp.set("edl_img_per_s", 6700.5, {"job": "rn50", "stage": "s1"})
        exp.set("edl_world_size", 8)
        body = urllib.request.urlopen("http://127.0.0.1:%d/metrics" % exp.port, timeout=5).read().decode()
    finally:
        exp.stop()
    assert 'edl_img_per_s{job="rn50",stage="s1"} 6700.5' in body and "edl_world_size 8" in body
