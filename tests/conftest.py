"""Shared fixtures: fast control-plane timing, the coordination store (Python and C++ builds), a store client."""
import os
import sys

import pytest

# fast control-plane timing for tests (the defaults mirror the reference: 15 s TTL, 3 s polls)
os.environ.setdefault("EDL_ETCD_TTL", "1.5")
os.environ.setdefault("EDL_POLL_INTERVAL", "0.3")
os.environ.setdefault("EDL_KILL_GRACE", "1")
os.environ.setdefault("EDL_BARRIER_TIMEOUT", "60")
os.environ.setdefault("EDL_RESCALE_BARRIER_TIMEOUT", "30")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu: test needs >= 2 CUDA devices")
    config.addinivalue_line("markers", "slow: long-running test")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        ngpu = torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        ngpu = 0
    skip_gpu = pytest.mark.skip(reason="needs a CUDA device")
    skip_multi = pytest.mark.skip(reason="needs >= 2 CUDA devices")
    # multi-process end-to-end modules take tens of seconds per test: they run against ONE store build (the C++
    # one when it can be built); the store semantics themselves are covered for both builds by test_store*.py,
    # test_cluster_model.py and test_discovery.py
    e2e = ("test_launch.py", "test_liveft.py", "test_jobserver_elastic.py", "test_inplace_rescale.py")
    drop = "[python]" if "native" in _kv_impls() else "[native]"
    kept, deselected = [], []
    for item in items:
        if item.nodeid.split("::")[0].endswith(e2e) and item.nodeid.endswith(drop):
            deselected.append(item)
        else:
            kept.append(item)
    if deselected:
        config.hook.pytest_deselected(items=deselected)
        items[:] = kept
    for item in items:
        if "gpu" in item.keywords and ngpu < 1:
            item.add_marker(skip_gpu)
        if "multigpu" in item.keywords and ngpu < 2:
            item.add_marker(skip_multi)


def _kv_impls():
    from edl_b200.store import native_server

    return ["python", "native"] if native_server.available() else ["python"]


@pytest.fixture(params=_kv_impls())
def kv_server(request):
    """The coordination store: every test that uses it runs against the Python server and the C++ one."""
    from edl_b200.store import KVServer, NativeKVServer

    srv = (NativeKVServer if request.param == "native" else KVServer)().start()
    yield srv
    srv.stop()


@pytest.fixture
def etcd(kv_server):
    import uuid
    from edl_b200.discovery.etcd_client import EtcdClient

    c = EtcdClient([kv_server.endpoint], root="job_" + uuid.uuid4().hex[:8])
    c.init()
    yield c
    c.close()


class FakeJobEnv:
    """Minimal JobEnv stand-in for in-process multi-pod tests (the reference's EtcdTestBase injects
    a fake PaddleCloud environment instead, tests/unittests/etcd_test_base.py:26-61)."""

    def __init__(self, endpoint, job_id, min_nodes=2, max_nodes=2, nproc=1):
        self.etcd_endpoints = [endpoint]
        self.job_id = job_id
        self.min_nodes, self.max_nodes = min_nodes, max_nodes
        self.nproc_per_node = nproc
        self.gpus = [str(i) for i in range(nproc)]
        from edl_b200.utils.network_utils import find_free_ports
        self.trainer_ports = [str(p) for p in find_free_ports(nproc)]
        self.log_dir = None
        self.hdfs_path = ""
