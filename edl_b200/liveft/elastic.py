"""``ElasticManager`` -- "the job runs iff exactly np nodes are registered" rendezvous
(reference: python/edl/liveft/elastic.py:36-313; doc/edl_live_fault_tolerance.md).

Store layout under ``/paddle/<job>``: the prefix key itself holds ``0`` (running) / ``1``
(completed); ``nodes/<timestamp> = host`` per live node (re-put by a watch if somebody deletes it);
``np`` = target node count (a scheduler resizes the job by writing it); ``endpoints`` =
``DISTRIBUTED_TRAINER_ENDPOINTS|PADDLE_TRAINERS``.

``watch()`` returns COMPLETED / RESTART (exit code 101) / ERROR / HOLD (membership != np: children are
stopped and the caller goes back to ``wait()``) / EXIT.  Fault-tolerance level
(``PADDLE_ELASTIC_FAULT_TOLERANC_LEVEL``): 1 = restart on failure, otherwise report the error.

Unlike the reference (whose ``LauncherInterface`` is abstract and whose ``ElasticManager()`` is
called without its required argument) this module is runnable: ``ProcessLauncher`` starts the
training command with the rank environment and node keys are leased, so a crashed node disappears
by itself.
"""
import logging
import os
import signal
import socket
import subprocess
import sys
import threading
import time

from ..store.client import KVClient

logger = logging.getLogger("edl.liveft")

ELASTIC_EXIT_CODE = 101


class ElasticStatus:
    COMPLETED = "completed"
    ERROR = "error"
    HOLD = "hold"
    RESTART = "restart"
    EXIT = "exit"


class _Proc:
    def __init__(self, proc, rank, log_fn=None):
        self.proc, self.rank, self.log_fn = proc, rank, log_fn


class LauncherInterface:
    def __init__(self, args):
        self.args = args
        self.procs = []

    def _terminate_procs(self, timeout=50):
        for p in self.procs:
            if p.proc.poll() is None:
                p.proc.terminate()
                if p.log_fn:
                    p.log_fn.close()
                logger.info("terminate process id:%d", p.proc.pid)
        deadline = time.time() + timeout
        while time.time() < deadline:
            alive = False
            for p in self.procs:
                if p.proc.poll() is None:
                    alive = True
                    if deadline - time.time() < timeout - 3:
                        os.kill(p.proc.pid, signal.SIGKILL)
            if not alive:
                logger.info("terminated all the procs")
                return True
            time.sleep(0.2)
        return False

    def _check_procs(self):
        """None while running; 0 when all exited cleanly; else the failing exit code."""
        alive, result = False, None
        for p in self.procs:
            ret = p.proc.poll()
            if ret is None:
                alive = True
            elif ret != 0:
                logger.error("rank %s exited with code %s", p.rank, ret)
                result = ret
        if result is not None:
            return result
        return None if alive else 0

    def launch(self):
        raise NotImplementedError

    def stop(self):
        raise NotImplementedError

    def watch(self):
        raise NotImplementedError


class ProcessLauncher(LauncherInterface):
    """Starts ``args.training_script`` once on this node with the environment the manager prepared."""

    def launch(self):
        cmd = [sys.executable, "-u", self.args.training_script] + list(getattr(self.args, "training_script_args", []))
        env = dict(os.environ)
        self.procs = [_Proc(subprocess.Popen(cmd, env=env), int(env.get("PADDLE_TRAINER_ID", "0")))]

    def stop(self):
        self._terminate_procs(timeout=10)

    def watch(self):
        return self._check_procs()


class ElasticManager:
    def __init__(self, args):
        self.args = args
        server = getattr(args, "elastic_server", None) or os.getenv("PADDLE_ELASTIC_SERVER")
        name = getattr(args, "job_id", None) or os.getenv("PADDLE_ELASTIC_JOB_ID")
        np_ = int(getattr(args, "np", None) or os.getenv("PADDLE_ELASTIC_NP", 0) or 0)
        host = getattr(args, "host", None) or os.getenv("POD_IP")
        scale = int(getattr(args, "scale", None) or os.getenv("PADDLE_ELASTIC_SCALE", 0) or 0)
        force = getattr(args, "force", None) or os.getenv("PADDLE_ELASTIC_FORCE")
        self.endpoints = os.getenv("DISTRIBUTED_TRAINER_ENDPOINTS", "")
        self.trainers = os.getenv("PADDLE_TRAINERS", "")
        self.elastic_level = int(os.getenv("PADDLE_ELASTIC_FAULT_TOLERANC_LEVEL", 1))
        self.poll_s = float(os.getenv("EDL_POLL_INTERVAL", 3))
        self.node_ttl = float(os.getenv("EDL_ETCD_TTL", 15))
        self.hosts = []
        self.stopped = False
        self.sigint = 0
        self.launcher = None
        self.job_done = False
        if not server or ":" not in server or not name or not np_:
            logger.info("Elastic is not enabled with server %s name %s and np %s", server, name, np_)
            self.enable = False
            return
        self.enable = True
        self.etcd = KVClient(server.split(","))
        self.etcd.connect()
        self.host = host if host else self._get_host()
        self.prefix = "/paddle/" + name
        self.node_prefix = self.prefix + "/nodes/"
        self.np_path = self.prefix + "/np"
        self.endpoints_path = self.prefix + "/endpoints"
        self.host_path = "{}{:.6f}-{}".format(self.node_prefix, time.time(), os.getpid())
        self.np = np_ + scale
        self.etcd.put(self.prefix, b"0")
        # node registration under a lease: a crashed node vanishes after node_ttl
        self._lease = self.etcd.lease(self.node_ttl)
        self.etcd.put(self.host_path, self.host, self._lease.id)
        self._keep = threading.Thread(target=self._keepalive, daemon=True, name="liveft-keepalive")
        self._keep.start()

        def host_call_back(events, rev):
            if self.stopped:
                return
            if any(e["type"] == "delete" and e["key"] == self.host_path for e in events):
                logger.info("register host again %s", self.host)
                try:
                    self._lease = self.etcd.lease(self.node_ttl)
                    self.etcd.put(self.host_path, self.host, self._lease.id)
                except Exception as e:  # noqa: BLE001
                    logger.warning("re-register failed: %s", e)

        host_watch = self.etcd.add_watch_prefix_callback(self.host_path, host_call_back)
        value, _ = self.etcd.get(self.np_path)
        inp = int(value or 0)
        if scale == 0 and not force:
            assert inp == np_ or inp == 0, "np {} is not consistent with np in the store {}".format(np_, inp)
        else:
            assert inp == np_ or inp == self.np or inp == 0 or force, \
                "np {} scale to {} by {} is not allowed".format(inp, self.np, scale)
        self.etcd.put(self.np_path, "%d" % self.np)

        def np_call_back(events, rev):
            v, _ = self.etcd.get(self.np_path)
            if v is not None and int(v) != self.np:
                logger.info("scale np %d to %d", self.np, int(v))
                self.np = int(v)

        np_watch = self.etcd.add_watch_prefix_callback(self.np_path, np_call_back)
        self.etcd.put(self.endpoints_path, "{}|{}".format(self.endpoints, self.trainers))

        def endpoints_call_back(events, rev):
            if not self.endpoints:
                return
            v, _ = self.etcd.get(self.endpoints_path)
            self.endpoints, self.trainers = (v or b"|").decode().split("|")

        endpoints_watch = self.etcd.add_watch_prefix_callback(self.endpoints_path, endpoints_call_back)
        self.watches = [host_watch, np_watch, endpoints_watch]

    def _keepalive(self):
        while not self.stopped:
            try:
                if self._lease.refresh() <= 0:
                    self._lease = self.etcd.lease(self.node_ttl)
                    self.etcd.put(self.host_path, self.host, self._lease.id)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(max(0.05, self.node_ttl / 3.0))

    def exit(self, completed=False):
        logger.info("manager exit, completed=%s", completed)
        if self.launcher is not None:
            self.launcher.stop()
        if not self.enable:
            return
        if completed:
            self.etcd.put(self.prefix, b"1")
        for w in self.watches:
            self.etcd.cancel_watch(w)
        self.stopped = True
        try:
            self._lease.revoke()
        except Exception:  # noqa: BLE001
            self.etcd.delete(self.host_path)
        kvs, _ = self.etcd.get_prefix(self.node_prefix)
        if len(kvs) == 0 and completed:
            self.etcd.delete_prefix(self.prefix + "/")

    def _get_host(self):
        try:
            return socket.gethostbyname(socket.getfqdn(socket.gethostname()))
        except OSError:
            return "127.0.0.1"

    def _completed(self):
        if not self.enable:
            return True
        v, _ = self.etcd.get(self.prefix)
        return v is not None and int(v) == 1

    def _match(self):
        kvs, _ = self.etcd.get_prefix(self.node_prefix)
        self._node_keys = [kv["key"] for kv in kvs]
        self.hosts = [kv["value"].decode() for kv in kvs]
        if len(self.hosts) != self.np and self.hosts:
            self._adapt_np(len(self.hosts))
        return len(self.hosts) == self.np

    def _adapt_np(self, live: int):
        """Fault-tolerance levels (doc/edl_live_fault_tolerance.md:81-96): 1 = wait for a replacement node
        (np never changes), 2 = shrink to the surviving nodes, 3 = fully elastic (shrink and grow).  The new np
        is published in the store so that every node takes the same decision; a node only shrinks after the
        membership has been stable for one lease TTL (a late starter is not a lost node)."""
        if self.elastic_level < 2 or live == self.np:
            return
        if live > self.np and self.elastic_level < 3:
            return
        now = time.time()
        if getattr(self, "_np_candidate", None) != live:
            self._np_candidate, self._np_since = live, now
            return
        if now - self._np_since < self.node_ttl:
            return
        logger.info("fault-tolerance level %d: np %d -> %d", self.elastic_level, self.np, live)
        self.np = live
        self.etcd.put(self.np_path, "%d" % live)
        self._np_candidate = None

    def _update_hosts(self):
        assert len(self.hosts) != 0, "hosts empty"
        if self.endpoints and self.host in self.endpoints:
            os.environ["DISTRIBUTED_TRAINER_ENDPOINTS"] = self.endpoints
            os.environ["PADDLE_TRAINERS"] = self.trainers
            return
        # rank = position of my registration key (several nodes may share one host ip in tests)
        idx = self._node_keys.index(self.host_path) if self.host_path in self._node_keys else self.hosts.index(self.host)
        os.environ["PADDLE_TRAINER_ID"] = "{}".format(idx)
        hosts = ",".join(self.hosts)
        self.args.ips = hosts
        os.environ["PADDLE_TRAINERS"] = hosts
        os.environ["PADDLE_TRAINERS_NUM"] = str(len(self.hosts))

    def wait(self):
        if not self.enable:
            return
        while not self.stopped:
            if self._completed():
                # the others already finished the job while this node was still waiting
                self.job_done = True
                return
            if self._match():
                logger.info("ready with hosts %s", self.hosts)
                self._update_hosts()
                return
            logger.info("not ready for np %d with hosts %s", self.np, self.hosts)
            time.sleep(self.poll_s)

    def run(self, launcher):
        if self.stopped:
            return
        self.launcher = launcher(self.args)
        self._np_at_launch = self.np
        self.launcher.launch()

    def watch(self):
        while not self.stopped:
            ret = self.launcher.watch()
            if ret is not None:
                logger.info("job exit with code %s", ret)
                completed = ret == 0
                self.exit(completed=completed)
                if completed:
                    return ElasticStatus.COMPLETED
                return ElasticStatus.RESTART if self.elastic_level == 1 else ElasticStatus.ERROR
            if self.enable and not self._completed() and (not self._match() or self.np != self._np_at_launch):
                # membership differs from np, or np itself changed (scheduler resize / level-2 shrink):
                # the running trainers have a stale world size
                self.launcher.stop()
                return ElasticStatus.HOLD
            time.sleep(self.poll_s)
        return ElasticStatus.EXIT

    def signal_handler(self, sigint, frame):
        if self.enable:
            self.exit()
        self.sigint = sigint
        self.stopped = True
