"""The elastic launcher: rendezvous -> barrier -> spawn trainers -> supervise -> re-barrier and
restart on membership change -> publish the final status
(reference: python/edl/utils/launcher.py:33-261, call stack in SURVEY 3.1)."""
import time

from . import cluster as edl_cluster
from . import (cluster_watcher, constants, exceptions, leader_pod, pod_server, pod_server_client,
               resource_pods, status as edl_status, train_process)
from .log_utils import logger


class Launcher:
    def __init__(self, job_env, pod, etcd, args):
        self._job_env, self._pod, self._etcd, self._args = job_env, pod, etcd, args
        self._pod_server = None
        self._resource_register = None
        self._leader_register = None
        self._watcher = None
        self._procs = []
        self._cluster = None
        self._metrics = None
        self._leave = False         # SIGTERM / request_leave(): announce the departure, let the job re-plan first
        self._leaving_since = None
        self.rescales = 0          # number of stage changes survived
        self.last_rescale_s = None  # wall seconds from "change seen" to "trainers restarted"

    # ------------------------------------------------------------------ set-up
    def _publish(self):
        """Prometheus text metrics of this pod's launcher (EDL_METRICS_PORT): what a scheduler / dashboard needs to see
        elastic events -- the reference only has the status tables in etcd."""
        if self._metrics is None:
            return
        labels = {"job": str(self._job_env.job_id), "pod": str(self._pod.id)[:8]}
        m = self._metrics
        m.set("edl_launcher_rescales_total", self.rescales, labels)
        m.set("edl_launcher_inplace_rescales_total", getattr(self, "inplace_rescales", 0), labels)
        m.set("edl_launcher_last_rescale_seconds", self.last_rescale_s or 0.0, labels)
        m.set("edl_launcher_world_size", self._cluster.get_trainers_nranks() if self._cluster is not None else 0, labels)
        m.set("edl_launcher_pods", len(self._cluster.pods) if self._cluster is not None else 0, labels)
        m.set("edl_launcher_is_leader", 1 if (self._leader_register is not None and self._leader_register.is_leader()) else 0,
              labels)
        m.set("edl_launcher_leaving", 1 if self._leaving_since is not None else 0, labels)
        m.set("edl_launcher_trainers_alive", sum(1 for tp in self._procs if tp.proc.poll() is None), labels)

    def init(self):
        import os

        self._metrics = None
        port = os.environ.get("EDL_METRICS_PORT")
        if port:
            from .metrics import MetricsExporter

            try:
                self._metrics = MetricsExporter(int(port)).start()
                logger.info("launcher metrics on port %d", self._metrics.port)
            except OSError as e:
                logger.warning("metrics endpoint not started: %s", e)
        self._pod_server = pod_server.PodServer(self._job_env, self._pod.id, etcd=self._etcd).start()
        self._pod.port = self._pod_server.port
        edl_status.save_pod_status_to_etcd(self._etcd, self._pod.id, edl_status.Status.INITIAL, timeout=30)
        return self

    def _barrier(self, timeout):
        """Wait until every pod of the current stage has arrived at the leader; returns the cluster."""
        begin = time.time()
        last_err = None
        while time.time() - begin < timeout:
            try:
                leader = leader_pod.load_from_etcd(self._etcd, timeout=3)
                c = pod_server_client.Client(leader.endpoint)
                try:
                    return c.barrier(self._job_env.job_id, self._pod.id,
                                     timeout=min(15.0, max(1.0, timeout - (time.time() - begin))))
                finally:
                    c.close()
            except exceptions.EdlException as e:
                last_err = e
                # evicted (scale-in / cluster full): the stored cluster of a NEW stage does not list this pod --
                # hand it to the caller, whose _adopt() then lets the pod leave quietly instead of timing out
                try:
                    cur = edl_cluster.load_from_etcd(self._etcd, timeout=3)
                except exceptions.EdlException:
                    cur = None
                if (self._cluster is not None and cur is not None and cur.stage != self._cluster.stage
                        and cur.get_pod_by_id(self._pod.id) is None):       # (a joiner keeps waiting to be appended)
                    return cur
                time.sleep(min(1.0, constants.POLL_INTERVAL))
        raise exceptions.EdlBarrierError("barrier did not complete in {}s: {}".format(timeout, last_err))

    def request_leave(self):
        """Ask the pod to leave the job gracefully (called from the SIGTERM handler of collective/launch.py)."""
        self._leave = True

    def _adopt(self, cluster):
        """Take this pod's record (rank, global trainer ranks) from the agreed cluster; False if evicted."""
        mine = cluster.get_pod_by_id(self._pod.id)
        if mine is None:
            return False
        mine.port = self._pod.port
        self._pod = mine
        self._cluster = cluster
        return True

    # ------------------------------------------------------------------ main loop
    def launch(self):
        ok = False
        try:
            ok = self._launch()
        except Exception:
            logger.exception("launcher of pod %s failed", self._pod.id)
            ok = False
        finally:
            self._exit(ok)
        return ok

    def _launch(self):
        je = self._job_env
        self._resource_register = resource_pods.Register(je, self._pod.id, self._pod.to_json(), etcd=self._etcd)
        self._leader_register = leader_pod.Register(je, self._pod.id, etcd=self._etcd)
        cluster = self._barrier(constants.BARRIER_TIMEOUT)
        if not self._adopt(cluster):
            logger.info("pod %s was not admitted (cluster is full); exiting quietly", self._pod.id)
            return True
        edl_status.save_pod_status_to_etcd(self._etcd, self._pod.id, edl_status.Status.RUNNING, timeout=30)
        self._watcher = cluster_watcher.Watcher(je, self._cluster, etcd=self._etcd)
        self._procs = train_process.start(je, self._cluster, self._pod, self._args.training_script,
                                          self._args.training_script_args, log_dir=je.log_dir)
        poll = min(constants.POLL_INTERVAL, 1.0)
        while True:
            self._publish()
            alive, failed = train_process.watch(self._procs)
            collateral = False
            if failed is not None:
                # A trainer that dies because a PEER pod died (its collective raised: gloo resets the connection at
                # once, NCCL after its watchdog) is collateral damage of a membership change, not a job failure: give
                # the store the time it needs to notice the dead pod (lease expiry + a generator round); if the
                # cluster changes, fall through to the rescale path below, which restarts this pod's trainers.
                logger.error("a trainer exited with code %s", failed)
                grace = constants.COLLATERAL_GRACE
                deadline = time.time() + grace
                while time.time() < deadline and not self._watcher.changed and self._leaving_since is None:
                    time.sleep(min(poll, 0.2))
                if not self._watcher.changed:
                    train_process.terminate(self._procs)
                    return False
                logger.warning("... while the membership changed: treating the exit as collateral, restarting trainers")
                collateral = True
            if not alive and not collateral:
                logger.info("all trainers of pod %s finished", self._pod.id)
                return True
            if self._leave and self._leaving_since is None:
                # graceful departure (SIGTERM from the scheduler, k8s pod deletion with a grace period): give up the
                # leadership and the resource key FIRST, keep the trainers alive.  The (new) leader publishes a stage
                # without this pod; in-place trainers agree on the switch and the ones of this pod leave by
                # themselves, restart-mode pods re-barrier -- nobody's collective is broken by a vanished peer.
                self._leaving_since = time.time()
                logger.info("pod %s is leaving: releasing leadership and resource registration", self._pod.id)
                self._leader_register.stop()
                self._resource_register.stop()
            if self._leaving_since is not None:
                if time.time() - self._leaving_since > constants.LEAVE_GRACE:
                    logger.warning("the job did not re-plan within %.0fs; stopping trainers", constants.LEAVE_GRACE)
                    train_process.terminate(self._procs)
                    return True
            elif self._resource_register.is_stopped() or self._leader_register.is_stopped():
                logger.error("lost the store registration; stopping trainers")
                train_process.terminate(self._procs)
                return False
            if self._watcher.changed:
                t0 = time.time()
                logger.info("cluster changed; re-barrier")
                new_cluster = self._barrier(constants.RESCALE_BARRIER_TIMEOUT)
                if self._inplace_possible():
                    done = self._rescale_in_place(new_cluster, t0)
                    if done is not None:
                        if done == "evicted":
                            return True
                        time.sleep(poll)
                        continue
                    logger.warning("in-place rescale was not acknowledged; falling back to stop-resume")
                train_process.terminate(self._procs)
                self._watcher.stop()
                if not self._adopt(new_cluster):
                    logger.info("pod %s is not part of the new cluster; exiting", self._pod.id)
                    return True
                self._watcher = cluster_watcher.Watcher(je, self._cluster, etcd=self._etcd)
                self._procs = train_process.start(je, self._cluster, self._pod, self._args.training_script,
                                                  self._args.training_script_args, log_dir=je.log_dir)
                self.rescales += 1
                self.last_rescale_s = time.time() - t0
                logger.info("rescaled to %d trainers in %.2fs", self._cluster.get_trainers_nranks(),
                            self.last_rescale_s)
                self._publish()
            time.sleep(poll)

    # ------------------------------------------------------------------ in-place rescale (edl_b200/elastic.py)
    def _inplace_possible(self):
        """Mode requested, every local trainer alive and owning an ElasticContext (it announced itself)."""
        from .. import elastic

        if not elastic.inplace_requested():
            return False
        try:
            for tp in self._procs:
                if tp.proc.poll() is not None:
                    return False
                key = elastic.capable_key(self._job_env.job_id, self._pod.id, tp.local_rank)
                if self._etcd.kv.get(key)[0] is None:
                    return False
        except Exception as e:  # noqa: BLE001
            logger.debug("in-place capability check failed: %s", e)
            return False
        return bool(self._procs)

    def _rescale_in_place(self, new_cluster, t0):
        """Leave the trainers running; they switch stages themselves.  Returns "ok", "evicted" or None (= fall
        back to stop-resume: some trainer did not enter the new stage in time)."""
        from .. import elastic

        job = self._job_env.job_id
        mine = new_cluster.get_pod_by_id(self._pod.id)
        if mine is None:
            logger.info("pod %s is not part of the new cluster; its trainers leave by themselves", self._pod.id)
            rc = train_process.wait_all(self._procs, timeout=constants.INPLACE_ACK_TIMEOUT)
            if rc == -1:
                train_process.terminate(self._procs)
            self._watcher.stop()
            return "evicted"
        deadline = time.time() + constants.INPLACE_ACK_TIMEOUT
        want = [elastic.ready_key(job, new_cluster.stage, t.global_rank) for t in mine.trainers]
        while time.time() < deadline:
            if any(tp.proc.poll() is not None for tp in self._procs):
                return None
            try:
                if all(self._etcd.kv.get(k)[0] == b"survivor" for k in want):
                    break
                latest = edl_cluster.load_from_etcd(self._etcd, timeout=3)
                if latest is not None and latest.stage != new_cluster.stage:
                    break      # superseded: the watcher below picks the newer stage up right away
            except exceptions.EdlException:
                pass
            time.sleep(0.1)
        else:
            return None
        self._watcher.stop()
        self._adopt(new_cluster)
        self._watcher = cluster_watcher.Watcher(self._job_env, self._cluster, etcd=self._etcd)
        self.rescales += 1
        self.inplace_rescales = getattr(self, "inplace_rescales", 0) + 1
        self.last_rescale_s = time.time() - t0
        logger.info("rescaled IN PLACE to %d trainers in %.2fs (trainer processes kept)",
                    self._cluster.get_trainers_nranks(), self.last_rescale_s)
        return "ok"

    # ------------------------------------------------------------------ tear-down
    def _exit(self, ok):
        try:
            edl_status.save_pod_flag_to_etcd(self._etcd, self._pod.id, ok, timeout=15)
        except Exception:  # noqa: BLE001
            logger.warning("could not persist pod status")
        is_leader = self._leader_register is not None and self._leader_register.is_leader()
        if self._watcher is not None:
            self._watcher.stop()
        if self._resource_register is not None:
            self._resource_register.stop()
        if is_leader:
            # the leader declares the job status after every follower has released its resource key
            released = resource_pods.wait_followers_release(self._etcd, self._pod.id,
                                                            timeout=constants.ETCD_TTL * 2 + 5)
            try:
                _, _, _, failed = edl_status.load_pods_status_from_etcd(self._etcd, timeout=5)
                job_ok = ok and released and not failed
                edl_status.save_job_flag_to_etcd(self._etcd, self._pod.id, job_ok, timeout=15)
            except Exception:  # noqa: BLE001
                logger.warning("could not persist job status")
        if self._leader_register is not None:
            self._leader_register.stop()
        if self._procs:
            train_process.terminate(self._procs)
        if self._pod_server is not None:
            self._pod_server.stop()
        if getattr(self, "_metrics", None) is not None:
            self._metrics.stop()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass
