"""In-place elastic rescale: surviving trainers keep their process (CUDA context, parameters, optimizer state,
data-loader workers) across a membership change instead of being killed and restarted.

The reference only knows stop-resume: on every stage change each launcher SIGTERMs its trainers, starts new
ones and they reload the checkpoint (python/edl/utils/launcher.py:221-244; doc/edl_collective_design_doc.md).
Its design document lists "no restart" as future work.  Here, with ``EDL_RESCALE_MODE=inplace`` (or
``--rescale_mode inplace`` on the launcher):

* every trainer owns an :class:`ElasticContext`.  It watches the job's cluster record in the store; once per
  ``check_every`` steps ``poll()`` folds the local "membership changed" flag into a 1-element MAX all-reduce, so
  ALL ranks leave the old process group at the same step boundary;
* ``rescale()`` then runs the *stage rendezvous* through the store -- every member of the new stage (survivors
  and freshly started trainers alike) publishes ``ready/<stage>/<rank> = survivor|joiner``; the first one to see
  all of them writes the stage's commit record in ONE transaction that re-checks every ready key, and a member
  that wants to move on to a newer stage may withdraw its key only in a transaction that checks the commit
  record is absent -- so "everybody proceeds with stage S" and "somebody abandoned S" are mutually exclusive;
* the stage's communication is bootstrapped through the same store (:class:`KVRendezvousStore`, a
  ``torch.distributed.Store`` on the job's KV store, one key prefix per stage -- no free port, no TCPStore that
  would die with rank 0; the reference re-broadcasts an ncclUniqueId over TCP among the new endpoints,
  utils/train_process.py:37-41).  On GPUs (backend ``"fabric"``, the default there) NO process group and no NCCL
  communicator exists at all: the trainers get a :class:`edl_b200.parallel.symm.Fabric` (store + rank + world) and
  the data-parallel engine maps a fresh symmetric slab through it (cuMem VMM handles, csrc/vmm.cpp); agreement,
  state hand-off and gradient reduction all run on our own kernels, which time out into an error word instead of
  hanging when a peer dies.  On CPU (backend ``"gloo"``) the store bootstraps a gloo group;
* ``StageInfo.root`` names the lowest-ranked survivor: joiners take parameters / optimizer state / the epoch
  cursor from it over the fabric (``ElasticDataParallel.broadcast_parameters``) instead of reading the checkpoint;
  ``root is None`` means nobody survived (cold start or stop-resume fallback) and the checkpoint is the source.

* a collective that FAILS because a peer died is the same thing triggered differently: ``recover()`` drops the
  broken group, waits for the store to publish the new stage and rejoins as a survivor (hot recovery; the
  reference restarts every trainer of the job and reloads the checkpoint).  If the membership does NOT change -- a
  false alarm: a stalled rank, a transient transport error -- the same members re-form the same stage under a fresh
  namespace (``StageInfo.generation``) and continue from rank 0's state (soft reset).

The launcher side (utils/launcher.py) leaves the trainers of a surviving pod alone when they have announced an
ElasticContext and acknowledges the switch by waiting for their ready keys; anything else -- a trainer that does
not answer in time, a trainer that died -- falls back to the reference's stop-resume for that pod, and the
restarted trainers simply show up as joiners of the same stage.
"""
from __future__ import annotations

import datetime
import json
import os
import threading
import time
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .utils import cluster as edl_cluster
from .utils import constants
from .utils.env import TrainerEnv
from .utils.log_utils import logger

INPLACE_TABLE = "inplace"


def inplace_requested(environ=None) -> bool:
    e = environ if environ is not None else os.environ
    return e.get("EDL_RESCALE_MODE", "").lower() == "inplace"


def _prefix(job_id: str) -> str:
    return "/%s/%s/" % (job_id, INPLACE_TABLE)


def capable_key(job_id: str, pod_id: str, rank_in_pod: int) -> str:
    return "%scapable/%s/%d" % (_prefix(job_id), pod_id, rank_in_pod)


def ready_key(job_id: str, stage: str, rank: int) -> str:
    return "%sready/%s/%d" % (_prefix(job_id), stage, rank)


def commit_key(job_id: str, stage: str) -> str:
    return "%scommit/%s" % (_prefix(job_id), stage)


class EdlEvicted(Exception):
    """This trainer's pod is not part of the new stage (scale-in): leave quietly with exit code 0."""


class KVRendezvousStore(dist.Store):
    """``torch.distributed.Store`` on the job's KV store: process-group bootstrap without a TCPStore."""

    def __init__(self, kv, prefix: str, timeout_s: float = 120.0):
        super().__init__()
        self._kv, self._p, self._t = kv, prefix, timeout_s

    def set(self, key, value):
        self._kv.put(self._p + key, value if isinstance(value, (bytes, bytearray)) else str(value).encode())

    def get(self, key):
        deadline = time.time() + self._t
        delay = 0.002
        while True:
            v, _ = self._kv.get(self._p + key)
            if v is not None:
                return v
            if time.time() > deadline:
                raise RuntimeError("rendezvous store: timed out waiting for key %r" % key)
            time.sleep(delay)
            delay = min(0.05, delay * 1.5)

    def add(self, key, amount):
        k = self._p + key
        while True:
            v, meta = self._kv.get(k)
            cur = int(v) if v is not None else 0
            cmp_ = [{"key": k, "target": "version", "op": "==", "value": meta["version"] if meta else 0}]
            ok, _ = self._kv.txn(cmp_, [{"op": "put", "key": k, "value": str(cur + int(amount))}])
            if ok:
                return cur + int(amount)

    def compare_set(self, key, expected, desired):
        k = self._p + key
        exp = expected if isinstance(expected, (bytes, bytearray)) else str(expected).encode()
        des = desired if isinstance(desired, (bytes, bytearray)) else str(desired).encode()
        v, meta = self._kv.get(k)
        if v is None:
            if len(exp) == 0:
                cmp_ = [{"key": k, "target": "version", "op": "==", "value": 0}]
                ok, _ = self._kv.txn(cmp_, [{"op": "put", "key": k, "value": des}])
                return des if ok else self.get(key)
            return exp
        if v == exp:
            ok, _ = self._kv.txn([{"key": k, "value": exp}], [{"op": "put", "key": k, "value": des}])
            return des if ok else self.get(key)
        return v

    def wait(self, keys, timeout=None):
        old = self._t
        if timeout is not None:
            self._t = timeout.total_seconds() if hasattr(timeout, "total_seconds") else float(timeout)
        try:
            for k in keys:
                self.get(k)
        finally:
            self._t = old

    def check(self, keys):
        return all(self._kv.get(self._p + k)[0] is not None for k in keys)

    def delete_key(self, key):
        return self._kv.delete(self._p + key) > 0

    def num_keys(self):
        return len(self._kv.get_prefix(self._p)[0])

    def set_timeout(self, timeout):
        self._t = timeout.total_seconds() if hasattr(timeout, "total_seconds") else float(timeout)

    # extended API (symmetric-memory and NCCL bootstrap use it when the store advertises it)
    def has_extended_api(self):
        return True

    def append(self, key, value):
        k = self._p + key
        add = value if isinstance(value, (bytes, bytearray)) else str(value).encode()
        while True:
            v, meta = self._kv.get(k)
            cmp_ = [{"key": k, "target": "version", "op": "==", "value": meta["version"] if meta else 0}]
            ok, _ = self._kv.txn(cmp_, [{"op": "put", "key": k, "value": (v or b"") + bytes(add)}])
            if ok:
                return

    def multi_get(self, keys):
        return [self.get(k) for k in keys]

    def multi_set(self, keys, values):
        ops = [{"op": "put", "key": self._p + k, "value": v if isinstance(v, (bytes, bytearray)) else str(v).encode()}
               for k, v in zip(keys, values)]
        self._kv.txn([], ops)


@dataclass
class StageInfo:
    stage: str
    rank: int
    size: int
    rank_in_pod: int
    root: Optional[int]          # rank to take the training state from; None = load the checkpoint
    survivor: bool               # this trainer carried its state over from the previous stage
    prev_size: int               # world size this trainer ran with before (== size on a cold start)
    rendezvous_s: float = 0.0    # seconds from leaving the old process group to having the new one
    generation: int = 0          # soft resets of this stage so far (same members, fresh group; see recover())

    @property
    def group_name(self) -> str:
        """Namespace of this stage's process group / fabric in the store: unique per (stage, generation)."""
        return self.stage if self.generation == 0 else "%s~%d" % (self.stage, self.generation)


class ElasticContext:
    def __init__(self, backend: Optional[str] = None, check_every: int = 10, timeout_s: float = 120.0,
                 environ=None, etcd=None):
        self.env = TrainerEnv(environ)
        self.backend = backend or os.environ.get("EDL_INPLACE_BACKEND") or (
            "fabric" if torch.cuda.is_available() else "gloo")
        self.fabric = None              # backend "fabric": what StudentTrainer / ElasticDataParallel take instead of a group
        self._bcast_seq = 0
        self.check_every = max(1, int(check_every))
        self.timeout_s = timeout_s
        self.info: Optional[StageInfo] = None
        self._steps = 0
        self._changed = threading.Event()
        self._stage_pods = set()
        self._stage_pod_list = []
        self._watch_id = None
        self._etcd = etcd
        self._own_etcd = etcd is None
        self.standalone = not self.env.etcd_endpoints and etcd is None
        if not self.standalone and etcd is None:
            from .discovery.etcd_client import EtcdClient

            self._etcd = EtcdClient(self.env.etcd_endpoints, root=self.env.job_id)
            self._etcd.init()

    # ------------------------------------------------------------------ helpers
    @property
    def kv(self):
        return self._etcd.kv

    def _locate(self, cluster):
        """(rank, size, rank_in_pod) of this trainer in ``cluster`` or None if its pod is not listed."""
        pod = cluster.get_pod_by_id(self.env.pod_id)
        if pod is None:
            return None
        for t in pod.trainers:
            if t.rank_in_pod == self.env.rank_in_pod:
                return t.global_rank, cluster.get_trainers_nranks(), t.rank_in_pod
        return None

    def _on_cluster_event(self, add, rm):
        try:
            c = edl_cluster.load_from_etcd(self._etcd, timeout=5)
        except Exception:  # noqa: BLE001 - poll() will look again
            return
        if c is not None and self.info is not None and c.stage != self.info.stage:
            self._changed.set()

    def _joiners_ready(self) -> bool:
        """Survivors keep training in the old stage until every trainer of a NEWLY ADDED pod is waiting at the new
        stage's rendezvous (its interpreter start-up and imports cost seconds; nobody should idle for them).
        A scale-in has no joiners and switches at once."""
        try:
            cluster = edl_cluster.load_from_etcd(self._etcd, timeout=5)
            if cluster is None:
                return False
            want = [ready_key(self.env.job_id, cluster.stage, t.global_rank) for p in cluster.pods
                    if p.id not in self._stage_pods for t in p.trainers]
            if not want:
                return True
            kvs, _ = self.kv.get_prefix("%sready/%s/" % (_prefix(self.env.job_id), cluster.stage))
            have = {kv["key"] for kv in kvs}
            return all(k in have for k in want)
        except Exception:  # noqa: BLE001 - ask again at the next poll
            return False

    def _rendezvous(self, survivor: bool, prev_size: int, soft: Optional[Tuple[str, int]] = None) -> StageInfo:
        """Stage rendezvous through the store (see module docstring); returns once the stage is committed.
        ``soft = (stage, generation)``: re-form THAT stage under a fresh key namespace (soft reset after a collective
        failed although nobody left); if the membership moves on meanwhile, the newer stage is joined as usual."""
        job = self.env.job_id
        t0 = time.time()
        deadline = t0 + self.timeout_s
        my_key = None
        while True:
            cluster = edl_cluster.load_from_etcd(self._etcd, timeout=10)
            loc = self._locate(cluster) if cluster is not None else None
            if cluster is not None and loc is None:
                raise EdlEvicted("pod %s is not part of stage %s" % (self.env.pod_id, cluster.stage))
            if cluster is None:
                if time.time() > deadline:
                    raise TimeoutError("no cluster record in the store")
                time.sleep(0.1)
                continue
            rank, size, rip = loc
            stage = cluster.stage
            gen = soft[1] if soft is not None and soft[0] == stage else 0
            skey = stage if gen == 0 else "%s~%d" % (stage, gen)          # key namespace of this (stage, generation)
            my_key = ready_key(job, skey, rank)
            self.kv.put(my_key, b"survivor" if survivor else b"joiner")
            keys = [ready_key(job, skey, r) for r in range(size)]
            ckey = commit_key(job, skey)
            while True:
                committed, _ = self.kv.get(ckey)
                if committed is None:
                    kvs, _ = self.kv.get_prefix("%sready/%s/" % (_prefix(job), skey))
                    have = {kv["key"]: kv["value"] for kv in kvs}
                    if all(k in have for k in keys):
                        flags = [have[k] for k in keys]
                        survivors = [r for r, f in enumerate(flags) if f == b"survivor"]
                        record = json.dumps({"size": size, "root": survivors[0] if survivors else None}).encode()
                        cmp_ = [{"key": k, "target": "version", "op": ">", "value": 0} for k in keys]
                        cmp_.append({"key": ckey, "target": "version", "op": "==", "value": 0})
                        self.kv.txn(cmp_, [{"op": "put", "key": ckey, "value": record}])
                        committed, _ = self.kv.get(ckey)
                if committed is not None:
                    rec = json.loads(committed.decode())
                    self._stage_pods = cluster.get_pods_ids_set()
                    self._stage_pod_list = cluster.get_pods_ids_list()
                    return StageInfo(stage=stage, rank=rank, size=size, rank_in_pod=rip, root=rec["root"],
                                     survivor=survivor, prev_size=prev_size, rendezvous_s=time.time() - t0,
                                     generation=gen)
                latest = edl_cluster.load_from_etcd(self._etcd, timeout=10)
                if latest is not None and latest.stage != stage:
                    # the membership moved on while this stage was still forming: withdraw -- unless the stage
                    # got committed in the meantime, in which case everybody (including us) runs it first
                    cmp_ = [{"key": ckey, "target": "version", "op": "==", "value": 0}]
                    ok, _ = self.kv.txn(cmp_, [{"op": "delete", "key": my_key}])
                    if ok:
                        logger.info("stage %s superseded by %s before it formed", stage, latest.stage)
                        break
                    continue
                if time.time() > deadline:
                    raise TimeoutError("stage %s did not form within %.0fs (%d trainers expected)" % (
                        stage, self.timeout_s, size))
                time.sleep(0.05)

    def _init_group(self, info: StageInfo):
        self._bcast_seq = 0
        if self.backend == "fabric":
            from .parallel.symm import Fabric

            dev = torch.device("cuda", info.rank_in_pod % max(1, torch.cuda.device_count()))
            torch.cuda.set_device(dev)
            store = KVRendezvousStore(self.kv, "%spg/%s/" % (_prefix(self.env.job_id), info.group_name), self.timeout_s)
            self.fabric = Fabric(store=store, rank=info.rank, world=info.size, tag="stage-%s" % info.group_name)
            return
        if info.size <= 1:
            return
        store = KVRendezvousStore(self.kv, "%spg/%s/" % (_prefix(self.env.job_id), info.group_name), self.timeout_s)
        kwargs = {}
        if self.backend == "nccl":
            dev = torch.device("cuda", info.rank_in_pod % max(1, torch.cuda.device_count()))
            torch.cuda.set_device(dev)
            kwargs["device_id"] = dev
        dist.init_process_group(self.backend, store=store, rank=info.rank, world_size=info.size,
                                timeout=datetime.timedelta(seconds=self.timeout_s), **kwargs)

    # ------------------------------------------------------------------ public API
    def start(self) -> StageInfo:
        """Join the stage the launcher started this trainer for (or whatever the newest stage is by now)."""
        if self.standalone:
            self.info = StageInfo("standalone", 0, 1, 0, None, False, 1)
            return self.info
        self.kv.put(capable_key(self.env.job_id, self.env.pod_id, self.env.rank_in_pod), str(os.getpid()).encode())
        info = self._rendezvous(survivor=False, prev_size=max(1, self.env.size))
        info.prev_size = info.size
        self._init_group(info)
        self.info = info
        self._steps = 0                 # poll() cadence restarts with every stage: identical on all ranks
        self._watch_id = self._etcd.watch_service(constants.ETCD_CLUSTER, self._on_cluster_event)
        self._on_cluster_event(None, None)        # a change may have landed between rendezvous and watch
        logger.info("trainer joined stage %s as rank %d/%d (state from %s)", info.stage, info.rank, info.size,
                    "checkpoint" if info.root is None else "rank %d" % info.root)
        return info

    def poll(self, force: bool = False, agree=None) -> bool:
        """Call once per training step.  True on EVERY rank of the current stage at the same step as soon as any
        of them has seen the membership change (the decision rides on a 1-element MAX all-reduce every
        ``check_every`` steps).  ``agree`` (e.g. ``ElasticDataParallel.agree``) replaces the library all-reduce by
        a collective with a timeout: ``flag -> (max flag, error word)``; a non-zero error word raises
        ``RuntimeError`` so that the caller's hot-recovery path (``recover()``) takes over."""
        if self.standalone or self.info is None:
            return False
        self._steps += 1
        if not force and self._steps % self.check_every != 0:
            return False
        flag = 1.0 if (self._changed.is_set() and self._joiners_ready()) else 0.0
        if agree is not None and self.info.size > 1:
            flag, err = agree(flag)
            if err:
                raise RuntimeError("collective timed out waiting for peer %d (dead pod?)" % (err - 1))
            return flag > 0.5
        if self.info.size > 1 and self.backend == "fabric":
            flag = max(float(v) for v in self.allgather_object(flag))        # no agree kernel given: through the store
        elif self.info.size > 1 and dist.is_initialized():
            dev = torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu")
            t = torch.tensor([flag], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            flag = float(t.item())
        return flag > 0.5

    def should_switch(self) -> bool:
        """Local (non-collective) form of ``poll()``: the membership changed and every joiner is waiting.  For loops
        whose ranks do not run in lock step (e.g. the elastic data reader, where pods consume different numbers of
        batches); each rank then calls ``rescale()`` on its own and the stage rendezvous brings them together."""
        return (not self.standalone) and self._changed.is_set() and self._joiners_ready()

    @property
    def pod_ids(self):
        """Pod ids of the committed stage, leader first (what ``collective.distribute_reader.Reader`` wants)."""
        c = edl_cluster.load_from_etcd(self._etcd, timeout=10)
        if c is not None and self.info is not None and c.stage == self.info.stage:
            return c.get_pods_ids_list()
        return list(self._stage_pod_list)

    @property
    def etcd(self):
        return self._etcd

    def allgather_object(self, obj):
        """Every rank's ``obj`` (rank order) over the current stage: through the process group, or -- fabric backend --
        through the store (JSON values: cursors, counters; this is control traffic, not tensors)."""
        if self.info is None or self.info.size <= 1:
            return [obj]
        if self.backend == "fabric":
            self._bcast_seq += 1
            st = self.fabric.store
            st.set("obj/%d/%d" % (self._bcast_seq, self.info.rank), json.dumps(obj).encode())
            return [json.loads(bytes(st.get("obj/%d/%d" % (self._bcast_seq, r))).decode()) for r in range(self.info.size)]
        if not dist.is_initialized():
            return [obj]
        out = [None] * self.info.size
        dist.all_gather_object(out, obj)
        return out

    def broadcast_object(self, obj, root: int = 0):
        """``root``'s ``obj`` on every rank of the current stage (the epoch / step cursor after a state hand-off)."""
        if self.info is None or self.info.size <= 1:
            return obj
        if self.backend == "fabric":
            self._bcast_seq += 1
            st = self.fabric.store
            key = "bcast/%d" % self._bcast_seq
            if self.info.rank == root:
                st.set(key, json.dumps(obj).encode())
                return obj
            return json.loads(bytes(st.get(key)).decode())
        box = [obj]
        dist.broadcast_object_list(box, src=root)
        return box[0]

    def barrier(self):
        """Host barrier of the current stage's trainers."""
        if self.info is None or self.info.size <= 1:
            return
        if self.backend == "fabric":
            self.allgather_object(0)
        elif dist.is_initialized():
            dist.barrier()

    def rescale(self) -> StageInfo:
        """Leave the old process group, run the stage rendezvous, build the new group.  Raises
        :class:`EdlEvicted` when this trainer's pod is not part of the new stage."""
        old = self.info
        t0 = time.time()
        if dist.is_initialized():
            dist.destroy_process_group()
        self._changed.clear()
        info = self._rendezvous(survivor=True, prev_size=old.size)
        self._init_group(info)
        info.rendezvous_s = time.time() - t0          # old group torn down -> new group usable
        self.info = info
        self._steps = 0
        self._on_cluster_event(None, None)
        logger.info("trainer moved in place from stage %s (%d ranks) to %s as rank %d/%d in %.2fs", old.stage,
                    old.size, info.stage, info.rank, info.size, info.rendezvous_s)
        return info

    def recover(self, wait_s: Optional[float] = None) -> StageInfo:
        """Hot recovery after a FAILED collective (a peer died mid-step; gloo raises, our all-reduce kernels time
        out into ``check_comm_error()``): abandon the broken group at once -- closing its sockets is what makes the
        other survivors' collectives fail fast too --, wait until the store has noticed the loss (the dead pod's
        lease expires, the leader publishes a new stage), then rejoin as a survivor exactly like ``rescale()``.
        The caller discards the interrupted step and takes ``StageInfo.root``'s state: every survivor ends up on the
        root's last completed optimizer step, whatever each of them had applied of the broken one."""
        old = self.info
        try:
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception as e:  # noqa: BLE001 - the group is broken anyway
            logger.warning("destroying the broken process group: %s", e)
        # Nobody may have died: a rank that stalled longer than the communication time-out or a transient transport error
        # fails a collective just the same, and every member of the stage lands here (the abandoned group makes the
        # others' collectives fail too).  So the members first try to re-form the SAME stage under a fresh namespace
        # (generation + 1) -- a soft reset.  If everybody is alive that rendezvous completes at once and they continue
        # from rank 0's state; if a pod did die, its ready key never appears, the store publishes the smaller stage
        # (lease expiry + one leader poll) and ``_rendezvous`` withdraws and joins THAT stage as a survivor: the hot
        # recovery.  EDL_SOFT_RESET=0: wait for a membership change first and give up after ``wait_s`` without one.
        soft = (old.stage, old.generation + 1)
        if os.environ.get("EDL_SOFT_RESET", "1") == "0":
            soft = None
            wait_s = wait_s if wait_s is not None else constants.ETCD_TTL * 2 + 4 * constants.POLL_INTERVAL + 10
            deadline = time.time() + wait_s
            while time.time() < deadline:
                c = edl_cluster.load_from_etcd(self._etcd, timeout=10)
                if c is not None and c.stage != old.stage:
                    break
                time.sleep(0.1)
            else:
                raise TimeoutError("a collective failed but the membership did not change within %.0fs" % wait_s)
        self._changed.clear()
        t0 = time.time()
        info = self._rendezvous(survivor=True, prev_size=old.size, soft=soft)
        self._init_group(info)
        info.rendezvous_s = time.time() - t0          # new stage published -> new group usable
        self.info = info
        self._steps = 0
        self._on_cluster_event(None, None)
        logger.info("trainer recovered in place from stage %s (%d ranks) to %s as rank %d/%d%s", old.stage, old.size,
                    info.group_name, info.rank, info.size,
                    " (soft reset: the membership did not change)" if info.generation > 0 else "")
        return info

    def close(self):
        if self._watch_id is not None:
            try:
                self._etcd.cancel_watch(self._watch_id)
            except Exception:  # noqa: BLE001
                pass
        if dist.is_initialized():
            dist.destroy_process_group()
        if self._own_etcd and self._etcd is not None:
            self._etcd.close()
