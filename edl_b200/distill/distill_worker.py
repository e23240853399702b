"""The DistillReader pipeline: slice the student's data stream into teacher-sized tasks, fan them out
to a dynamic pool of teacher connections, and hand the samples back *in order* with the teacher's
predictions appended.

Behavioural contract of the reference (python/edl/distill/distill_worker.py:46-847):
* tasks of ``teacher_batch_size`` samples, at most ``2*N+2`` in flight (back-pressure), results
  re-ordered by task id, original batch boundaries restored for the three reader kinds;
* a manager polls service discovery (<= 2 s) and adds / retires teacher connections while data
  flows; a failing teacher's task is re-queued and served by another teacher;
* the epoch ends only after every task produced by the reader has been delivered, whatever
  happened to individual workers.

Design here: threads instead of forked processes (the hot path is a blocking RPC plus numpy
(de)serialisation, both GIL-free), one worker thread per teacher connection, plain queues -- no
poison-pill counting across processes.  The device-resident NVSwitch path bypasses all of this
(``device_feed.py``).
"""
import logging
import queue
import threading
import time

import numpy as np

from . import timeline

logger = logging.getLogger("edl.distill")

_NOP_PREDICT_TEST = False   # tests flip this to use NopPredictClient (reference distill_worker.py:36)


class ReaderType:
    SAMPLE = 0        # generator yields one sample tuple at a time
    SAMPLE_LIST = 1   # generator yields a list of sample tuples (a batch)
    BATCH = 2         # generator yields a tuple of stacked arrays


class Task:
    __slots__ = ("task_id", "batch_id", "last_of_batch", "samples", "preds", "epoch")

    def __init__(self, task_id, batch_id, last_of_batch, samples, epoch):
        self.task_id, self.batch_id, self.last_of_batch = task_id, batch_id, last_of_batch
        self.samples, self.preds, self.epoch = samples, None, epoch


class _EpochEnd:
    def __init__(self, n_tasks, epoch, error=None):
        self.n_tasks, self.epoch, self.error = n_tasks, epoch, error


def _chunks(seq, n):
    for i in range(0, len(seq), n):
        yield seq[i:i + n]


def reader_worker(reader, reader_type, teacher_batch_size, in_q, out_q, sem, stop, epoch):
    """Producer: user data -> tasks.  Runs in its own thread for one epoch."""
    task_id = 0
    try:
        if reader_type == ReaderType.SAMPLE:
            buf = []
            for sample in reader():
                buf.append(tuple(sample))
                if len(buf) == teacher_batch_size:
                    if not _acquire(sem, stop):
                        return
                    in_q.put(Task(task_id, task_id, True, buf, epoch))
                    task_id, buf = task_id + 1, []
            if buf and _acquire(sem, stop):
                in_q.put(Task(task_id, task_id, True, buf, epoch))
                task_id += 1
        else:
            for batch_id, batch in enumerate(reader()):
                if reader_type == ReaderType.BATCH:
                    slots = [np.asarray(s) for s in batch]
                    n = len(slots[0])
                    samples = [tuple(s[i] for s in slots) for i in range(n)]
                else:
                    samples = [tuple(s) for s in batch]
                parts = list(_chunks(samples, teacher_batch_size))
                for j, part in enumerate(parts):
                    if not _acquire(sem, stop):
                        return
                    in_q.put(Task(task_id, batch_id, j == len(parts) - 1, part, epoch))
                    task_id += 1
        out_q.put(_EpochEnd(task_id, epoch))
    except Exception as e:  # noqa: BLE001 - surface reader errors in the consumer
        logger.exception("reader failed")
        out_q.put(_EpochEnd(task_id, epoch, error=e))


def _reader_process_main(reader, reader_type, teacher_batch_size, mp_q, stop, epoch):
    """Child process (forked): the USER's reader runs here, off the student's GIL.  Tasks travel as plain tuples; the
    parent's pump thread applies the back-pressure semaphore and feeds the predict pool."""
    local_in, local_out = queue.Queue(), queue.Queue()

    class _NoSem:                      # the bounded mp queue is the back-pressure on this side of the pipe
        def acquire(self, timeout=None):
            return True

    t = threading.Thread(target=reader_worker, daemon=True,
                         args=(reader, reader_type, teacher_batch_size, local_in, local_out, _NoSem(), stop, epoch))
    t.start()
    try:
        while True:
            try:
                task = local_in.get(timeout=0.05)
                mp_q.put(("task", task.task_id, task.batch_id, task.last_of_batch, task.samples))
                continue
            except queue.Empty:
                pass
            if not t.is_alive() and local_in.empty():
                break
        end = local_out.get(timeout=5)
        mp_q.put(("end", end.n_tasks, repr(end.error) if end.error is not None else None))
    except Exception as e:  # noqa: BLE001
        mp_q.put(("end", -1, repr(e)))


def reader_process_pump(reader, reader_type, teacher_batch_size, in_q, out_q, sem, stop, epoch):
    """``reader_worker`` with the user's generator in a FORKED PROCESS (the reference's design,
    distill_worker.py:46-120: reader and predict workers are processes): a Python-heavy reader (cv2 decode, augmentation
    in numpy / pure Python) no longer competes with the training loop and the predict threads for the GIL.  Samples are
    pickled once across the pipe; the thread version stays the default because it hands numpy arrays over by reference."""
    import multiprocessing as mp

    ctx = mp.get_context("fork")       # user readers are closures: not picklable, so no spawn
    mp_q = ctx.Queue(maxsize=8)
    child_stop = ctx.Event()
    proc = ctx.Process(target=_reader_process_main, daemon=True,
                       args=(reader, reader_type, teacher_batch_size, mp_q, child_stop, epoch))
    proc.start()
    n_tasks, error = 0, None
    try:
        while True:
            if stop.is_set():
                return
            try:
                msg = mp_q.get(timeout=0.2)
            except queue.Empty:
                if not proc.is_alive():
                    error = RuntimeError("the distill reader process died (exit code %s)" % proc.exitcode)
                    break
                continue
            if msg[0] == "end":
                if msg[2] is not None:
                    error = RuntimeError("reader failed in its process: %s" % msg[2])
                n_tasks = max(n_tasks, msg[1])
                break
            _, task_id, batch_id, last, samples = msg
            if not _acquire(sem, stop):
                return
            in_q.put(Task(task_id, batch_id, last, samples, epoch))
            n_tasks = task_id + 1
        out_q.put(_EpochEnd(n_tasks, epoch, error=error))
    finally:
        child_stop.set()
        proc.join(2)
        if proc.is_alive():
            proc.terminate()


def _acquire(sem, stop):
    while not stop.is_set():
        if sem.acquire(timeout=0.1):
            return True
    return False


class PredictWorker(threading.Thread):
    """One teacher connection: pull task -> predict -> push result; on failure re-queue the task and
    retire (the manager will reconnect or replace the teacher)."""

    def __init__(self, client, feeds, fetchs, in_q, out_q):
        super().__init__(daemon=True, name="distill-predict-%s" % client.server)
        self.client, self.feeds, self.fetchs, self.in_q, self.out_q = client, feeds, fetchs, in_q, out_q
        self.stop_event = threading.Event()
        self.failed = False
        self.n_done = 0

    def run(self):
        tl = timeline.TimeLine()
        try:
            self.client.connect()
        except Exception as e:  # noqa: BLE001
            logger.warning("cannot connect teacher %s: %s", self.client.server, e)
            self.failed = True
            return
        names = [(i, n) for i, n in enumerate(self.feeds) if n is not None and
                 (getattr(self.client, "teacher_feeds", None) is None or n in self.client.teacher_feeds)]
        while not self.stop_event.is_set():
            try:
                task = self.in_q.get(timeout=0.2)
            except queue.Empty:
                continue
            tl.record("get_data")
            try:
                feed_batch = [{n: s[i] for i, n in names} for s in task.samples]
                tl.record("predict_preprocess")
                preds = self.client.predict(feed_batch)
                tl.record("real_predict")
                task.preds = [tuple(p[k] for k in self.fetchs) for p in preds]
                tl.record("postprocess")
            except Exception as e:  # noqa: BLE001
                logger.warning("teacher %s failed (%s); re-queueing task %d", self.client.server, e, task.task_id)
                self.in_q.put(task)
                self.failed = True
                break
            self.out_q.put(task)
            self.n_done += 1
            tl.record("put_data")
        self.client.close()


class PredictPool:
    """Keeps <= require_num PredictWorkers matched to what service discovery currently returns."""

    def __init__(self, discover, make_client, feeds, fetchs, in_q, out_q, require_num, poll_s=2.0):
        self.discover, self.make_client = discover, make_client
        self.feeds, self.fetchs, self.in_q, self.out_q = feeds, fetchs, in_q, out_q
        self.require_num, self.poll_s = require_num, poll_s
        self.workers = {}
        self._stop = threading.Event()
        self._lock = threading.Lock()
        self._t = threading.Thread(target=self._manage, daemon=True, name="distill-predict-manager")
        self._t.start()

    def _reconcile(self):
        try:
            servers = list(self.discover.get_servers() or [])
        except Exception as e:  # noqa: BLE001
            logger.warning("service discovery failed: %s", e)
            return
        with self._lock:
            for srv, w in list(self.workers.items()):
                if not w.is_alive() or w.failed:
                    self.workers.pop(srv)
                elif srv not in servers:
                    logger.info("teacher %s retired", srv)
                    w.stop_event.set()
                    self.workers.pop(srv)
            for srv in servers:
                if len(self.workers) >= self.require_num:
                    break
                if srv not in self.workers:
                    w = PredictWorker(self.make_client(srv), self.feeds, self.fetchs, self.in_q, self.out_q)
                    w.start()
                    self.workers[srv] = w
                    logger.info("teacher %s connected (%d/%d)", srv, len(self.workers), self.require_num)

    def _manage(self):
        while not self._stop.is_set():
            self._reconcile()
            # react quickly while we are short of teachers, lazily otherwise
            self._stop.wait(0.2 if len(self.workers) < self.require_num else self.poll_s)

    def num_workers(self):
        with self._lock:
            return sum(1 for w in self.workers.values() if w.is_alive() and not w.failed)

    def stop(self):
        self._stop.set()
        self._t.join(3)
        with self._lock:
            for w in self.workers.values():
                w.stop_event.set()
            for w in self.workers.values():
                w.join(3)
            self.workers.clear()


def fetch_out(reader_type, out_q, sem, stop, epoch, idle_warn_s=30.0):
    """Consumer side: re-order by task id, restore batch boundaries, release back-pressure."""
    pending = {}
    next_id = 0
    total = None
    batch_acc = []
    last_progress = time.time()
    while not stop.is_set():
        if total is not None and next_id >= total:
            return
        try:
            item = out_q.get(timeout=0.5)
        except queue.Empty:
            if time.time() - last_progress > idle_warn_s:
                logger.warning("distill reader waiting for teachers (%d tasks pending)", len(pending))
                last_progress = time.time()
            continue
        if isinstance(item, _EpochEnd):
            if item.epoch != epoch:
                continue
            if item.error is not None:
                raise item.error
            total = item.n_tasks
            continue
        if item.epoch != epoch:
            continue  # stale result of an aborted epoch
        pending[item.task_id] = item
        while next_id in pending:
            task = pending.pop(next_id)
            next_id += 1
            last_progress = time.time()
            sem.release()
            merged = [s + p for s, p in zip(task.samples, task.preds)]
            if reader_type == ReaderType.SAMPLE:
                for m in merged:
                    yield m
            else:
                batch_acc.extend(merged)
                if task.last_of_batch:
                    if reader_type == ReaderType.SAMPLE_LIST:
                        yield batch_acc
                    else:
                        nslot = len(batch_acc[0])
                        yield tuple(np.stack([np.asarray(m[i]) for m in batch_acc]) for i in range(nslot))
                    batch_acc = []
