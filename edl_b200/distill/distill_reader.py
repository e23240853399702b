"""``DistillReader`` -- the user-facing distillation data API.

    dr = DistillReader(ins=['image', 'label'], predicts=['score'])
    dr.set_teacher_batch_size(16)
    dr.set_fixed_teacher("10.0.0.1:9292,10.0.0.2:9292")            # or
    dr.set_dynamic_teacher(["10.0.0.9:7001"], "ResNeXt101", require_max_teacher=4)
    train_reader = dr.set_sample_list_generator(batch_reader)
    for batch in train_reader():        # each sample = original slots + one slot per `predicts`
        ...

Same constructor, setters, environment overrides (``PADDLE_DISTILL_*`` win over setters) and
generator protocol as the reference (python/edl/distill/distill_reader.py:85-416); the pipeline
behind it is ``distill_worker`` (threads + gRPC teachers) and teachers are found through
``FixedServiceDiscover`` / ``DynamicServiceDiscover`` (etcd-style discovery service or the
redis-style balance server, chosen by ``PADDLE_DISTILL_BALANCE_TYPE``)."""
import logging
import os
import queue
import threading

from . import distill_worker
from .serving_conf import load_serving_conf

logger = logging.getLogger("edl.distill")


class ServiceDiscover:
    def get_servers(self):
        raise NotImplementedError

    def stop(self):
        pass


class FixedServiceDiscover(ServiceDiscover):
    def __init__(self, servers):
        self._servers = list(servers)

    def get_servers(self):
        return self._servers


class DynamicServiceDiscover(ServiceDiscover):
    """``PADDLE_DISTILL_BALANCE_TYPE`` = etcd | redis (default redis, like the reference :53)."""

    def __init__(self, discovery_servers, service_name, require_num, balance_type=None):
        kind = (balance_type or os.environ.get("PADDLE_DISTILL_BALANCE_TYPE", "redis")).lower()
        if kind == "etcd":
            from .discovery_client import DiscoveryClient

            self._client = DiscoveryClient(discovery_servers, service_name, require_num)
        else:
            from .redis.client import Client

            self._client = Client(discovery_servers, service_name, require_num)
        self._client.start()

    def get_servers(self):
        return self._client.get_servers()

    def stop(self):
        self._client.stop()


_service_discover = None
_service_discover_lock = threading.Lock()


class DistillReader:
    def __init__(self, ins, predicts):
        self._feeds = list(ins)
        self._fetchs = list(predicts)
        self._serving_conf_file = "./serving_conf/serving_client_conf.prototxt"
        self._teacher_batch_size = 1
        self._mode = None
        self._teachers = []
        self._require_num = 1
        self._discovery_servers = []
        self._service_name = None
        self._reader = None
        self._reader_type = None
        self._is_args_init = False
        self._pool = None
        self._discover = None
        self._in_q = self._out_q = self._sem = None
        self._stop = threading.Event()
        self._epoch = 0
        self._client_factory = None
        self._lock = threading.Lock()
        self._reader_in_process = os.environ.get("EDL_DISTILL_READER_PROCESS", "0") == "1"

    # ------------------------------------------------------------------ configuration
    def set_serving_conf_file(self, conf_file):
        assert os.path.isfile(conf_file), "{} is not file".format(conf_file)
        self._serving_conf_file = conf_file

    def set_teacher_batch_size(self, teacher_batch_size=1):
        self._teacher_batch_size = int(teacher_batch_size)

    def set_fixed_teacher(self, teachers):
        if isinstance(teachers, (list, tuple)):
            self._teachers = list(teachers)
        elif isinstance(teachers, str):
            self._teachers = [t for t in teachers.split(",") if t]
        else:
            raise TypeError("teachers must be list|tuple|str")
        self._mode = "fixed"
        self._require_num = len(self._teachers)

    def set_dynamic_teacher(self, discovery_servers, teacher_service_name, require_max_teacher=1):
        if isinstance(discovery_servers, (list, tuple)):
            self._discovery_servers = list(discovery_servers)
        elif isinstance(discovery_servers, str):
            self._discovery_servers = [s for s in discovery_servers.split(",") if s]
        else:
            raise TypeError("discovery_servers must be list|tuple|str")
        self._mode = "discover"
        self._service_name = teacher_service_name
        self._require_num = int(require_max_teacher)

    def set_require_max_teacher(self, require_max_teacher):
        if self._mode == "fixed":
            return
        self._require_num = int(require_max_teacher)

    def set_predict_client_factory(self, factory):
        """Plug another teacher transport: ``factory(server, feeds, fetchs, conf) -> PredictClient``."""
        self._client_factory = factory

    def set_reader_process(self, enabled=True):
        """Run the user's reader generator in a forked process instead of a thread (the reference forks its reader
        worker): worth it when the reader is Python-heavy (decode / augmentation holding the GIL); costs one pickle of
        every sample.  Also ``EDL_DISTILL_READER_PROCESS=1``."""
        self._reader_in_process = bool(enabled)
        return self

    def set_sample_generator(self, reader):
        assert self._reader is None, "reader has already set"
        self._reader, self._reader_type = reader, distill_worker.ReaderType.SAMPLE
        return self

    def set_sample_list_generator(self, reader):
        assert self._reader is None, "reader has already set"
        self._reader, self._reader_type = reader, distill_worker.ReaderType.SAMPLE_LIST
        return self

    def set_batch_generator(self, reader):
        assert self._reader is None, "reader has already set"
        self._reader, self._reader_type = reader, distill_worker.ReaderType.BATCH
        return self

    def print_config(self):
        print("------ DistillReader Configuration Arguments ------")
        if not self._is_args_init:
            print("DistillReader not start yet, some args may change.")
        for k, v in {
            "ins": self._feeds, "predicts": self._fetchs, "serving_conf_file": self._serving_conf_file,
            "teacher_batch_size": self._teacher_batch_size, "distill_mode": self._mode,
            "teachers": self._teachers, "require_max_teacher": self._require_num,
            "discovery_servers": self._discovery_servers, "teacher_service_name": self._service_name,
            "reader_type": self._reader_type,
        }.items():
            print("%s: %s" % (k, v))
        print("------------------------------------------------")

    # ------------------------------------------------------------------ start-up
    def _init_from_env(self):
        """Environment has the highest priority (reference :255-298)."""
        conf = os.environ.get("PADDLE_DISTILL_CONF_FILE")
        if not os.path.isfile(self._serving_conf_file) and conf and os.path.isfile(conf):
            self._serving_conf_file = conf
        servers = os.environ.get("PADDLE_DISTILL_BALANCE_SERVER")
        if servers is not None:
            name = os.environ.get("PADDLE_DISTILL_SERVICE_NAME")
            assert name is not None, "PADDLE_DISTILL_SERVICE_NAME must accompany PADDLE_DISTILL_BALANCE_SERVER"
            self._mode, self._discovery_servers, self._service_name = "discover", servers.split(","), name
            mt = os.environ.get("PADDLE_DISTILL_MAX_TEACHER")
            if mt is not None:
                self._require_num = int(mt)
        assert self._mode is not None, (
            "Teacher is empty: use set_fixed_teacher / set_dynamic_teacher or the PADDLE_DISTILL_* environment")

    def _make_client(self, server):
        from . import predict_client

        conf = load_serving_conf(self._serving_conf_file) if os.path.isfile(self._serving_conf_file) else None
        if self._client_factory is not None:
            return self._client_factory(server, self._feeds, self._fetchs, conf)
        if distill_worker._NOP_PREDICT_TEST:
            return predict_client.NopPredictClient(server, self._feeds, self._fetchs, conf)
        return predict_client.GrpcPredictClient(server, self._feeds, self._fetchs, conf)

    def _get_discover(self):
        global _service_discover
        if self._mode == "fixed":
            return FixedServiceDiscover(self._teachers)
        with _service_discover_lock:   # one registration per process, shared by all readers
            if _service_discover is None:
                _service_discover = DynamicServiceDiscover(self._discovery_servers, self._service_name,
                                                           self._require_num)
            return _service_discover

    def _init_args(self):
        if self._is_args_init:
            return
        self._init_from_env()
        self._in_q, self._out_q = queue.Queue(), queue.Queue()
        self._sem = threading.Semaphore(2 * self._require_num + 2)
        self._discover = self._get_discover()
        self._pool = distill_worker.PredictPool(self._discover, self._make_client, self._feeds, self._fetchs,
                                                self._in_q, self._out_q, self._require_num)
        self._is_args_init = True

    # ------------------------------------------------------------------ the generator
    def __call__(self):
        assert self._reader is not None, "must set reader before iter DistillReader"
        with self._lock:
            self._init_args()
            self._epoch += 1
            epoch = self._epoch
        epoch_stop = threading.Event()
        target = distill_worker.reader_process_pump if self._reader_in_process else distill_worker.reader_worker
        t = threading.Thread(target=target, daemon=True, name="distill-reader",
                             args=(self._reader, self._reader_type, self._teacher_batch_size, self._in_q,
                                   self._out_q, self._sem, epoch_stop, epoch))
        t.start()
        finished = False
        try:
            for data in distill_worker.fetch_out(self._reader_type, self._out_q, self._sem, self._stop, epoch):
                yield data
            finished = True
        finally:
            epoch_stop.set()
            if not finished:
                self._abort_epoch(t)
            t.join(5)

    def _abort_epoch(self, reader_thread):
        """The consumer stopped early: drop queued tasks and give the semaphore permits back."""
        reader_thread.join(2)
        dropped = 0
        while True:
            try:
                self._in_q.get_nowait()
                dropped += 1
            except queue.Empty:
                break
        while True:
            try:
                self._out_q.get_nowait()
            except queue.Empty:
                break
        self._sem = threading.Semaphore(2 * self._require_num + 2)

    def stop(self):
        self._stop.set()
        if self._pool is not None:
            self._pool.stop()
            self._pool = None

    def __del__(self):
        try:
            self.stop()
        except Exception:  # noqa: BLE001
            pass
