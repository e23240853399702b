"""Background-process base with a cooperative stop event (reference: python/edl/utils/process.py:21-51)."""
import multiprocessing
import threading


class ProcessWrapper:
    """Runs ``self._worker_func()`` in a child process (or a thread when ``use_thread``)."""

    def __init__(self, use_thread=False):
        self._use_thread = use_thread
        self._stop = threading.Event() if use_thread else multiprocessing.Event()
        self._worker = None

    def _worker_func(self):
        raise NotImplementedError

    def start(self):
        if self._use_thread:
            self._worker = threading.Thread(target=self._worker_func, daemon=True)
        else:
            self._worker = multiprocessing.Process(target=self._worker_func, daemon=True)
        self._worker.start()
        return self

    def stop(self, timeout=10):
        self._stop.set()
        if self._worker is not None:
            self._worker.join(timeout)
            if not self._use_thread and self._worker.is_alive():
                self._worker.terminate()
            self._worker = None

    def is_stopped(self):
        return self._worker is None or not self._worker.is_alive()

    def should_stop(self):
        return self._stop.is_set()
