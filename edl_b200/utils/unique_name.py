"""Process-wide unique-name generator (reference: python/edl/utils/unique_name.py)."""
import collections
import threading


class UniqueNameGenerator:
    def __init__(self, prefix=None):
        self.ids = collections.defaultdict(int)
        self.prefix = prefix or ""
        self._lock = threading.Lock()

    def __call__(self, key):
        with self._lock:
            n = self.ids[key]
            self.ids[key] += 1
        return self.prefix + "_".join([key, str(n)])


generator = UniqueNameGenerator()


def generate(key):
    return generator(key)
