"""Retry envelope (reference: python/edl/utils/error_utils.py:22-39)."""
import functools
import time

from . import constants
from .exceptions import EdlException
from .log_utils import logger


def handle_errors_until_timeout(f):
    """Retry ``f`` on any :class:`EdlException` every ``retry_interval`` (default: poll interval)
    until the ``timeout=`` keyword elapses, then re-raise the last error."""

    @functools.wraps(f)
    def handler(*args, **kwargs):
        timeout = float(kwargs.get("timeout", constants.ETCD_OPERATION_TIMEOUT))
        interval = float(kwargs.pop("retry_interval", min(constants.POLL_INTERVAL, 3.0)))
        begin = time.time()
        while True:
            try:
                return f(*args, **kwargs)
            except EdlException as e:
                if time.time() - begin >= timeout:
                    logger.warning("%s timed out after %.1fs: %s", f.__name__, timeout, e)
                    raise
                time.sleep(min(interval, max(0.0, timeout - (time.time() - begin))) or 0.01)

    return handler
