"""Logger in the reference's ``LEVEL time file:line] msg`` format (python/edl/utils/log_utils.py:21-32)."""
import logging

logger = logging.getLogger("edl")


def get_logger(log_level=20, name="edl"):
    lg = logging.getLogger(name)
    lg.setLevel(int(log_level))
    if not lg.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter(fmt="%(levelname)s %(asctime)s %(filename)s:%(lineno)d] %(message)s"))
        lg.addHandler(h)
        lg.propagate = False
    return lg
