"""Small string helpers (reference: python/edl/utils/string_utils.py)."""


def bytes_to_string(o, codec="utf-8"):
    if o is None:
        return None
    return o.decode(codec) if isinstance(o, (bytes, bytearray)) else o


def string_to_bytes(o, codec="utf-8"):
    if o is None:
        return None
    return o.encode(codec) if isinstance(o, str) else o


def dataset_to_string(o):
    """Render a collection of ids for log messages."""
    return "[" + ",".join(str(x) for x in o) + "]"


def trim_brackets(s: str) -> str:
    return s.strip().lstrip("[").rstrip("]")
