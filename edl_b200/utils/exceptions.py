"""EDL error hierarchy + (de)serialisation into the ``common.Status{type, detail}`` RPC message
(reference: python/edl/utils/exceptions.py:20-117)."""


class EdlException(Exception):
    pass


def _mk(name, base=EdlException):
    return type(name, (base,), {"__module__": __name__})


EdlRegisterError = _mk("EdlRegisterError")
EdlBarrierError = _mk("EdlBarrierError")
EdlUnkownError = _mk("EdlUnkownError")
EdlRankError = _mk("EdlRankError")
EdlInternalError = _mk("EdlInternalError")
EdlWaitFollowersReleaseError = _mk("EdlWaitFollowersReleaseError")
EdlLeaderError = _mk("EdlLeaderError")
EdlGenerateClusterError = _mk("EdlGenerateClusterError")
EdlTableError = _mk("EdlTableError")
EdlEtcdIOError = _mk("EdlEtcdIOError")
EdlDataEndError = _mk("EdlDataEndError")
EdlPodIDNotExistError = _mk("EdlPodIDNotExistError")
EdlReaderNameError = _mk("EdlReaderNameError")
EdlFileListNotMatchError = _mk("EdlFileListNotMatchError")
EdlDataGenerateError = _mk("EdlDataGenerateError")
EdlAccessDataError = _mk("EdlAccessDataError")
EdlStopIteration = _mk("EdlStopIteration")
EdlNotLeaderError = _mk("EdlNotLeaderError")
EdlNotFoundLeader = _mk("EdlNotFoundLeader")
EdlCommTimeoutError = _mk("EdlCommTimeoutError")   # a peer vanished mid-collective (device flag)

_BY_NAME = {k: v for k, v in list(globals().items()) if isinstance(v, type) and issubclass(v, EdlException)}


def serialize(pb_status, exc: Exception) -> None:
    """Fill a ``common.Status`` from an exception (empty type == success)."""
    pb_status.type = type(exc).__name__
    pb_status.detail = str(exc)


def deserialize(pb_status) -> None:
    """Raise the exception carried by a ``common.Status``; no-op on success."""
    if not pb_status.type:
        return
    cls = _BY_NAME.get(pb_status.type, EdlUnkownError)
    raise cls(pb_status.detail)
