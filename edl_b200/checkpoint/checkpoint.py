"""Versioned atomic checkpoints.

Layout:  ``<path>/__edl_checkpoint__.<N>/{state.pt, meta.json}``.  A save writes into
``...<N>.tmp.<uuid>`` and atomically renames it, finished versions are never mutated, only rank 0
(``trainer_id == 0``) writes, loads pick the highest complete version -- the rules spelled out in
the reference's doc/fault_tolerance.md:13-25 and used through ``fleet.save_check_point`` /
``fleet.load_check_point`` (example/collective/resnet50/train_with_fleet.py:422-434,562-570).
"""
import json
import os
import re
import shutil
import tempfile
import time
import uuid

import torch

from .fs import LocalFS

_PREFIX = "__edl_checkpoint__"
_RE = re.compile(r"^%s\.(\d+)$" % re.escape(_PREFIX))


class TrainStatus:
    """Minimal epoch cursor (fleet's ``TrainStatus(epoch)`` / ``.next()``)."""

    def __init__(self, epoch_no=-1, global_step=0, extra=None):
        self._epoch_no, self.global_step, self.extra = epoch_no, global_step, extra or {}

    def next(self):
        return self._epoch_no + 1

    @property
    def epoch_no(self):
        return self._epoch_no

    def to_dict(self):
        return {"epoch_no": self._epoch_no, "global_step": self.global_step, "extra": self.extra}

    @staticmethod
    def from_dict(d):
        return TrainStatus(d.get("epoch_no", -1), d.get("global_step", 0), d.get("extra"))

    def __eq__(self, other):
        return isinstance(other, TrainStatus) and self.to_dict() == other.to_dict()


def list_versions(path, fs=None):
    fs = fs or LocalFS()
    dirs, _ = fs.ls_dir(path)
    out = []
    for d in dirs:
        m = _RE.match(d)
        if m and fs.is_exist(os.path.join(path, d, "meta.json")):
            out.append(int(m.group(1)))
    return sorted(out)


def latest_version(path, fs=None):
    v = list_versions(path, fs)
    return v[-1] if v else -1


def clean_redundant(path, fs=None, keep=2):
    fs = fs or LocalFS()
    for v in list_versions(path, fs)[:-keep]:
        fs.delete(os.path.join(path, "%s.%d" % (_PREFIX, v)))
    dirs, _ = fs.ls_dir(path)
    for d in dirs:  # stale temp dirs of crashed writers (their name carries the creation time)
        if d.startswith(_PREFIX) and ".tmp." in d:
            full = os.path.join(path, d)
            try:
                born = int(d.rsplit(".", 1)[-1].split("-")[0], 16) if "-" in d.rsplit(".", 1)[-1] else None
                if born is None:
                    born = os.path.getmtime(full) if not fs.need_upload_download() else time.time()
                if time.time() - born > 3600:
                    fs.delete(full)
            except (OSError, ValueError):
                pass


def _fsync_dir(path):
    try:
        fd = os.open(path, os.O_RDONLY)
        try:
            os.fsync(fd)
        finally:
            os.close(fd)
    except OSError:
        pass


def _write_version_dir(d, tensors, meta):
    """state.pt + meta.json into the LOCAL directory ``d``, both flushed to stable storage."""
    with open(os.path.join(d, "state.pt"), "wb") as f:
        torch.save(tensors, f)
        f.flush()
        os.fsync(f.fileno())
    with open(os.path.join(d, "meta.json"), "w") as f:
        json.dump(meta, f)
        f.flush()
        os.fsync(f.fileno())
    _fsync_dir(d)


def save_check_point(path, tensors, train_status=None, fs=None, trainer_id=0, state_json=None, keep=2):
    """Write version N+1.  ``tensors``: anything ``torch.save`` accepts (e.g. trainer.state_dict()).
    Only ``trainer_id == 0`` writes; returns the new version number (or -1 for non-writers).

    Every byte goes through ``fs``: on a local / POSIX-shared file system the version directory is written in place
    under a temporary name and renamed; on a remote one (``fs.need_upload_download()``, e.g. ``HDFSClient`` / the
    reference's BDFS) it is written to a local scratch directory, uploaded under the temporary name and renamed there
    -- the rename stays the commit point."""
    if trainer_id != 0:
        return -1
    fs = fs or LocalFS()
    fs.mkdirs(path)
    version = latest_version(path, fs) + 1
    final = os.path.join(path, "%s.%d" % (_PREFIX, version))
    tmp = final + ".tmp.%x-%s" % (int(time.time()), uuid.uuid4().hex[:8])
    meta = {"version": version, "time": time.time(),
            "train_status": (train_status or TrainStatus()).to_dict(), "state_json": state_json}
    if fs.need_upload_download():
        scratch = tempfile.mkdtemp(prefix="edl_ckpt_")
        try:
            local = os.path.join(scratch, os.path.basename(tmp))
            os.makedirs(local)
            _write_version_dir(local, tensors, meta)
            fs.upload(local, tmp)
        finally:
            shutil.rmtree(scratch, ignore_errors=True)
    else:
        fs.mkdirs(tmp)
        _write_version_dir(tmp, tensors, meta)
    fs.mv(tmp, final)   # the commit point
    if not fs.need_upload_download():
        _fsync_dir(path)
    clean_redundant(path, fs, keep)
    return version


def load_check_point(path, fs=None, trainer_id=0, map_location="cpu", version=None):
    """-> (tensors, TrainStatus, state_json) of the newest complete version, or (None, TrainStatus(), None)."""
    fs = fs or LocalFS()
    v = latest_version(path, fs) if version is None else version
    if v < 0:
        return None, TrainStatus(), None
    d = os.path.join(path, "%s.%d" % (_PREFIX, v))
    scratch = None
    try:
        if fs.need_upload_download():
            scratch = tempfile.mkdtemp(prefix="edl_ckpt_")
            fs.download(d, scratch)
            d = os.path.join(scratch, os.path.basename(d))
        with open(os.path.join(d, "meta.json")) as f:
            meta = json.load(f)
        tensors = torch.load(os.path.join(d, "state.pt"), map_location=map_location, weights_only=False)
    finally:
        if scratch is not None:
            shutil.rmtree(scratch, ignore_errors=True)
    return tensors, TrainStatus.from_dict(meta["train_status"]), meta.get("state_json")
