"""Checkpoint-to-file-system: atomic, versioned, rank-0 writer (the semantics the reference relies
on from ``fleet.save_check_point / load_check_point`` + ``LocalFS``/``BDFS``; doc/fault_tolerance.md:13-25)."""
from .fs import BDFS, LocalFS, HDFSClient, get_fs
from .checkpoint import (save_check_point, load_check_point, latest_version, list_versions, clean_redundant,
                         TrainStatus)

__all__ = ["LocalFS", "HDFSClient", "BDFS", "get_fs", "save_check_point", "load_check_point", "latest_version",
           "list_versions", "clean_redundant", "TrainStatus"]
