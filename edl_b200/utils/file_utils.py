"""File helpers (reference: python/edl/utils/file_utils.py)."""
import os


def read_txt_lines(path):
    with open(path, "r") as f:
        return [ln.rstrip("\n") for ln in f]


def make_dirs(path):
    os.makedirs(path, exist_ok=True)
    return path
