"""Fetch the teacher serving conf from a shared file system
(reference: python/edl/distill/utils.py:19-34 downloads it from HDFS via BDFS)."""
import os
import shutil

from ..checkpoint.fs import HDFSClient


def get_conf_file(default_path="./serving_conf/serving_client_conf.prototxt"):
    """Resolve the conf file: PADDLE_DISTILL_CONF_FILE, else HDFS download when
    PADDLE_DISTILL_HDFS_{NAME,UGI,PATH} are set, else ``default_path``."""
    env = os.environ.get("PADDLE_DISTILL_CONF_FILE")
    if env and os.path.isfile(env):
        return env
    name, ugi, path = (os.environ.get("PADDLE_DISTILL_HDFS_NAME"), os.environ.get("PADDLE_DISTILL_HDFS_UGI"),
                       os.environ.get("PADDLE_DISTILL_HDFS_PATH"))
    if path:
        os.makedirs(os.path.dirname(default_path) or ".", exist_ok=True)
        if name and ugi:
            fs = HDFSClient(name, ugi)
            if fs.available:
                fs.download(path, default_path)
                return default_path
        if os.path.isfile(path):  # shared POSIX path
            shutil.copyfile(path, default_path)
            return default_path
    return default_path
