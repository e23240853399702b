"""File-system abstraction for checkpoints: ``LocalFS`` (any POSIX path shared by the trainers) and an
``HDFSClient`` that shells out to the ``hadoop fs`` CLI when it exists (reference uses Paddle's
``LocalFS`` / ``BDFS(hdfs_name, ugi, 20 min timeout, 3 s sleep)``,
example/collective/resnet50/train_with_fleet.py:422-424)."""
import os
import shutil
import subprocess
import time


class FS:
    def ls_dir(self, path): raise NotImplementedError
    def is_exist(self, path): raise NotImplementedError
    def mkdirs(self, path): raise NotImplementedError
    def delete(self, path): raise NotImplementedError
    def mv(self, src, dst): raise NotImplementedError
    def need_upload_download(self): return False


class LocalFS(FS):
    def ls_dir(self, path):
        """-> (dirs, files)"""
        if not os.path.isdir(path):
            return [], []
        dirs, files = [], []
        for f in sorted(os.listdir(path)):
            (dirs if os.path.isdir(os.path.join(path, f)) else files).append(f)
        return dirs, files

    def is_exist(self, path): return os.path.exists(path)
    def is_dir(self, path): return os.path.isdir(path)
    def is_file(self, path): return os.path.isfile(path)

    def mkdirs(self, path):
        os.makedirs(path, exist_ok=True)

    def delete(self, path):
        if os.path.isdir(path):
            shutil.rmtree(path, ignore_errors=True)
        elif os.path.exists(path):
            os.remove(path)

    def mv(self, src, dst):
        os.replace(src, dst)  # atomic on one file system

    def touch(self, path):
        open(path, "a").close()


class HDFSClient(FS):
    """Thin wrapper over ``hadoop fs``; every call retries until ``time_out`` ms like BDFS."""

    def __init__(self, hdfs_name=None, hdfs_ugi=None, time_out=20 * 60 * 1000, sleep_inter=3000, hadoop_bin="hadoop"):
        self._base = [hadoop_bin, "fs"]
        if hdfs_name:
            self._base += ["-D", "fs.default.name=%s" % hdfs_name]
        if hdfs_ugi:
            self._base += ["-D", "hadoop.job.ugi=%s" % hdfs_ugi]
        self._time_out, self._sleep = time_out / 1000.0, sleep_inter / 1000.0
        self.available = shutil.which(hadoop_bin) is not None

    def _run(self, args):
        if not self.available:
            raise RuntimeError("hadoop CLI not found; use LocalFS on a shared path")
        begin = time.time()
        while True:
            r = subprocess.run(self._base + args, capture_output=True, text=True)
            if r.returncode == 0:
                return r.stdout
            if time.time() - begin > self._time_out:
                raise RuntimeError("hadoop fs %s failed: %s" % (args, r.stderr[-500:]))
            time.sleep(self._sleep)

    def need_upload_download(self): return True

    def ls_dir(self, path):
        out = self._run(["-ls", path])
        dirs, files = [], []
        for ln in out.splitlines():
            f = ln.split()
            if len(f) >= 8:
                (dirs if f[0].startswith("d") else files).append(os.path.basename(f[-1]))
        return dirs, files

    def is_exist(self, path):
        if not self.available:
            return False
        return subprocess.run(self._base + ["-test", "-e", path]).returncode == 0

    def mkdirs(self, path): self._run(["-mkdir", "-p", path])
    def delete(self, path): self._run(["-rm", "-r", "-f", path])
    def mv(self, src, dst): self._run(["-mv", src, dst])
    def upload(self, local, remote): self._run(["-put", "-f", local, remote])
    def download(self, remote, local): self._run(["-get", remote, local])


BDFS = HDFSClient        # the reference's name for the same client (fleet.utils.fs / BDFS(hdfs_name, ugi, time_out, sleep))


def get_fs(hdfs_name=None, hdfs_ugi=None):
    """HDFS when configured *and* the CLI exists, else LocalFS."""
    if hdfs_name and hdfs_ugi:
        fs = HDFSClient(hdfs_name, hdfs_ugi)
        if fs.available:
            return fs
    return LocalFS()
