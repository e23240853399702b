"""``python -m paddle_edl.liveft.launch --elastic_server h:p --job_id J --np N train.py ...``
(reference: python/edl/liveft/launch.py:24-59): ``while True: wait(); run(); watch()``."""
import argparse
import signal
import sys

from .elastic import ELASTIC_EXIT_CODE, ElasticManager, ElasticStatus, ProcessLauncher


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="live fault tolerant launcher")
    p.add_argument("--elastic_server", type=str, default=None, help="store endpoint host:port")
    p.add_argument("--job_id", type=str, default=None)
    p.add_argument("--np", type=int, default=None, help="number of nodes the job needs")
    p.add_argument("--host", type=str, default=None)
    p.add_argument("--scale", type=int, default=0)
    p.add_argument("--force", type=str, default=None)
    p.add_argument("training_script", type=str)
    p.add_argument("training_script_args", nargs=argparse.REMAINDER)
    return p.parse_args(argv)


def launch(argv=None, launcher_cls=ProcessLauncher):
    args = parse_args(argv)
    elastic_manager = ElasticManager(args)
    signal.signal(signal.SIGTERM, elastic_manager.signal_handler)
    signal.signal(signal.SIGINT, elastic_manager.signal_handler)
    while True:
        elastic_manager.wait()                 # block until exactly np nodes are registered
        if elastic_manager.job_done:
            elastic_manager.exit(completed=True)
            return 0
        elastic_manager.run(launcher_cls)      # start the local training process(es)
        ret = elastic_manager.watch()          # supervise
        if ret == ElasticStatus.COMPLETED:
            return 0
        if ret == ElasticStatus.HOLD:
            continue                           # membership changed: children stopped, re-wait
        if ret == ElasticStatus.EXIT:
            break
        if ret == ElasticStatus.ERROR:
            return 3
        if ret == ElasticStatus.RESTART:
            return ELASTIC_EXIT_CODE
    if int(elastic_manager.sigint) > 0:
        return 128 + int(elastic_manager.sigint)
    return 0


if __name__ == "__main__":
    sys.exit(launch())
