"""Launcher CLI (reference: python/edl/utils/args_utils.py:31-100)."""
import argparse


def parse_args(argv=None):
    p = argparse.ArgumentParser(
        description="Start an elastic training pod: rendezvous through the store, spawn one trainer "
                    "process per GPU, re-barrier and restart them when the membership changes.")
    p.add_argument("--nodes_range", type=str, default=None, help="min:max number of pods, e.g. 2:8")
    p.add_argument("--nproc_per_node", type=int, default=None, help="trainer processes on this node")
    p.add_argument("--etcd_endpoints", type=str, default=None, help="store endpoints, comma separated")
    p.add_argument("--job_id", type=str, default=None, help="unique job id (store namespace)")
    p.add_argument("--log_level", type=int, default=20, help="logging level (20 = INFO)")
    p.add_argument("--log_dir", type=str, default="./log", help="directory of workerlog.N files")
    p.add_argument("--hdfs_name", type=str, default=None)
    p.add_argument("--hdfs_ugi", type=str, default=None)
    p.add_argument("--hdfs_path", type=str, default=None, help="checkpoint path visible to all trainers")
    p.add_argument("--rescale_mode", type=str, default=None, choices=["restart", "inplace"],
                   help="restart (reference behaviour: kill and restart trainers on every membership change) or "
                        "inplace (trainers that own an ElasticContext keep running; env EDL_RESCALE_MODE)")
    p.add_argument("training_script", type=str, help="the single-GPU training program")
    p.add_argument("training_script_args", nargs=argparse.REMAINDER)
    return p.parse_args(argv)


def convert_args_to_dict(args):
    return {k: v for k, v in vars(args).items() if v is not None}
