"""``python -m paddle_edl.collective.launch [flags] train.py [args...]`` -- start one elastic pod
(reference: python/edl/collective/launch.py:32-55; console script ``edlrun``)."""
import sys

from ..discovery.etcd_client import EtcdClient
from ..utils import args_utils, constants, env as edl_env, launcher as edl_launcher, log_utils, status as edl_status
from ..utils.pod import Pod


def main(argv=None):
    args = args_utils.parse_args(argv)
    if args.rescale_mode:
        import os

        os.environ["EDL_RESCALE_MODE"] = args.rescale_mode      # trainers inherit it
    logger = log_utils.get_logger(args.log_level)
    job_env = edl_env.JobEnv(args_utils.convert_args_to_dict(args))
    etcd = EtcdClient(endpoints=job_env.etcd_endpoints, root=job_env.job_id, timeout=constants.ETCD_CONN_TIMEOUT)
    etcd.init()
    if edl_status.load_job_status_from_etcd(etcd, timeout=15) == edl_status.Status.SUCCEED:
        logger.info("job %s already succeeded; nothing to do", job_env.job_id)
        return 0
    pod = Pod().from_env(job_env)
    launcher = edl_launcher.Launcher(job_env=job_env, pod=pod, etcd=etcd, args=args)
    launcher.init()
    try:
        import signal

        # SIGTERM = "this pod has to go" (scheduler, k8s): leave gracefully instead of vanishing (launcher.request_leave);
        # a second SIGTERM, or SIGKILL after the grace period, still ends the process the hard way
        def _on_term(signum, frame):
            if launcher._leave:
                raise SystemExit(128 + signum)
            launcher.request_leave()

        signal.signal(signal.SIGTERM, _on_term)
    except ValueError:      # not the main thread (embedded use): no handler
        pass
    ok = launcher.launch()
    etcd.close()
    return 0 if ok else 1


def run_commandline():
    sys.exit(main())


if __name__ == "__main__":
    run_commandline()
