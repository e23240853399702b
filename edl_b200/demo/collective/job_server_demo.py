"""JobServer demo: an HTTP service that decides, over time, WHICH pods of a job should be running
-- a stand-in for a cluster scheduler that grows and shrinks an elastic job.

The reference ships only the shell scripts and README for this demo (README.md:121-154,
example/demo/collective/start_job_server.sh:26-30; the python module is absent from the snapshot).
Documented surface kept here: flags ``--node_ips --pod_num_of_node --time_interval_to_change
--gpu_num_of_node``, HTTP on port 8180 (``PADDLE_JOBSERVER=http://ip:8180``), pods flip every
``time_interval_to_change`` seconds (900 in the published accuracy run).

HTTP API (JSON):
  GET  /job/<job_id>/pods            -> {"version": n, "pods": {pod_id: {"node_ip", "gpus", "running"}}}
  GET  /job/<job_id>/node/<node_ip>  -> the pods placed on that node
  POST /job/<job_id>/schedule        -> body {"running": [pod_id, ...]}: set the running set explicitly
  GET  /healthz
"""
import argparse
import json
import logging
import random
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

logger = logging.getLogger("edl.jobserver")


class JobState:
    def __init__(self, node_ips, pod_num_of_node, gpu_num_of_node, min_running=1, seed=0):
        self.lock = threading.Lock()
        self.version = 0
        self.pods = {}
        self.min_running = min_running
        self.rng = random.Random(seed)
        for ip in node_ips:
            per = max(1, gpu_num_of_node // max(1, pod_num_of_node))
            for j in range(pod_num_of_node):
                pid = "pod_%s_%d" % (ip.replace(".", "_"), j)
                gpus = list(range(j * per, (j + 1) * per)) if gpu_num_of_node > 0 else []
                self.pods[pid] = {"node_ip": ip, "gpus": gpus, "running": True}

    def snapshot(self, node_ip=None):
        with self.lock:
            pods = {k: dict(v) for k, v in self.pods.items() if node_ip is None or v["node_ip"] == node_ip}
            return {"version": self.version, "pods": pods}

    def set_running(self, running_ids):
        with self.lock:
            for pid, p in self.pods.items():
                p["running"] = pid in running_ids
            self.version += 1
            return self.version

    def flip(self):
        """Pick a new random running subset (never fewer than ``min_running`` pods)."""
        with self.lock:
            ids = sorted(self.pods)
            n = self.rng.randint(max(self.min_running, 1), len(ids))
            keep = set(self.rng.sample(ids, n))
            changed = False
            for pid, p in self.pods.items():
                r = pid in keep
                changed |= r != p["running"]
                p["running"] = r
            if changed:
                self.version += 1
            logger.info("schedule v%d: running %s", self.version, sorted(keep))
            return self.version


class _Handler(BaseHTTPRequestHandler):
    def log_message(self, fmt, *args):  # quiet
        logger.debug(fmt, *args)

    def _send(self, code, obj):
        body = json.dumps(obj).encode()
        self.send_response(code)
        self.send_header("Content-Type", "application/json")
        self.send_header("Content-Length", str(len(body)))
        self.end_headers()
        self.wfile.write(body)

    def do_GET(self):
        parts = [p for p in self.path.split("?")[0].split("/") if p]
        st = self.server.state
        if parts == ["healthz"]:
            return self._send(200, {"ok": True})
        if len(parts) >= 3 and parts[0] == "job" and parts[2] == "pods":
            return self._send(200, st.snapshot())
        if len(parts) == 4 and parts[0] == "job" and parts[2] == "node":
            return self._send(200, st.snapshot(parts[3]))
        self._send(404, {"error": "not found"})

    def do_POST(self):
        parts = [p for p in self.path.split("/") if p]
        n = int(self.headers.get("Content-Length", "0"))
        body = json.loads(self.rfile.read(n) or b"{}")
        if len(parts) == 3 and parts[0] == "job" and parts[2] == "schedule":
            v = self.server.state.set_running(set(body.get("running", [])))
            return self._send(200, {"version": v})
        self._send(404, {"error": "not found"})


class JobServer:
    def __init__(self, node_ips, pod_num_of_node=1, gpu_num_of_node=8, time_interval_to_change=900, port=8180,
                 host="0.0.0.0", min_running=1, seed=0):
        self.state = JobState(node_ips, pod_num_of_node, gpu_num_of_node, min_running, seed)
        self.interval = time_interval_to_change
        self._httpd = ThreadingHTTPServer((host, port), _Handler)
        self._httpd.state = self.state
        self.port = self._httpd.server_address[1]
        self._stop = threading.Event()

    def start(self):
        threading.Thread(target=self._httpd.serve_forever, kwargs={"poll_interval": 0.2}, daemon=True).start()
        if self.interval and self.interval > 0:
            threading.Thread(target=self._flipper, daemon=True).start()
        logger.info("job server on port %d, %d pods, change every %ss", self.port, len(self.state.pods), self.interval)
        return self

    def _flipper(self):
        while not self._stop.wait(self.interval):
            self.state.flip()

    def stop(self):
        self._stop.set()
        self._httpd.shutdown()
        self._httpd.server_close()

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()


def main(argv=None):
    ap = argparse.ArgumentParser(description="EDL JobServer demo")
    ap.add_argument("--node_ips", type=str, required=True, help="comma separated node ips")
    ap.add_argument("--pod_num_of_node", type=int, default=1)
    ap.add_argument("--time_interval_to_change", type=int, default=900)
    ap.add_argument("--gpu_num_of_node", type=int, default=8)
    ap.add_argument("--port", type=int, default=8180)
    ap.add_argument("--log_level", type=int, default=20)
    args = ap.parse_args(argv)
    logging.basicConfig(level=args.log_level)
    srv = JobServer(args.node_ips.split(","), args.pod_num_of_node, args.gpu_num_of_node,
                    args.time_interval_to_change, args.port).start()
    try:
        threading.Event().wait()
    except KeyboardInterrupt:
        srv.stop()


if __name__ == "__main__":
    main()
