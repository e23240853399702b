"""Feed / fetch description of a teacher's serving conf (reference: example/distill/qps_tools/parse_config.py:19-44,
which asks paddle_serving_client for it).  Here the file (JSON or the Paddle-Serving prototxt) is parsed by
``paddle_edl.distill.serving_conf``."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from paddle_edl.distill.serving_conf import load_serving_conf  # noqa: E402


def get_ins_predicts(conf_file=None):
    """-> (feed names, feed shapes, feed dtypes, fetch names).  Search order of the reference: the argument,
    ./serving_conf/serving_client_conf.prototxt, $PADDLE_DISTILL_CONF_FILE."""
    cands = [conf_file, "./serving_conf/serving_client_conf.prototxt", os.getenv("PADDLE_DISTILL_CONF_FILE")]
    path = next((c for c in cands if c and os.path.isfile(c)), None)
    assert path is not None, "no serving conf file: pass one or set PADDLE_DISTILL_CONF_FILE"
    conf = load_serving_conf(path)
    return (conf.feed_names, [tuple(f.get("shape") or ()) for f in conf.feeds],
            [f.get("dtype", "float32") for f in conf.feeds], conf.fetch_names)


if __name__ == "__main__":
    print(get_ins_predicts(sys.argv[1] if len(sys.argv) > 1 else None))
