"""Teacher serving configuration: which feed variables the teacher consumes (name + shape) and
which outputs it can fetch.  Accepts JSON (``{"feed": [{"name", "shape", "dtype"}], "fetch": [...]}``)
or the Paddle-Serving ``serving_client_conf.prototxt`` text format the reference reads through
``paddle_serving_client.Client.load_client_config`` (distill_worker.py:213-241;
tests/unittests/serving_conf/serving_client_conf.prototxt)."""
import json
import re

_FEED_TYPE = {0: "int64", 1: "float32", 2: "int32"}


class ServingConf:
    def __init__(self, feeds=None, fetches=None):
        self.feeds = feeds or []      # [{"name", "shape", "dtype"}]
        self.fetches = fetches or []  # [{"name", "shape", "dtype"}]

    @property
    def feed_names(self):
        return [f["name"] for f in self.feeds]

    @property
    def fetch_names(self):
        return [f["name"] for f in self.fetches]

    def feed_shape(self, name):
        for f in self.feeds:
            if f["name"] == name:
                return list(f.get("shape") or [])
        return None

    def to_json(self):
        return json.dumps({"feed": self.feeds, "fetch": self.fetches})


def _parse_prototxt(text):
    conf = ServingConf()
    for kind, body in re.findall(r"(feed_var|fetch_var)\s*\{(.*?)\}", text, flags=re.S):
        name = re.search(r'alias_name\s*:\s*"([^"]+)"', body) or re.search(r'name\s*:\s*"([^"]+)"', body)
        shape = [int(x) for x in re.findall(r"shape\s*:\s*(-?\d+)", body)]
        t = re.search(r"(?:feed_type|fetch_type)\s*:\s*(\d+)", body)
        rec = {"name": name.group(1), "shape": shape, "dtype": _FEED_TYPE.get(int(t.group(1)) if t else 1, "float32")}
        (conf.feeds if kind == "feed_var" else conf.fetches).append(rec)
    return conf


def load_serving_conf(path):
    with open(path, "r") as f:
        text = f.read()
    s = text.lstrip()
    if s.startswith("{"):
        d = json.loads(s)
        return ServingConf(d.get("feed", []), d.get("fetch", []))
    return _parse_prototxt(text)
