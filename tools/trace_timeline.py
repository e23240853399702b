#!/usr/bin/env python
"""Per-stream occupancy and idle-gap analysis of a chrome trace written by ``bench.py --kineto P``
(``P.trace.json``): how busy is each CUDA stream during one graph replay, where does the main stream
wait, and which kernels sit on the longest dependency chain by time.

    python tools/trace_timeline.py gpurun_out/kineto.trace.json
"""
import collections
import json
import sys


def main(path):
    ev = json.load(open(path))
    ev = ev["traceEvents"] if isinstance(ev, dict) else ev
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and e.get("ph") == "X"]
    ks.sort(key=lambda e: e["ts"])
    if not ks:
        print("no kernels in trace")
        return
    t0, t1 = ks[0]["ts"], max(e["ts"] + e["dur"] for e in ks)
    streams = collections.defaultdict(list)
    for e in ks:
        streams[e.get("args", {}).get("stream", e.get("tid"))].append(e)
    print("trace span %.1f us, %d kernels, %d streams" % (t1 - t0, len(ks), len(streams)))
    for sid, lst in sorted(streams.items(), key=lambda kv: -len(kv[1])):
        busy = sum(e["dur"] for e in lst)
        span = lst[-1]["ts"] + lst[-1]["dur"] - lst[0]["ts"]
        print("stream %-6s kernels %5d  busy %9.1f us  span %9.1f us  (%.0f%% busy)" % (sid, len(lst), busy, span, 100.0 * busy / max(span, 1e-9)))
    main_sid = max(streams, key=lambda s: sum(e["dur"] for e in streams[s]))
    lst = streams[main_sid]
    gaps = []
    for a, b in zip(lst[:-1], lst[1:]):
        g = b["ts"] - (a["ts"] + a["dur"])
        gaps.append((g, a["name"][:60], b["name"][:60]))
    tot_gap = sum(g for g, _, _ in gaps if g > 0)
    print("main stream %s: total idle between kernels %.1f us over %d gaps; mean %.2f us" % (main_sid, tot_gap, len(gaps), tot_gap / max(1, len(gaps))))
    by_next = collections.defaultdict(lambda: [0.0, 0])
    for g, a, b in gaps:
        if g > 0:
            by_next[b][0] += g
            by_next[b][1] += 1
    print("idle time attributed to the kernel that FOLLOWS the gap (top 15):")
    for name, (g, n) in sorted(by_next.items(), key=lambda kv: -kv[1][0])[:15]:
        print("  %9.1f us in %4d gaps (%.1f us avg) before %s" % (g, n, g / n, name))
    print("largest single gaps:")
    for g, a, b in sorted(gaps, reverse=True)[:10]:
        print("  %7.1f us  after %-60s before %s" % (g, a, b))
    # concurrency: time during which >= 2 kernels are resident
    pts = []
    for e in ks:
        pts.append((e["ts"], 1))
        pts.append((e["ts"] + e["dur"], -1))
    pts.sort()
    lvl, last, hist = 0, pts[0][0], collections.defaultdict(float)
    for t, d in pts:
        hist[min(lvl, 3)] += t - last
        last = t
        lvl += d
    print("time with N kernels in flight: " + ", ".join("%d%s: %.0f us" % (k, "+" if k == 3 else "", v) for k, v in sorted(hist.items())))


if __name__ == "__main__":
    main(sys.argv[1])
