#!/usr/bin/env python
"""Turn the raw artefacts that come back in gpurun_out/ (ncu reports, launch lists, bench JSON, kernel
microbench logs) into the small committed summaries under profiles/.

    python tools/summarize_profiles.py
"""
import collections
import csv
import glob
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def kernel_key(name):
    m = re.search(r"edl::(?:<unnamed>::)?(\w+)", name)
    if m:
        t = re.search(r"<([^>]*)>", name[m.end():m.end() + 80])
        return "edl::" + m.group(1) + (("<" + t.group(1) + ">") if t else "")
    return re.sub(r"<.*", "", name)[:60]


def ncu_report(path):
    try:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    except FileNotFoundError:
        return None
    rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
    if len(rows) < 3:
        return None
    hdr, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = {"kernel": kernel_key(row[hdr.index("Kernel Name")]), "grid": row[hdr.index("Grid Size")],
             "block": row[hdr.index("Block Size")]}
        for m in RAW_METRICS:
            if m in hdr:
                d[m] = "%s %s" % (row[hdr.index(m)], units[hdr.index(m)])
        res.append(d)
    return res


def launch_list(path):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    rows = collections.OrderedDict()
    for row in r:
        d = rows.setdefault(int(row[idx["ID"]]), {"name": row[idx["Kernel Name"]]})
        d[row[idx["Metric Name"]]] = float(row[idx["Metric Value"]].replace(",", ""))
    ids = list(rows)
    sgd = [i for i in ids if "sgd_momentum" in rows[i]["name"]]
    if len(sgd) < 3:
        return None
    step = [rows[i] for i in ids if sgd[-3] < i <= sgd[-1]]
    tot = sum(x["gpu__time_duration.sum"] for x in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in step:
        a = agg[kernel_key(x["name"])]
        a[0] += 1
        a[1] += x["gpu__time_duration.sum"]
    lines = ["last training step: %d kernel launches, %.1f us summed kernel time (ncu serialised, caches flushed per kernel)" % (
        len(step), tot / 1000), "", "%6s %10s %7s %9s  kernel" % ("count", "total us", "share", "avg us")]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        lines.append("%6d %10.1f %6.1f%% %9.1f  %s" % (c, t / 1000, 100 * t / tot, t / 1000 / c, k))
    return "\n".join(lines)


def main():
    os.makedirs(OUT, exist_ok=True)
    for rep in sorted(glob.glob(os.path.join(SRC, "*.ncu-rep"))):
        res = ncu_report(rep)
        if res:
            with open(os.path.join(OUT, os.path.basename(rep).replace(".ncu-rep", ".ncu.txt")), "w") as f:
                f.write("# ncu --set full --clock-control none, from %s\n" % os.path.basename(rep))
                for d in res:
                    f.write("\n")
                    for k, v in d.items():
                        f.write("%-78s %s\n" % (k, v))
    for ll in sorted(glob.glob(os.path.join(SRC, "launches*.csv"))):
        s = launch_list(ll)
        if s:
            open(os.path.join(OUT, os.path.basename(ll).replace(".csv", ".txt")), "w").write(s + "\n")
    for f in sorted(glob.glob(os.path.join(SRC, "kernels*.log")) + glob.glob(os.path.join(SRC, "kineto*.txt"))):
        txt = "".join(l for l in open(f) if not l.startswith("W0"))
        open(os.path.join(OUT, os.path.basename(f).replace(".log", ".txt")), "w").write(txt)
    # earlier rounds' records stay: gpurun_out/ is scratch and starts empty in a re-created container
    bench_path = os.path.join(OUT, "bench_runs.json")
    try:
        bench = json.load(open(bench_path))
    except (OSError, ValueError):
        bench = {}
    pats = ("bench*.json", "ab_*.json", "ab2_*.json", "b4_*.json", "b6_*.json", "b7_*.json", "b8_*.json", "b9_*.json",
            "b10_*.json", "b1[1-9]_*.json", "b[2-9][0-9]_*.json", "final_bench_*.json")
    for f in sorted(x for p_ in pats for x in glob.glob(os.path.join(SRC, p_))):
        try:
            line = [l for l in open(f).read().splitlines() if l.startswith("{")][-1]
            d = json.loads(line)
            bench[os.path.basename(f)[:-5]] = {k: d.get(k) for k in ("impl", "n_gpus", "value", "unit", "ms_per_step", "gpu_launches", "clocks")}
            bench[os.path.basename(f)[:-5]]["e2e"] = (d.get("e2e") or {}).get("value")
            keys = ("model", "batch_per_gpu", "cuda_graph", "conv_impl", "allreduce", "parallelism", "own_wgrad3",
                    "fuse_bn_bwd", "conv3_s2", "own_stem1", "no_library")
            bench[os.path.basename(f)[:-5]]["config"] = {k: (d.get("config") or {}).get(k) for k in keys}
            for extra in ("exposed_comm_ms", "library_fallbacks"):
                if d.get(extra) is not None:
                    bench[os.path.basename(f)[:-5]][extra] = d[extra]
            for extra in ("rescale", "distill", "allreduce", "exposed_comm"):        # multi-GPU metric terms (bench.py extras)
                if isinstance(d.get(extra), dict):
                    v = dict(d[extra])
                    v.pop("buckets", None)
                    bench[os.path.basename(f)[:-5]][extra] = v
        except Exception:
            pass
    json.dump(bench, open(bench_path, "w"), indent=1)
    # round 2: multi-GPU sweeps, elastic-launch recovery times, timelines, A/B summaries
    for pat, dst in (("comm_*gpu.json", None), ("rescale_*gpu.json", None), ("elastic_launch_*gpu*.json", None),
                     ("ctr_sweep_*gpu.json", None), ("ctr_deepfm_*gpu.json", None), ("prof_allreduce_*gpu.jsonl", None),
                     ("timeline_r2*.txt", None), ("call*_summary.txt", None), ("wgrad3.json", "wgrad3_microbench.json"),
                     ("trace_persist_c*.txt", None), ("teacher_c*.txt", None), ("final_*tests.log", None),
                     ("final_smoke.log", None)):
        for f in sorted(glob.glob(os.path.join(SRC, pat))):
            open(os.path.join(OUT, dst or os.path.basename(f)), "w").write(open(f).read())
    for name in ("experimental_summary.txt", "loader_bench.jsonl"):       # scripts/gpu_validate_experimental.sh
        src = os.path.join(SRC, name)
        if os.path.exists(src):
            open(os.path.join(OUT, name), "w").write(open(src).read())
    print("wrote", len(os.listdir(OUT)), "files to", OUT)


if __name__ == "__main__":
    main()
