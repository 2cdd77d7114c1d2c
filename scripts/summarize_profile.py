#!/usr/bin/env python
"""Turn one `scripts/gpu_profile.sh <tag>` pass (files under gpurun_out/) into the tracked evidence under
profiles/: the bench lines, the ncu launch list, the per-launch DRAM traffic of the hot kernels from the
`--set full` capture, and a summary table.  usage: python scripts/summarize_profile.py <tag>"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
           "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "pcie__read_bytes.sum.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "nvlrx__bytes.sum.per_second", "nvltx__bytes.sum.per_second"]


def short(name):
    name = name.replace("void ", "")
    if "unnamed>::" in name:
        name = name.split("unnamed>::", 1)[1]
    return name.split("(")[0]


def launch_list(tag, suffix=""):
    path = os.path.join(SRC, f"launches_{tag}{suffix}.csv")
    rows = []
    with open(path) as f:
        text = f.read()
    start = text.index('"ID"')
    for r in csv.DictReader(io.StringIO(text[start:])):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]))
    agg = collections.OrderedDict()
    for k, us, g, b in rows:
        a = agg.setdefault(k, {"n": 0, "us": 0.0, "grid": g, "block": b})
        a["n"] += 1
        a["us"] += us
    total = sum(a["us"] for a in agg.values())
    return agg, total


def full_capture(tag, suffix=""):
    rep = os.path.join(SRC, f"hot_{tag}{suffix}.ncu-rep")
    if not os.path.exists(rep):
        return []
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    start = out.index('"ID"')
    rd = list(csv.reader(io.StringIO(out[start:])))
    head, units, body = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(head)}
    res = []
    for r in body:
        d = {"kernel": short(r[col["Kernel Name"]])}
        for m in METRICS:
            if m in col:
                d[m] = (r[col[m]], units[col[m]])
        res.append(d)
    return res


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(val, unit):
    v = float(val.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def main():
    tag = sys.argv[1]
    os.makedirs(DST, exist_ok=True)
    lines = [f"# profiles/{tag} — measured pass on B200", ""]
    lines.append(f"All numbers from one `gpurun` call (`bash scripts/gpu_profile.sh {tag}`). Raw files: `launches_{tag}.csv` "
                 f"(ncu launch list), `bench_{tag}*.json` (bench lines, NOT under a profiler), `ncu_traffic_{tag}.json` "
                 f"(per-launch DRAM bytes from the `--set full` capture). The `.ncu-rep` stays in `gpurun_out/` (scratch).")
    lines.append("")
    lines.append("## 1. bench.py lines (not under a profiler)\n")
    lines.append("| arm | value (M vectors/s) | ms/step | e2e (M vectors/s) | hit rate | probe kernel GB/s (frac of measured HBM peak) | host link GB/s (frac of memcpy) |")
    lines.append("|---|---|---|---|---|---|---|")
    for suffix, label in (("", "ours, direct pull (default `python bench.py`)"), ("_staged", "ours, staged CPU gather (`--miss-path staged`)"),
                          ("_reference", "CPU parameter-server path (`--impl reference`)")):
        p = os.path.join(SRC, f"bench_{tag}{suffix}.json")
        if not os.path.exists(p) or os.path.getsize(p) == 0:
            continue
        shutil.copy(p, os.path.join(DST, f"bench_{tag}{suffix}.json"))
        d = json.loads(open(p).read().strip().splitlines()[-1])
        rf, hl = d.get("roofline"), d.get("roofline_host_link")
        lines.append("| %s | %.1f | %.3f | %.1f | %s | %s | %s |" % (
            label, d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6,
            "%.4f" % d["config"]["hit_rate_measured"] if "hit_rate_measured" in d["config"] else "",
            "%.0f (%.3f)" % (rf["achieved"], rf["frac"]) if rf else "",
            "%.1f (%.3f)" % (hl["achieved"], hl["frac"]) if hl else ""))
        if suffix == "":
            ch = d.get("cache_hit")
            if ch:
                lines.append("| 100 %% cache-hit pass | %.1f | %.3f | | 1.0 | %.0f (%.3f) | |" % (
                    ch["vectors_per_s"] / 1e6, ch["ms_per_step"], ch["hbm_gbs"], ch["frac_of_peak"]))
            lines.append("")
            lines.append(f"<!-- clocks: {d.get('clocks')} ; cpu_baseline: {d.get('cpu_baseline')} -->")
            lines.pop(-2)
    lines.append("")
    agg, total = launch_list(tag)
    shutil.copy(os.path.join(SRC, f"launches_{tag}.csv"), os.path.join(DST, f"launches_{tag}.csv"))
    lines.append("## 2. ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised)\n")
    lines.append("| kernel | launches | total us | avg us | share of all kernel time | grid | block |")
    lines.append("|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        lines.append("| `%s` | %d | %.1f | %.1f | %.1f %% | %s | %s |" % (k, a["n"], a["us"], a["us"] / a["n"], 100 * a["us"] / total,
                                                                       a["grid"], a["block"]))
    lines.append("")
    caps = full_capture(tag)
    traffic = {"source": f"ncu --set full --clock-control none, gpurun_out/hot_{tag}.ncu-rep, per launch"}
    if caps:
        lines.append("## 3. `ncu --set full` of the hot kernels (per launch)\n")
        for c in caps:
            lines.append(f"**{c['kernel']}**\n")
            lines.append("| metric | value | unit |")
            lines.append("|---|---|---|")
            for m in METRICS:
                if m in c:
                    lines.append(f"| `{m}` | {c[m][0]} | {c[m][1]} |")
            lines.append("")
            key = c["kernel"].split("<")[0].replace("_kernel", "")
            if key not in traffic and "dram__bytes_read.sum" in c:
                traffic[key] = {"dram_bytes_read": int(to_bytes(*c["dram__bytes_read.sum"])),
                                "dram_bytes_write": int(to_bytes(*c["dram__bytes_write.sum"])),
                                "gpu_time_us": to_us(*c["gpu__time_duration.sum"])}
        with open(os.path.join(DST, f"ncu_traffic_{tag}.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    # extra passes: the single-GPU form of the model-parallel workload and the dense head
    for suffix, title in (("_c4", "model-parallel workload on one GPU (`bench.py --workload c4 --gpus 1`)"),
                          ("_tier", "NVLink tier with one rank: misses pulled from the local shard (`bench.py --value-only --local-tier`)"),
                          ("_gather", "tier-only table, 125 M rows, 1.7 M keys per request (`scripts/c4_repro.py`)"),
                          ("_mlp", "dense head, Criteo shape (`scripts/mlp_profile_driver.py`)")):
        if not os.path.exists(os.path.join(SRC, f"launches_{tag}{suffix}.csv")):
            for c in full_capture(tag, suffix):
                lines.append(f"**{c['kernel']}** (`ncu --set full`) — {title}\n")
                lines.append("| metric | value | unit |")
                lines.append("|---|---|---|")
                for m in METRICS:
                    if m in c:
                        lines.append(f"| `{m}` | {c[m][0]} | {c[m][1]} |")
                lines.append("")
            continue
        agg2, total2 = launch_list(tag, suffix)
        shutil.copy(os.path.join(SRC, f"launches_{tag}{suffix}.csv"), os.path.join(DST, f"launches_{tag}{suffix}.csv"))
        lines.append(f"## ncu launch list — {title}\n")
        lines.append("| kernel | launches | total us | avg us | share of all kernel time | grid | block |")
        lines.append("|---|---|---|---|---|---|---|")
        for k, a in sorted(agg2.items(), key=lambda kv: -kv[1]["us"]):
            lines.append("| `%s` | %d | %.1f | %.1f | %.1f %% | %s | %s |" % (k, a["n"], a["us"], a["us"] / a["n"], 100 * a["us"] / total2,
                                                                           a["grid"], a["block"]))
        lines.append("")
        for c in full_capture(tag, suffix):
            lines.append(f"**{c['kernel']}** (`ncu --set full`)\n")
            lines.append("| metric | value | unit |")
            lines.append("|---|---|---|")
            for m in METRICS:
                if m in c:
                    lines.append(f"| `{m}` | {c[m][0]} | {c[m][1]} |")
            lines.append("")
    with open(os.path.join(DST, f"{tag}_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
