"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
  python tools/summarize_profiles.py <round-tag> <launches.csv> [<report.ncu-rep> ...]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    return agg


def ncu_raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_dmma.sum",
            "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum"]
    res = collections.OrderedDict()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        if name in res:
            continue
        d = {}
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                d[w] = f"{r[i]} {units[i]}".strip()
        res[name] = d
    return res


def write_launches(tag, lpath, ci, suffix=""):
    agg = launches(lpath)
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches{suffix}.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), one local BA call on config {ci}\n\n")
        f.write(f"Command: `PPO_BA_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv python tools/profile_one.py {ci}`\n")
        f.write("(host-driven LM loop: the same kernels as the captured graph, one launch per node).\n")
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"total kernel time {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches\n\n| kernel | launches | total us | avg us | max us | share |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {v[2]:.2f} | {100 * v[1] / tot:.1f}% |\n")


def main():
    """python tools/summarize_profiles.py <tag> <launches.csv> [<launches_config4.csv>] [<report.ncu-rep> ...]
    A report whose name contains `config4` feeds the config4 entry of <tag>_traffic.json, every other one the config2 entry."""
    tag, lpath, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    write_launches(tag, lpath, 2)
    if reps and reps[0].endswith(".csv"):
        write_launches(tag, reps[0], 4, "_config4")
        reps = reps[1:]
    traffic_all = {}
    for rp in reps:
        traffic = traffic_all.setdefault("config4" if "config4" in os.path.basename(rp) else "config2", {})
        res = ncu_raw(rp)
        base = os.path.splitext(os.path.basename(rp))[0]
        if base.startswith(tag + "_"):
            base = base[len(tag) + 1:]
        for k, d in res.items():  # dram__bytes_read.sum + dram__bytes_write.sum per launch, in bytes (bench.py: roofline.traffic)
            try:
                tot_b = 0.0
                for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v, u = d[m].split()
                    tot_b += float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                traffic[re.sub(r"<.*", "", k.replace("ppo::", ""))] = tot_b
            except Exception:
                pass
        with open(os.path.join(ROOT, "profiles", f"{tag}_{base}.md"), "w") as f:
            f.write(f"# {tag}: ncu --set full summary from {os.path.basename(rp)} (first launch of each kernel)\n\n")
            for k, d in res.items():
                f.write(f"## {k}\n\n")
                for m, v in d.items():
                    f.write(f"- `{m}` = {v}\n")
                f.write("\n")
        json.dump(res, open(os.path.join(ROOT, "profiles", f"{tag}_{base}.json"), "w"), indent=1)
    if traffic_all:
        json.dump(traffic_all, open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
