"""Turn gpurun_out/launches.csv + prof_*.ncu-rep into committed text summaries under profiles/.

usage: python tools/summarize_ncu.py <tag>     (e.g. r01a)
"""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(tag):
    path = os.path.join(GO, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            agg[d["Kernel Name"].split("(")[0]].append(float(d["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu --no-big-sweep --no-config2 --no-config3 --no-config4`\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 500` (cold-cache, serialised: "
                "compare shares, not absolutes).\n\n| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[:70]}` | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / tot:.1%} |\n")


def kernels(tag):
    for rep in sorted(glob.glob(os.path.join(GO, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        with open(os.path.join(OUT, f"{tag}_{name}.md"), "w") as f:
            f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on -k regex:{name}` (bench.py)\n\n")
            for li, r in enumerate(rows[2:]):
                d = dict(zip(hdr, r))
                u = dict(zip(hdr, units))
                f.write(f"## launch {li}: {d.get('Kernel Name', '')[:120]}\n\n| metric | value | unit |\n|---|---:|---|\n")
                for k in WANT:
                    if k in d:
                        f.write(f"| {k} | {d[k]} | {u.get(k, '')} |\n")
                # warp stall reasons (cycles stalled per issued instruction), largest first
                stalls = []
                for k in hdr:
                    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and k in d:
                        try:
                            stalls.append((float(d[k].replace(",", "")), k))
                        except ValueError:
                            pass
                for v, k in sorted(stalls, reverse=True)[:8]:
                    f.write(f"| {k} | {v:.2f} | warps per issue |\n")
                f.write("\n")
            src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
            tmp = f"/tmp/src_{name}.csv"
            open(tmp, "w").write(src)
            mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_mix.py"), tmp], capture_output=True,
                                 text=True).stdout
            f.write("## SASS opcode mix and stall samples (all captured launches summed)\n\n```\n" + mix + "```\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    kernels(tag)
