#!/usr/bin/env python
"""Turn the ncu artefacts a GPU pass left in gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summary.py --tag r01 [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof_pa_v1.ncu-rep]

Writes profiles/<tag>_launches.md (per-kernel count / mean / share of the step from the
`--metrics gpu__time_duration.sum` pass), profiles/<tag>_pa_v1_metrics.csv (selected metrics of the
`--set full` capture of the attention kernel, one column per captured launch) and profiles/traffic.json
(dram bytes per attention launch, read by bench.py for roofline.traffic).
"""
import argparse
import collections
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def launches(path, out_md, cmd, step_kernels=("pa_v1_kernel", "reshape_and_cache")):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg.setdefault(row["Kernel Name"], []).append(v)
    step_tot = sum(sum(v) for k, v in agg.items() if any(s in k for s in step_kernels))
    with open(out_md, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n"
                "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"Command: `{cmd}`\n\n")
        f.write("| kernel | launches | mean us | total us | share of hot-path kernel time |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            mine = any(s in k for s in step_kernels)
            share = f"{sum(v) / step_tot * 100:.1f}%" if mine else "(synthetic input generation, outside the step)"
            f.write(f"| `{k[:100]}` | {len(v)} | {sum(v) / len(v):.2f} | {sum(v):.1f} | {share} |\n")
    return agg


def full(rep, out_csv):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        for m in KEEP:
            if m in hdr:
                i = hdr.index(m)
                w.writerow([m, units[i]] + [r[i] for r in data])

    def col(name):
        i = hdr.index(name)
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
        return [float(r[i].replace(",", "")) * scale for r in data]

    rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    return sum(rd) / len(rd), sum(wr) / len(wr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--launches", default=os.path.join(ROOT, "gpurun_out", "launches.csv"))
    ap.add_argument("--rep", default=os.path.join(ROOT, "gpurun_out", "prof_pa_v1.ncu-rep"))
    ap.add_argument("--cmd", default="ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv "
                                     "python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if os.path.isfile(a.launches):
        launches(a.launches, os.path.join(ROOT, "profiles", f"{a.tag}_launches.md"), a.cmd)
    if os.path.isfile(a.rep):
        rd, wr = full(a.rep, os.path.join(ROOT, "profiles", f"{a.tag}_pa_v1_metrics.csv"))
        try:
            commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True,
                                    text=True).stdout.strip() or "unknown"
        except Exception:
            commit = "unknown"
        import sys
        sys.path.insert(0, ROOT)
        from vllmini_b200 import _native
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump({"pa_v1_dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "source": f"profiles/{a.tag}_pa_v1_metrics.csv (ncu --set full, mean of captured launches)",
                       "kernel": _native.plan(256, 12, 12, 64, 16, 1024).split(" grid")[0],
                       "commit": f"{commit} (+ working tree at capture time)",
                       "workload": "bs=256 ctx=1024 12x64 block 16"}, f, indent=1)
        print("traffic per launch: read %.1f MB write %.1f MB" % (rd / 1e6, wr / 1e6))


if __name__ == "__main__":
    main()
