#!/usr/bin/env python
"""Turn the raw output of profiles/r02_profile_run.sh (gpurun_out/prof/) into the tracked summaries under profiles/:
r02_launches_{strict,fast}.csv (raw launch lists), r02_launches.md, r02_step_kernel_ncu.md."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "prof")
DST = os.path.join(ROOT, "profiles")
CELLS = 8192 * 2048
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
]


def launch_table(arith):
    rows = list(csv.reader(open(os.path.join(SRC, f"launches_{arith}.csv"))))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[hi + 1:]:
        try:
            agg[r[ik]].append(float(r[iv].replace(",", "")))
        except Exception:
            pass
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | mean us | share of profiled GPU time |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k.split('(')[0]}` | {len(v)} | {sum(v) / len(v) / 1000:.1f} | {sum(v) / tot * 100:.1f}% |")
    return "\n".join(out)


def raw_metrics(arith):
    out = subprocess.run(["ncu", "-i", os.path.join(SRC, f"prof_{arith}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    res = {}
    for m in METRICS:
        if m in h:
            i = h.index(m)
            res[m] = (units[i], [r[i] for r in data])
    res["Kernel Name"] = ("", [r[h.index("Kernel Name")] for r in data])
    res["Grid Size"] = ("", [r[h.index("Grid Size")] for r in data])
    res["Block Size"] = ("", [r[h.index("Block Size")] for r in data])
    return res


def mix(arith):
    out = subprocess.run(["ncu", "-i", os.path.join(SRC, f"prof_{arith}.ncu-rep"), "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    i_s, i_e = hdr.index("Source"), hdr.index("Instructions Executed")
    by, tot = collections.Counter(), 0
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        try:
            e = int(r[i_e])
        except Exception:
            continue
        op = r[i_s].split()
        by[(op[1] if op[0].startswith("@") else op[0]).split(".")[0]] += e
        tot += e
    threads = CELLS / 2 / 32
    return tot / threads, [(o, c / threads) for o, c in by.most_common(14)]


def live(arith):
    d = json.loads(open(os.path.join(SRC, f"plain_{arith}.json")).read().strip().splitlines()[-1])
    return d["ms_per_step"] * 1000, d["value"], d["roofline"]["frac"]


def main():
    for a in ("strict", "fast"):
        shutil.copy(os.path.join(SRC, f"launches_{a}.csv"), os.path.join(DST, f"r02_launches_{a}.csv"))
    with open(os.path.join(DST, "r02_launches.md"), "w") as f:
        f.write("# Round 2 -- ncu launch lists (`ncu --metrics gpu__time_duration.sum --clock-control none -c 400`), FINAL build\n\n"
                "Command: `python bench.py --quick --steps 200 --windows 3 --warmup 20 --arith A` (urban 8192x2048, register kernel, 2 cells per\n"
                "thread, 128-thread CTAs, ring warps, PDL + early start); script `profiles/r02_profile_run.sh`, raw lists\n"
                "`profiles/r02_launches_strict.csv` / `_fast.csv`.  Per-launch times under ncu are serialised and cold-cache: compare shares.\n")
        for a, title in (("strict", "default arithmetic (strict, bit-exact)"), ("fast", "optional fast arithmetic")):
            us, mlups, frac = live(a)
            f.write(f"\n## {title}\n\nLive (same command without ncu, CUDA events): **{us:.1f} us per step = {mlups:.0f} MLUPS = {frac * 100:.1f} % "
                    f"of the measured HBM roofline**.\n\n{launch_table(a)}\n")
        f.write("\nThe step is ONE kernel in both builds (99+ % of the profiled GPU time incl. its EMIT variant, the last step of every `lbm_run`);\n"
                "the first 400 launches cover init + warm-up + the start of the timed windows.\n")
    with open(os.path.join(DST, "r02_step_kernel_ncu.md"), "w") as f:
        f.write("# Round 2 -- `ncu --set full` of the dominant kernel (FINAL build), strict (default) and fast arithmetic\n\n"
                "Command: `ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 3 python bench.py --quick --steps 200\n"
                "--windows 3 --warmup 20 --arith A` after the plain run of the same command exited 0 (`profiles/r02_profile_run.sh`).\n"
                "Algorithmic bytes per launch: 72 B x 16 777 216 cells = 1 207.96 MB (+ 2.1 MB of mask bits).\n")
        for a in ("strict", "fast"):
            m = raw_metrics(a)
            us, mlups, frac = live(a)
            per_thread, top = mix(a)
            n = len(m["Kernel Name"][1])
            f.write(f"\n## arith = {a}: `{m['Kernel Name'][1][0]}`, grid {m['Grid Size'][1][0]}, block {m['Block Size'][1][0]}\n\n")
            f.write(f"Live: {us:.1f} us per step, {mlups:.0f} MLUPS, {frac * 100:.1f} % of the measured HBM copy peak (6 550.7 GB/s).\n\n")
            f.write("| metric | " + " | ".join(f"launch {i}" for i in range(n)) + " |\n|---|" + "---|" * n + "\n")
            for k in METRICS:
                if k in m:
                    f.write(f"| {k} [{m[k][0]}] | " + " | ".join(m[k][1]) + " |\n")
            rd = [float(v.replace(",", "")) for v in m["dram__bytes_read.sum"][1]]
            wr = [float(v.replace(",", "")) for v in m["dram__bytes_write.sum"][1]]
            unit = m["dram__bytes_read.sum"][0]
            f.write(f"\nDRAM traffic per launch: {sum(rd) / n:.1f} + {sum(wr) / n:.1f} = {(sum(rd) + sum(wr)) / n:.1f} {unit} "
                    f"(algorithmic 1 208 MB; the writes still dirty in the 126 MB L2 at kernel end are not counted).\n")
            f.write(f"Dynamic instruction mix (source page): {per_thread:.0f} warp instructions per thread (= per 2 cells): "
                    + ", ".join(f"{o} {c:.0f}" for o, c in top) + ".\n")
    print("wrote profiles/r02_launches.md, r02_step_kernel_ncu.md")


if __name__ == "__main__":
    sys.exit(main())
