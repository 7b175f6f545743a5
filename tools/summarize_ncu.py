"""Summarise ncu outputs brought back in gpurun_out/ into tracked files under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python tools/summarize_ncu.py full     gpurun_out/prof_r1a.ncu-rep profiles/r1_kernels_full.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        k = f'{row["Kernel Name"][:110]} grid={row["Grid Size"]}'
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): gpu__time_duration.sum, --clock-control none\n\n")
        f.write("Per-launch times are cold-cache and serialised by the profiler: compare SHARES.\n\n")
        f.write(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n\n| us | n | avg us | share | kernel |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {v[1]:.1f} | {v[0]} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% | `{k}` |\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src} (--clock-control none)\n\n")
        for r in data:
            f.write(f'## `{r[idx["Kernel Name"]][:120]}` grid={r[idx["Grid Size"]]} block={r[idx["Block Size"]]}\n\n')
            for k in KEYS:
                if k in idx:
                    f.write(f"- {k}: {r[idx[k]]} {units[idx[k]]}\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
