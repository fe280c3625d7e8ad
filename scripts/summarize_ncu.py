"""Summarise ncu outputs into markdown for profiles/ (run here, no GPU needed).

  python scripts/summarize_ncu.py launches <csv> <title>       per-kernel launch list
  python scripts/summarize_ncu.py full <ncu-rep> <title>       key metrics of a --set full capture
"""
import collections
import csv
import subprocess
import sys


def launches(path, title):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}\n")
    print("Per-launch times are cold-cache and serialised (ncu replays every kernel): compare "
          "SHARES with bench.py's CUDA-event `kernel_ms_per_unit`, not absolutes.\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:72]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |")


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_op_dmma.sum",
        "sm__inst_executed_pipe_lsu.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def full(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n")
    print(f"`ncu --set full --clock-control none --import-source on`, report `{path}` "
          "(kept out of git; numbers below are copied from it).\n")
    for r in rows[2:]:
        print(f"## `{r[hdr.index('Kernel Name')]}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"| {w} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
