"""Turn ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep  > profiles/rNN_<kernel>.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.3f} ms summed device time "
          "(ncu: cold-cache, serialised -- compare SHARES)")
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'share %':>8s} {'avg us':>9s} {'max us':>9s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {a[0]:8d} {a[1] / 1e3:10.3f} {100 * a[1] / tot:8.1f} {a[1] / a[0]:9.1f} {a[2]:9.1f}")


def report(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"== {d.get('Kernel Name', '?')}   grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}")
        for h, u in zip(hdr, units):
            if h in KEYS:
                print(f"   {h:86s} {d[h]:>16s} {u}")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
