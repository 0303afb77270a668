"""Summarise one kernel of an `ncu --set full` report for profiles/ (markdown table + raw CSV).

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_x "title" ["command line"]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and keeps the metrics
B200_PROFILING.md names: duration, clocks, DRAM bytes and throughput, pipe utilisation, issue
rate, occupancy and the top warp-stall reasons.
"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.per_cycle_active",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
]


def main() -> None:
    report, out_prefix, title = sys.argv[1], sys.argv[2], sys.argv[3]
    command = sys.argv[4] if len(sys.argv) > 4 else ""
    if report.endswith(".csv"):       # the raw page, already exported on the GPU box (reports are ~18 MB each)
        raw = open(report).read()
    else:
        raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True,
                             check=True).stdout
    open(out_prefix + "_ncu_raw.csv", "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    header, units, values = rows[0], rows[1], rows[2]
    column = {name: i for i, name in enumerate(header)}
    lines = [f"# {title}", ""]
    if command:
        lines += ["Command (after the same command exited 0 without ncu):", "", "```", command, "```", ""]
    lines += [f"Kernel: `{values[column['Kernel Name']]}`. Raw page: `{out_prefix.split('/')[-1]}_ncu_raw.csv`.",
              "Numbers under the profiler are cold-cache and serialised; bench values come from CUDA events "
              "without ncu.", "", "| metric | value | unit |", "|---|---|---|"]
    for name in KEEP:
        if name in column:
            lines.append(f"| `{name}` | {values[column[name]]} | {units[column[name]]} |")
    stalls = []
    for name, i in column.items():
        if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(values[i].replace(",", "")), name))
            except ValueError:
                pass
    lines += ["", "Warp stall reasons (warps stalled per issue-active cycle, top 6):", "",
              "| reason | ratio |", "|---|---|"]
    for value, name in sorted(stalls, reverse=True)[:6]:
        short = name[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
        lines.append(f"| {short} | {value:.3f} |")
    open(out_prefix + "_summary.md", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
