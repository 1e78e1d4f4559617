"""Text summary of an .ncu-rep: per-kernel headline metrics, opcode mix, hottest source lines.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex] > profiles/rN_summary.txt
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 else None
    sel = ["--kernel-name", "regex:" + rx] if rx else []
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"] + sel))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print("kernel:", r[hdr.index("Kernel Name")], " grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
        for mname in METRICS:
            if mname in hdr:
                i = hdr.index(mname)
                print(f"  {mname:90s} {r[i]:>16s} {units[i]}")
    src = run(["-i", rep, "--page", "source", "--csv"] + sel)
    rows = list(csv.reader(io.StringIO(src)))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    for a, b in zip(secs[:-1], secs[1:]):
        name = rows[a][1]
        hdr = rows[a + 1]
        body = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
        ci, si, st = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
        tot = sum(int(r[ci]) for r in body) or 1
        tots = sum(int(r[st]) for r in body) or 1
        print("=" * 100)
        print("kernel:", name, " executed warp instructions:", tot, " stall samples:", tots)
        op, ops = collections.Counter(), collections.Counter()
        for r in body:
            o = [x for x in r[si].split() if not x.startswith("@")]
            k = o[0].split(".")[0] if o else "?"
            op[k] += int(r[ci]); ops[k] += int(r[st])
        print("  opcode mix (share of executed warp instructions | share of stall samples)")
        for k, v in op.most_common(18):
            print(f"    {k:10s} {100 * v / tot:5.1f}%  {100 * ops[k] / tots:5.1f}%")
        print("  hottest SASS by stall samples")
        for r in sorted(body, key=lambda r: -int(r[st]))[:12]:
            print(f"    {100 * int(r[st]) / tots:5.1f}%  exec {int(r[ci]):>10d}  {r[si].strip()[:90]}")


if __name__ == "__main__":
    main()
