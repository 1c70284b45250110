"""Summarise an .ncu-rep (read here, no GPU needed) into the small text files committed under
profiles/: per-kernel key metrics (raw page) and, for one kernel, the SASS opcode mix with the
stall-sample share (source page).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01 [kernel-regex-for-opcode-mix ...]
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def raw_summary(rep, out_prefix):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    h, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(h)}
    seen = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        seen.setdefault(name, []).append(r)
    with open(out_prefix + "_kernels.md", "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of `{rep}`\n\n")
        f.write("Per-launch values (first captured launch of each kernel; launches captured: "
                + ", ".join(f"{k} x{len(v)}" for k, v in seen.items()) + ").\n\n")
        for name, rs in seen.items():
            r = rs[0]
            f.write(f"## {name}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in ix:
                    f.write(f"| {k} | {r[ix[k]]} | {units[ix[k]]} |\n")
            rd = float(r[ix["dram__bytes_read.sum"]].replace(",", ""))
            wr = float(r[ix["dram__bytes_write.sum"]].replace(",", ""))
            f.write(f"| traffic = dram read + write | {rd + wr:.3f} | {units[ix['dram__bytes_read.sum']]} |\n\n")


def opcode_mix(rep, out_prefix, pattern):
    text = ncu(["-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pattern}"])
    rows = list(csv.reader(io.StringIO(text)))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    if not secs:
        return
    s = secs[0]
    h = s["rows"][0]
    data = [r for r in s["rows"][1:] if len(r) > 10]
    ix = {n: i for i, n in enumerate(h)}
    ops, ops_s = collections.Counter(), collections.Counter()
    tot_i = tot_s = 0
    bars, cum_s, cum_i = [], 0, 0
    for k, r in enumerate(data):
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
        base = (m.group(2) if m else r[ix["Source"]].split()[0]).split(".")[0]
        sm, n = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
        ops[base] += n
        ops_s[base] += sm
        tot_i += n
        tot_s += sm
        if base == "BAR":
            bars.append((k, cum_s, cum_i))
        cum_s += sm
        cum_i += n
    tag = re.sub(r"[^a-z0-9]+", "_", pattern.lower()).strip("_")
    with open(f"{out_prefix}_{tag}_opcodes.md", "w") as f:
        f.write(f"# SASS opcode mix of `{s['name']}`\n\nstatic instructions: {len(data)}; warp-instructions executed: {tot_i}; "
                f"stall samples: {tot_s}\n\n")
        for (k, cs, ci) in bars:
            f.write(f"- BAR.SYNC at static instruction {k}: {100 * ci / tot_i:.1f}% of executed instructions and "
                    f"{100 * cs / tot_s:.1f}% of samples lie before it\n")
        f.write("\n| opcode | warp-instructions | share | sample share |\n|---|---|---|---|\n")
        for op, n in ops.most_common(30):
            f.write(f"| {op} | {n} | {100 * n / tot_i:.2f}% | {100 * ops_s[op] / tot_s:.2f}% |\n")


if __name__ == "__main__":
    rep, prefix = sys.argv[1], sys.argv[2]
    raw_summary(rep, prefix)
    for pat in sys.argv[3:]:
        opcode_mix(rep, prefix, pat)
