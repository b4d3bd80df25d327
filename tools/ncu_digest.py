"""Digest of an ncu report: headline metrics, per-source-line instruction counts, opcode mix.
usage: python tools/ncu_digest.py gpurun_out/prof.ncu-rep [n_qps]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; nq = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h:95s} {v} {u}")  # ncu scales units per metric (Mbyte / Kbyte, us / ms): print them
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[2]; iI = h.index("Instructions Executed"); iS = h.index("# Samples")
cur = None; srcl = {}; per = collections.Counter(); samp = collections.Counter(); ops = collections.Counter()
for r in rows[3:]:
    if len(r) <= iI: continue
    if r[2] == "-" and r[0].isdigit():
        cur = int(r[0]); srcl[cur] = r[1].strip(); continue
    try: inst = int(r[iI]); s = int(r[iS])
    except ValueError: continue
    per[cur] += inst; samp[cur] += s
    t = r[3].split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    ops[op.split(".")[0]] += inst
tot = sum(per.values()); ts = sum(samp.values())
print(f"\ntotal warp instructions / QP: {tot/nq:.1f}   (samples {ts})")
print("line   inst/QP  stall%  source")
for ln in sorted(k for k in per if k is not None):
    if per[ln] / nq >= 12:
        print(f"{ln:4d} {per[ln]/nq:8.1f} {100*samp[ln]/max(ts,1):6.1f}  {srcl.get(ln,'')[:105]}")
print("\nopcode mix (inst/QP):", ", ".join(f"{o} {c/nq:.0f}" for o, c in ops.most_common(24)))
