"""Digest of an ncu report: headline metrics, executed warp instructions and stall samples per source FUNCTION
(source files are found from the report's own paths; functions by their definition lines), opcode mix.
usage: python tools/ncu_digest.py gpurun_out/prof.ncu-rep [n_qps]"""
import collections, csv, io, os, re, subprocess, sys

rep = sys.argv[1]
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h:95s} {v} {u}")  # ncu scales units per metric (Mbyte / Kbyte, us / ms): print them


def func_ranges(path):
    out = []
    if not os.path.exists(path):
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "quadruped_control_b200", "csrc", os.path.basename(path))
    if not os.path.exists(path):
        return out
    for i, l in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"^(?:QPB_HD|__device__ __forceinline__|__device__ __noinline__|inline|__global__)[^(]*?\b(\w+)\(", l) or re.match(r"^(\w+)\(const __grid_constant__", l)
        if m:
            out.append([m.group(1), i, None])
        if l.startswith("}") and out and out[-1][2] is None:
            out[-1][2] = i
    return out


src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
per = collections.Counter(); samp = collections.Counter(); ops = collections.Counter(); thr = collections.Counter()
line_inst = collections.Counter(); line_samp = collections.Counter(); line_src = {}
fpath = None; ranges = {}; iI = iS = iT = None; cur = None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]; ranges[fpath] = func_ranges(fpath); continue
    if r[0] == "Line No":
        iI = r.index("Instructions Executed"); iS = r.index("# Samples"); iT = r.index("Thread Instructions Executed"); continue
    if r[0] == "Function Name" or iI is None or len(r) <= iI:
        continue
    if r[0].isdigit() and r[2] == "-":
        cur = int(r[0]); line_src[(fpath, cur)] = r[1].strip(); continue
    try:
        inst = int(r[iI]); s = int(r[iS]); t = int(r[iT])
    except ValueError:
        continue
    fn = os.path.basename(fpath or "?")
    for name, a, b in ranges.get(fpath, []):
        if b and cur is not None and a <= cur <= b:
            fn += ":" + name
    per[fn] += inst; samp[fn] += s; thr[fn] += t
    line_inst[(fpath, cur)] += inst; line_samp[(fpath, cur)] += s
    tk = r[3].split()
    op = tk[1] if tk and tk[0].startswith("@") else (tk[0] if tk else "?")
    ops[op.split(".")[0]] += inst
tot = sum(per.values()); ts = sum(samp.values())
# (tot sums the source view, which lists an instruction of an inlined function under every file of its inline stack that has
# line information -- 5-20 % above the hardware counter smsp__inst_executed.sum printed with the headline metrics, which is the
# figure the documents quote; shares per function are what this table is for)
print(f"\nwarp instructions executed / QP: {tot/nq:.1f} summed over the source view (hardware count: smsp__inst_executed.sum / QPs, above)   (stall samples {ts})")
print("  inst/QP  inst%  stall%  lanes  function")
for fn, c in per.most_common(40):
    if c / nq >= 1.0:
        print(f"{c/nq:9.1f} {100*c/tot:6.1f} {100*samp[fn]/max(ts,1):7.1f} {thr[fn]/max(c,1):6.1f}  {fn}")
print("\nhottest source lines (stall samples):")
for (f, ln), s in line_samp.most_common(14):
    print(f"{100*s/max(ts,1):6.1f}% {line_inst[(f,ln)]/nq:8.1f} inst/QP  {os.path.basename(f)}:{ln}  {line_src.get((f,ln),'')[:90]}")
print("\nopcode mix (inst/QP):", ", ".join(f"{o} {c/nq:.1f}" for o, c in ops.most_common(28)))
