#!/usr/bin/env python
"""Benchmark of the balance-controller hot path (BASELINE.json metric: balance QPs/sec, batched).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference CPU path (oracle/_ref) on the host cores

A "step" is one pass of the hot path over one batch of synthetic robot states: BASELINE config 2
(65 536 states, 4 feet in contact, mu = 0.6, seed 20260102) per GPU; weak scaling: rank r owns records
[r*65536, (r+1)*65536) of the same stream.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: whatever NCCL logs (the version banner when NCCL_DEBUG is set) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

from quadruped_control_b200 import (ALGO_BYTES_PER_QP, OUT_DTYPE, STATE_DTYPE, WIRE_OUT_DTYPE, WIRE_STATE_DTYPE,  # noqa: E402
                                    default_params, states, to_wire)
from quadruped_control_b200.sharding import reduce_report  # noqa: E402

WORKLOADS = {
    # name: (records per GPU, seed, masks, profile, description)
    "cfg2": (65536, 20260102, "all4", "default", "cfg2: 65536 random CoM poses/twists per GPU, 4 feet in contact, mu=0.6"),
    "cfg3": (1048576, 20260103, "mixed", "default", "cfg3: 1048576 mixed 2/3/4-foot contact states per GPU, mu=0.6"),
    "cfg5": (1048576, 20260105, "mixed", "default", "cfg5: 8388608 mixed-contact states over 8 GPUs (1048576 per GPU), mu=0.6"),
}
MU = 0.6
METRIC = "balance QPs/sec (batched)"
UNIT = "QP/s"
N_ROTATE = 8  # distinct device batches cycled through so a step never re-reads L2-resident inputs

# BASELINE config 4 (10-step convex-MPC QP, 120 variables): a secondary workload with its own
# metric.  The reference has no code for it, so its CPU arm is the oracle port (kind "port", parity unpinned).
MPC_N, MPC_SEED = 65536, 20260104
MPC_METRIC = "MPC QPs/sec (10-step horizon, 120 variables, batched)"
MPC_DESC = "cfg4: 65536 10-step convex-MPC QPs per GPU (stand/trot/crawl gaits mixed), mu=0.6"
FP64_PEAK_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # nominal; tools/ubench_fp64 measures 64.0 DFMA lanes/clk/SM


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_explanatory(workload):
    """What actually bounds the kernel, from the committed `ncu --set full` capture (profiles/ncu_summary.json): issue-slot
    and FP64-pipe utilisation.  Static evidence quoted beside the live HBM figure, not measured by this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            d = json.load(f).get(workload, {})
        keys = ("issue_slots_busy_pct", "fp64_pipe_pct", "dram_pct_of_peak", "source")
        return {k: d[k] for k in keys if k in d} or None
    except Exception:
        return None


def ncu_traffic_per_launch(workload):
    """dram bytes per launch from the committed `ncu --set full` summary, if present."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            d = json.load(f)
        return d.get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clock/throttle sampler running beside the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows = []
        self.proc = None
        self.dev = device_index
        self.t0 = None
        self.marks = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.dev)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t_begin - 0.05 <= t <= t_end + 0.15] or [r for _, r in self.rows[-3:]]
        if not rows:  # nvidia-smi had not delivered its first sample yet: one direct query
            try:
                q = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.dev)],
                                   capture_output=True, text=True, timeout=10)
                rows = [l.strip() for l in q.stdout.splitlines() if l.strip()]
            except Exception:
                rows = []
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _cpu_arm():
    """The CPU implementation that is timed: oracle/_ref (the reference's own balance_controller.cpp +
    kinematics.cpp compiled against stand-in Armadillo/qpOASES headers) when it was
    built, else the plain-C oracle port.  Test infrastructure, used here only as the timed baseline."""
    import oracle

    if oracle.ref_available():
        return "reference", oracle.ref_control_batch, ("oracle/_ref: reference sources balance_controller.cpp + kinematics.cpp, "
                                                        "stand-in Armadillo/qpOASES (QP solved by the oracle's Goldfarb-Idnani)")
    return "port", oracle.control_batch, "oracle port (plain C restatement)"


def cpu_baseline(params, S, budget_s=12.0, check=None):
    """The reference CPU path on the host cores over a bounded sample of the workload.  ``check`` = (states, GPU
    outputs): their max relative GRF error against the oracle is returned beside the baseline."""
    import oracle

    kind, fn, desc = _cpu_arm()
    cores = os.cpu_count() or 1
    fn(params, S[:2048], cores)  # warm
    t0 = time.perf_counter()
    fn(params, S[:8192], cores)
    rate = 8192 / (time.perf_counter() - t0)
    n = int(min(len(S), max(8192, rate * budget_s / 3)))
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        fn(params, S[:n], cores)
        best = max(best, n / (time.perf_counter() - t0))
    t0 = time.perf_counter()
    m = min(n, 8192)
    fn(params, S[:m], 1)
    one = m / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    oracle.control_batch(params, S[:n], cores)
    port = n / (time.perf_counter() - t0)
    err = None
    if check is not None:
        ref = oracle.control_batch(params, check[0], cores)
        err = float((np.abs(check[1]["grf_body"] - ref["grf_body"]).max(axis=1) / np.maximum(np.abs(ref["grf_body"]).max(axis=1), 1.0)).max())
    return {"value": best, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"first {n} records of the workload, best of 3, {cores} threads; {desc}; 1 thread: {one:.0f} QP/s",
            "one_thread": one, "c_port_all_cores": port}, err


def static_config(desc, n, profile):
    """The part of `config` both arms print identically (the driver compares the two dicts key by key)."""
    return {"workload": desc, "qps_per_step_per_gpu": n, "profile": profile, "mu": MU}


# FP64 work of the range-space path per QP, as a model (qpb_tpq_core.h): PD target + lever arms (~450 flops), one 6x6
# solve on a set of faces (~700: G from four legs, Cholesky, two triangular solves, per-leg projections and multipliers)
# for the starting pair and once more for the polish, ~650 per working-set change (block round or loop iteration:
# Cholesky + solves + rank-one update + per-leg step and multiplier directions), ~600 for the epilogue (12 sincos, J^T f).
def balance_algorithmic_flops(iters):
    return 450.0 + 2 * 700.0 + 600.0 + 650.0 * np.asarray(iters, dtype=np.float64)


FP64_PEAK_TFLOPS = 37.2  # DFMA rate measured with tools/ubench_fp64 on the pool's B200 (profiles/r01_ubench_fp64.txt)


def run_reference(args):
    """--impl reference: the reference's CPU path for this hot path on the host cores (see _cpu_arm), with all
    the host threads, on our arm's config.  Rank 0 only; other ranks exit without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, fn, desc = _cpu_arm()
    n, seed, masks, profile, wdesc = WORKLOADS[args.workload]
    if args.profile:
        profile = args.profile
        wdesc += f" ({profile} disturbance profile)"
    params = default_params(MU)
    cores = os.cpu_count() or 1
    S_full = states.generate_states(n, seed, profile=profile, masks=masks)
    fn(params, S_full[:4096], cores)
    t0 = time.perf_counter()
    fn(params, S_full[:8192], cores)
    rate = 8192 / (time.perf_counter() - t0)
    total_steps = args.steps + args.warmup
    per_step = int(min(n, max(4096, rate * 120.0 / total_steps)))  # whole run within ~2 minutes
    S = np.ascontiguousarray(S_full[:per_step])
    for _ in range(args.warmup):
        fn(params, S, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = fn(params, S, cores)
    dt = time.perf_counter() - t0
    assert (out["status"] == 0).all()
    value = per_step * args.steps / dt
    sample = f"first {per_step} of {n} records per step, {cores} host threads; {desc}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": static_config(wdesc, n, profile),
        "details": {"qps_per_step": per_step, "note": "CPU path on the host cores, rank 0 only; qpOASES/Armadillo/Drake are not installable here (DESIGN.md section 5)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def kernel_path():
    """Which balance kernels the default bench command runs (QPB_QPS_PER_WARP / QPB_TPQ_LPQ select experiments)."""
    m = os.environ.get("QPB_QPS_PER_WARP", "")
    if m == "1":
        return "balance_qp_kernel<PackedIO>"
    if m == "2":
        return "balance_qp_kernel16<PackedIO>"
    pdl = "" if os.environ.get("QPB_TPQ_PDL", "1") == "0" else " (loop and finishing passes: programmatic dependent launches)"
    return f"tpq_setup_kernel<PackedIO, false> + tpq_loop_kernel<{os.environ.get('QPB_TPQ_LPQ', '1')}> + tpq_finish_kernel<PackedIO, false>{pdl}"


def fp64_block(iters, n, step_ms):
    """The explanatory roofline of the balance path: algorithmic FP64 work (balance_algorithmic_flops) over the FP64 pipe."""
    flops = float(balance_algorithmic_flops(iters).mean())
    achieved = flops * n / (step_ms * 1e-3) / 1e12
    return {"achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS, "flop_per_qp": flops,
            "model": "450 + 2 x 700 + 600 + 650 x working-set changes (balance_algorithmic_flops); peak = DFMA rate of tools/ubench_fp64"}


def bind_to_local_cpus(local_rank):
    """Pin this rank to the CPUs next to its GPU (sysfs local_cpulist of the device's PCI function), so that the pinned
    buffers it allocates next land on that NUMA node and its copies do not cross the socket interconnect."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        devn = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devn:02x}.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def pcie_bound(pin_in, pin_out, n, dev, barrier, reps=6):
    """Concurrent pinned H2D + D2H of one step's bytes on two streams, no kernels: ms per step on this rank."""
    import torch

    d_i = torch.empty(pin_in[0].nbytes, dtype=torch.uint8, device=dev)
    d_o = torch.empty(pin_out[0].nbytes, dtype=torch.uint8, device=dev)
    h_i = [torch.from_numpy(b.array.view(np.uint8).reshape(-1)) for b in pin_in]
    h_o = [torch.from_numpy(b.array.view(np.uint8).reshape(-1)) for b in pin_out]
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def step(i):
        with torch.cuda.stream(s1):
            d_i.copy_(h_i[i % 2], non_blocking=True)
        with torch.cuda.stream(s2):
            h_o[i % 2].copy_(d_o, non_blocking=True)

    step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(reps):
        step(i)
    s1.synchronize()
    s2.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    return {"ms_per_step": ms}


def secondary_block(solver, params, rank, world, local_rank, dev, dist, barrier):
    """BASELINE configs beside the headline one, each timed with >= 3 launches after the headline region:
    cfg3 (1 048 576 mixed-contact states per GPU) at every N, cfg5 (8 388 608 states over 8 GPUs) when N = 8, and the
    10-step MPC QP of cfg4.  Values are whole-job QP/s (max over ranks of the device time)."""
    import torch

    from quadruped_control_b200 import lib
    from quadruped_control_b200.records import MPC_OUT_DTYPE, default_mpc_params

    out = {}
    stream = torch.cuda.current_stream()
    peak, _ = measured_peak_hbm()

    def timed(fn, launches):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(launches):
            fn()
        e1.record(stream)
        barrier()
        ms, _ = reduce_report(e0.elapsed_time(e1), [0], dist, dev)
        return ms / launches

    jobs = [("cfg3", "cfg3")] + ([("cfg5", "cfg5")] if world == 8 else [])
    for name, wl in jobs:
        n, seed, masks, profile, desc = WORKLOADS[wl]
        S = states.generate_states(n, seed, lo=rank * n, profile=profile, masks=masks)
        d_in = torch.from_numpy(S.view(np.uint8).reshape(-1)).to(dev)
        d_out = torch.empty(n * OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        ms = timed(lambda: solver.control_packed(d_in, d_out, n, stream.cuda_stream), 3)
        res = d_out.cpu().numpy().view(OUT_DTYPE)
        entry = {"workload": desc, "value": world * n / (ms * 1e-3), "unit": UNIT, "ms_per_launch": ms, "launches": 3,
                 "failed_qps": int((res["status"] != 0).sum()), "iters_mean": float(res["iters"].mean()),
                 "hbm_frac": ALGO_BYTES_PER_QP * n / (ms * 1e-3) / 1e9 / peak,
                 "fp64_frac": fp64_block(res["iters"], n, ms)["frac"]}
        if rank == 0:
            import oracle

            stride = max(1, n // 1024)
            ref = oracle.control_batch(params, np.ascontiguousarray(S[::stride]), os.cpu_count() or 1)
            entry["max_rel_grf_err_vs_oracle"] = float((np.abs(res[::stride]["grf_body"] - ref["grf_body"]).max(axis=1) /
                                                        np.maximum(np.abs(ref["grf_body"]).max(axis=1), 1.0)).max())
        out[name] = entry
        del d_in, d_out, S
    # cfg2 as a controller sees it in its loop: every record carries the working set its previous tick ended on (the
    # reference's hotstart), the states have moved by 1 ms of a 1 kHz loop since.  One launch (tpq_one_kernel) at this size.
    n, seed, masks, profile, desc = WORKLOADS["cfg2"]
    S = states.generate_states(n, seed, lo=rank * n, profile=profile, masks=masks)
    d_in = torch.from_numpy(S.view(np.uint8).reshape(-1)).to(dev)
    d_out = torch.empty(n * OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    solver.control_packed(d_in, d_out, n, stream.cuda_stream)
    prev = d_out.cpu().numpy().view(OUT_DTYPE)
    rng = np.random.default_rng(1000 + rank)
    S["pad"][:, :4] = prev["pad"][:, :4]
    S["x"] += rng.normal(0, 2e-4, S["x"].shape)
    S["xdot"] += rng.normal(0, 2e-3, S["xdot"].shape)
    S["w"] += rng.normal(0, 2e-3, S["w"].shape)
    d_in = torch.from_numpy(S.view(np.uint8).reshape(-1)).to(dev)
    solver.set_warm_batches(True)
    try:
        l0 = solver.launches
        ms = timed(lambda: solver.control_packed(d_in, d_out, n, stream.cuda_stream), 5)
        per_call = (solver.launches - l0) // 6
    finally:
        solver.set_warm_batches(False)
    res = d_out.cpu().numpy().view(OUT_DTYPE)
    entry = {"workload": desc + "; the next tick (states moved by 1 ms), records carry the previous tick's working sets",
             "value": world * n / (ms * 1e-3), "unit": UNIT, "ms_per_launch": ms, "launches": 5, "kernels_per_call": per_call,
             "failed_qps": int((res["status"] != 0).sum()), "iters_mean": float(res["iters"].mean()),
             "hbm_frac": ALGO_BYTES_PER_QP * n / (ms * 1e-3) / 1e9 / peak}
    if rank == 0:
        import oracle

        stride = max(1, n // 1024)
        ref = oracle.control_batch(params, np.ascontiguousarray(S[::stride]), os.cpu_count() or 1)
        entry["max_rel_grf_err_vs_oracle"] = float((np.abs(res[::stride]["grf_body"] - ref["grf_body"]).max(axis=1) /
                                                    np.maximum(np.abs(ref["grf_body"]).max(axis=1), 1.0)).max())
    out["cfg2_warm_tick"] = entry
    if world > 1:
        # SURVEY.md 8e, optional: a single consumer wants every shard's results -- one NCCL all-gather of the 200-B result
        # records (forces, torques, status, iteration count: the first 200 bytes of qpb_out_rec) over NVLink / NVSwitch,
        # timed by itself.  Not part of the hot path (no QP waits for another rank) and not in `value`.
        try:
            from quadruped_control_b200.sharding import gather_outputs

            local = d_out.view(n, OUT_DTYPE.itemsize)[:, :200].contiguous()
            whole = torch.empty(world * local.numel(), dtype=torch.uint8, device=dev)
            ms = timed(lambda: gather_outputs(local, dist, whole), 5)
            sums = torch.stack([local.sum(dtype=torch.int64), torch.zeros((), dtype=torch.int64, device=dev)])
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            mine = whole[rank * local.numel():(rank + 1) * local.numel()]
            ok = bool(torch.equal(mine, local.reshape(-1))) and int(whole.sum(dtype=torch.int64)) == int(sums[0])
            recv = (world - 1) * local.numel()
            out["allgather_outputs"] = {
                "what": "torch.distributed.all_gather_into_tensor (NCCL) of 200-B result records, cfg2 shard of every rank -> every rank",
                "bytes_received_per_rank": recv, "ms_per_call": ms, "calls": 5, "gbs_received_per_rank": recv / (ms * 1e-3) / 1e9,
                "nvlink_peak_gbs_per_direction": 900.0, "frac": recv / (ms * 1e-3) / 1e9 / 900.0,
                "peak_source": "nominal NVLink 5 figure (no measured peer bandwidth in MEASURED_PEAKS.json)",
                "gathered_equals_shards": ok}
        except Exception as exc:  # never lose the bench line to the optional leg
            out["allgather_outputs"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    del d_in, d_out, S
    # cfg4: the 10-step convex-MPC QP
    mp = default_mpc_params(MU)
    mpc = lib.MpcSolver(mp, device=local_rank)
    R = states.generate_mpc(MPC_N, MPC_SEED, lo=rank * MPC_N)
    d_in = torch.from_numpy(R.view(np.uint8).reshape(-1)).to(dev)
    d_out = torch.empty(MPC_N * MPC_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: mpc.solve_packed(d_in, d_out, MPC_N, stream.cuda_stream), 3)
    res = d_out.cpu().numpy().view(MPC_OUT_DTYPE)
    flops = float(mpc_algorithmic_flops(R["contact"], res["iters"]).mean())
    out["cfg4_mpc"] = {"workload": MPC_DESC, "value": world * MPC_N / (ms * 1e-3), "unit": UNIT, "ms_per_launch": ms, "launches": 3,
                       "failed_qps": int((res["status"] != 0).sum()), "iters_mean": float(res["iters"].mean()),
                       "fp64_frac": flops * MPC_N / (ms * 1e-3) / 1e12 / 37.0, "parity": "unpinned by construction: no reference MPC code (DESIGN.md section 9)"}
    mpc.close()
    return out


def run_ours(args):
    import torch

    from quadruped_control_b200 import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    n, seed, masks, profile, desc = WORKLOADS[args.workload]
    if args.profile:
        profile = args.profile
        desc += f" ({profile} disturbance profile)"
    params = default_params(MU)
    solver = lib.BalanceSolver(params, device=local_rank)

    # rank r owns records [r*n, (r+1)*n) of the stream; N_ROTATE further disjoint slices rotate through HBM
    host_batches = [states.generate_states(n, seed + 1000 * k, lo=rank * n, profile=profile, masks=masks)
                    for k in range(N_ROTATE if n <= 131072 else 2)]
    n_rot = len(host_batches)
    d_in = [torch.from_numpy(b.view(np.uint8).reshape(-1)).to(dev) for b in host_batches]
    d_out = [torch.empty(n * OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev) for _ in range(n_rot)]
    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------
    for i in range(args.warmup):
        solver.control_packed(d_in[i % n_rot], d_out[i % n_rot], n, stream.cuda_stream)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    launches0 = solver.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    ev0.record(stream)
    for i in range(args.steps):
        solver.control_packed(d_in[i % n_rot], d_out[i % n_rot], n, stream.cuda_stream)
    ev1.record(stream)
    barrier()
    t_end = time.perf_counter()
    launches = solver.launches - launches0
    elapsed_ms_local = ev0.elapsed_time(ev1)
    last = d_out[(args.steps - 1) % n_rot].cpu().numpy().view(OUT_DTYPE)  # results of the last timed step
    # The timed region lasts milliseconds, nvidia-smi samples every 100 ms: keep the very same loop running (untimed) for
    # half a second more so that the clock / throttle samples are taken under this load, and say so in the line.
    clocks = None
    if rank == 0:
        t_more = time.perf_counter()
        while time.perf_counter() - t_more < 0.5:
            for i in range(32):
                solver.control_packed(d_in[i % n_rot], d_out[i % n_rot], n, stream.cuda_stream)
            torch.cuda.synchronize()
        clocks = sampler.stop(t_begin, time.perf_counter())
        clocks["window"] = f"timed region ({(t_end - t_begin) * 1e3:.1f} ms) + 0.5 s untimed continuation of the same loop"
    barrier()

    # correctness of what was just timed (status + checksum vs the oracle on a sample, rank 0)
    failed = int((last["status"] != 0).sum())

    elapsed_ms, (tot_launches, tot_failed) = reduce_report(elapsed_ms_local, [launches, failed], dist, dev)
    total_qps = world * n * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region --
    # Every step uploads its own records from pinned host memory and downloads its results, in the wire format of the
    # host-buffer calls (qpb_wire_state 488 B up, qpb_wire_out 200 B down: the device records without their padding --
    # the call is bound by the PCIe link, so bytes are speed).  Steps are queued with the asynchronous entry point, two
    # batches in flight, so the upload of step k+1 overlaps the download of step k (PCIe is full duplex); the final
    # qpb_host_sync() is inside the timed region.  The synchronous call and the padded 512 / 256-B records are timed beside it.
    bind_to_local_cpus(local_rank)
    e2e_steps = max(3, min(args.steps, 20))

    def time_host_calls(dt_in, dt_out, fill, call_sync, call_async):
        pin_i = [lib.PinnedBuffer(n, dt_in) for _ in range(2)]
        pin_o = [lib.PinnedBuffer(n, dt_out) for _ in range(2)]
        for k in range(2):
            fill(pin_i[k].array, host_batches[k % n_rot])
        for i in range(2):
            call_sync(pin_i[i % 2].array, pin_o[i % 2].array)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            call_sync(pin_i[i % 2].array, pin_o[i % 2].array)
        sync_local = (time.perf_counter() - t0) * 1e3
        barrier()
        for i in range(2):  # untimed: the asynchronous path creates its streams and staging buffers on first use
            call_async(pin_i[i % 2].array, pin_o[i % 2].array)
        solver.host_sync()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            call_async(pin_i[i % 2].array, pin_o[i % 2].array)
        solver.host_sync()
        async_local = (time.perf_counter() - t0) * 1e3
        barrier()
        async_ms, _ = reduce_report(async_local, [0], dist, dev)
        sync_ms, _ = reduce_report(sync_local, [0], dist, dev)
        ok = bool((pin_o[0].array["status"] == 0).all() and (pin_o[1].array["status"] == 0).all())
        return pin_i, pin_o, world * n * e2e_steps / (async_ms * 1e-3), world * n * e2e_steps / (sync_ms * 1e-3), ok

    def fill_padded(dst, src):
        dst[:] = src

    pad_in, pad_out, padded_qps, padded_sync_qps, padded_ok = time_host_calls(
        STATE_DTYPE, OUT_DTYPE, fill_padded, solver.control_host, solver.control_host_async)
    for b in pad_in + pad_out:
        b.free()
    pin_in, pin_out, e2e_qps, e2e_sync_qps, e2e_checksum_ok = time_host_calls(
        WIRE_STATE_DTYPE, WIRE_OUT_DTYPE, lambda dst, src: to_wire(src, dst), solver.control_wire_host, solver.control_wire_host_async)
    e2e_checksum_ok = e2e_checksum_ok and padded_ok
    # the same bytes with no solver in between: what the PCIe link (and the host memory behind it) gives this rank while
    # every rank does the same -- the bound of the end-to-end number
    pcie = pcie_bound(pin_in, pin_out, n, dev, barrier)
    pcie_ms, _ = reduce_report(pcie["ms_per_step"], [0], dist, dev)
    pcie_qps = world * n / (pcie_ms * 1e-3)

    secondary = secondary_block(solver, params, rank, world, local_rank, dev, dist, barrier) if (args.secondary and args.workload == "cfg2") else None

    if rank == 0:
        # the CPU-baseline leg also checks what was just timed against the oracle (strided sample of the last step)
        stride = max(1, n // 2048)
        cpu, err = cpu_baseline(params, host_batches[0], check=(np.ascontiguousarray(host_batches[(args.steps - 1) % n_rot][::stride]),
                                                                np.ascontiguousarray(last[::stride])))
        peak, peak_src = measured_peak_hbm()
        kernel_ms = elapsed_ms / args.steps  # one kernel launch per step on the timed stream
        achieved = ALGO_BYTES_PER_QP * n / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": total_qps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": static_config(desc, n, profile),
            "details": {"global_batch": world * n, "parallelism": f"batch-sharded x{world}",
                        "l2": f"{n_rot} distinct device batches rotate ({n_rot * n * 768 / 1e6:.0f} MB > 126 MB L2)" if n_rot * n * 768 > 126e6 else "inputs larger than L2",
                        "iters_mean": float(last["iters"].mean()), "iters_max": int(last["iters"].max()),
                        "kernel_path": kernel_path(),
                        "active_rows_hist": active_rows_histogram(host_batches[(args.steps - 1) % n_rot], last, params.mu, params.fzmin, params.fzmax)},
            "max_rel_grf_err_vs_oracle": err, "failed_qps": int(tot_failed),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch(args.workload), "peak_source": peak_src,
                         "ncu": ncu_explanatory(args.workload), "static": "traffic and ncu are copied from profiles/ncu_summary.json (one ncu capture of this command), not measured by this run",
                         "algorithmic_bytes_per_qp": ALGO_BYTES_PER_QP, "kernel": kernel_path(),
                         "kernel_ms": kernel_ms,
                         "fp64": fp64_block(last["iters"], n, kernel_ms),
                         "note": "the path is FP64-issue/latency bound, not DRAM bound (DESIGN.md); per-GPU figure"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": n * WIRE_STATE_DTYPE.itemsize,
                    "d2h_bytes_per_step": n * WIRE_OUT_DTYPE.itemsize, "steps": e2e_steps, "ok": e2e_checksum_ok,
                    "api": "qpb_control_batch_wire_host_async x steps + qpb_host_sync (pinned host buffers of 488-B / 200-B wire records, two batches in flight, staged H2D / kernels / D2H)",
                    "sync_call_value": e2e_sync_qps, "sync_api": "qpb_control_batch_wire_host (returns when the batch is complete)",
                    "padded_records_value": padded_qps, "padded_records_sync_value": padded_sync_qps,
                    "padded_records_api": "qpb_control_batch_host_async / qpb_control_batch_host on the 512-B / 256-B device records (round 1's and the C++ shim's format)",
                    "pcie_bound_qps": pcie_qps, "pcie_bound_gbs": pcie_qps * (WIRE_STATE_DTYPE.itemsize + WIRE_OUT_DTYPE.itemsize) / 1e9,
                    "pcie_frac": e2e_qps / pcie_qps,
                    "pcie_note": "bound = the same pinned buffers copied H2D and D2H concurrently (cudaMemcpyAsync on two streams) on every rank at once, no kernels"},
            "secondary": secondary,
            "gpu_launches": int(tot_launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    for b in pin_in + pin_out:
        b.free()
    solver.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def active_rows_histogram(S, out, mu, fzmin, fzmax):
    """Active pyramid rows at the returned solution, per QP (SURVEY.md 8d asks for the histogram beside every
    throughput number): world-frame forces f_w = -R f_b, rows |fx| <= mu fz, |fy| <= mu fz, fzmin <= fz <= fzmax."""
    R = S["Rwb"].reshape(-1, 3, 3)
    fb = out["grf_body"].reshape(-1, 4, 3)
    fw = -np.einsum("nij,nlj->nli", R, fb)
    stance = S["contact"].astype(bool)
    fx, fy, fz = fw[..., 0], fw[..., 1], fw[..., 2]
    tol = 1e-7
    act = ((mu * fz - np.abs(fx) <= tol * (1 + mu * np.abs(fz))).astype(int) + (mu * fz - np.abs(fy) <= tol * (1 + mu * np.abs(fz))).astype(int)
           + (fz - fzmin <= tol * (1 + abs(fzmin))).astype(int) + (fzmax - fz <= tol * (1 + abs(fzmax))).astype(int))
    per_qp = (act * stance).sum(axis=1)[out["status"] == 0]
    hist = np.bincount(per_qp, minlength=1)
    return [int(v) for v in hist]


def mpc_algorithmic_flops(contact, iters):
    """FP64 flops the MPC path needs per QP, as a model: Cholesky + triangular inverse of the n x n Hessian (2 n^3 / 3),
    its assembly (about 40 per lower-triangle entry) and two triangular mat-vecs plus the N* updates per iteration."""
    n = 3.0 * contact.reshape(len(contact), -1).sum(axis=1)
    q = np.minimum(iters, n)
    return 2.0 * n**3 / 3.0 + 20.0 * n * n + iters * (2.0 * n * n + 8.0 * 0.5 * q * n)


def run_mpc_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle
    from quadruped_control_b200.records import default_mpc_params

    p = default_mpc_params(MU)
    cores = os.cpu_count() or 1
    R_full = states.generate_mpc(4096, MPC_SEED)
    t0 = time.perf_counter()
    oracle.mpc_batch(p, R_full[:256], cores)
    rate = 256 / (time.perf_counter() - t0)
    total_steps = args.steps + args.warmup
    per_step = int(min(len(R_full), max(64, rate * 100.0 / total_steps)))
    R = np.ascontiguousarray(R_full[:per_step])
    for _ in range(args.warmup):
        oracle.mpc_batch(p, R, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = oracle.mpc_batch(p, R, cores)
    dt = time.perf_counter() - t0
    assert (out["status"] == 0).all()
    value = per_step * args.steps / dt
    sample = f"first {per_step} of {MPC_N} records per step, {cores} host threads; oracle port (the reference has no MPC code)"
    print(json.dumps({
        "impl": "reference", "metric": MPC_METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": MPC_DESC, "qps_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)


def run_mpc(args):
    """--workload cfg4: the batched 10-step convex-MPC QP through qpb_mpc_batch_packed / qpb_mpc_batch_host."""
    import torch

    from quadruped_control_b200 import lib
    from quadruped_control_b200.records import MPC_ALGO_BYTES_PER_QP, MPC_OUT_DTYPE, MPC_REC_DTYPE, default_mpc_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    n = MPC_N
    p = default_mpc_params(MU)
    solver = lib.MpcSolver(p, device=local_rank)
    host = [states.generate_mpc(n, MPC_SEED + 1000 * k, lo=rank * n) for k in range(2)]  # 143 MB each: larger than L2
    d_in = [torch.from_numpy(b.view(np.uint8).reshape(-1)).to(dev) for b in host]
    d_out = [torch.empty(n * MPC_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev) for _ in range(2)]
    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        solver.solve_packed(d_in[i % 2], d_out[i % 2], n, stream.cuda_stream)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    launches0 = solver.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    ev0.record(stream)
    for i in range(args.steps):
        solver.solve_packed(d_in[i % 2], d_out[i % 2], n, stream.cuda_stream)
    ev1.record(stream)
    barrier()
    t_end = time.perf_counter()
    launches = solver.launches - launches0
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    last = d_out[(args.steps - 1) % 2].cpu().numpy().view(MPC_OUT_DTYPE)
    failed = int((last["status"] != 0).sum())
    elapsed_ms, (tot_launches, tot_failed) = reduce_report(ev0.elapsed_time(ev1), [launches, failed], dist, dev)
    total_qps = world * n * args.steps / (elapsed_ms * 1e-3)

    pin_in, pin_out = lib.PinnedBuffer(n, MPC_REC_DTYPE), lib.PinnedBuffer(n, MPC_OUT_DTYPE)
    pin_in.array[:] = host[0]
    e2e_steps = max(3, min(args.steps, 10))
    solver.solve_host(pin_in.array, pin_out.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        solver.solve_host(pin_in.array, pin_out.array)
    e2e_local = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms, _ = reduce_report(e2e_local, [0], dist, dev)
    e2e_qps = world * n * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        import oracle

        cores = os.cpu_count() or 1
        hb = host[(args.steps - 1) % 2]
        idx = np.arange(0, n, n // 256)
        ref = oracle.mpc_batch(p, np.ascontiguousarray(hb[idx]), cores)
        err = float((np.abs(last["U"][idx] - ref["U"]).max(axis=1) / np.maximum(np.abs(ref["U"]).max(axis=1), 1.0)).max())
        t0 = time.perf_counter()
        oracle.mpc_batch(p, hb[:512], cores)
        rate = 512 / (time.perf_counter() - t0)
        m = int(min(n, max(512, rate * 10.0)))
        t0 = time.perf_counter()
        oracle.mpc_batch(p, hb[:m], cores)
        cpu_rate = m / (time.perf_counter() - t0)
        peak, peak_src = measured_peak_hbm()
        kernel_ms = elapsed_ms / args.steps
        achieved = MPC_ALGO_BYTES_PER_QP * n / (kernel_ms * 1e-3) / 1e9
        flops = float(mpc_algorithmic_flops(hb["contact"], last["iters"].astype(np.float64)).mean())
        tflops = flops * n / (kernel_ms * 1e-3) / 1e12
        print(json.dumps({
            "metric": MPC_METRIC, "value": total_qps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": MPC_DESC, "qps_per_step_per_gpu": n, "global_batch": world * n, "parallelism": f"batch-sharded x{world}",
                       "l2": "inputs larger than L2 (143 MB of records per step, two batches alternate)",
                       "iters_mean": float(last["iters"].mean()), "iters_max": int(last["iters"].max())},
            "max_rel_force_err_vs_oracle": err, "failed_qps": int(tot_failed),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_per_launch("cfg4"),
                         "peak_source": peak_src, "algorithmic_bytes_per_qp": MPC_ALGO_BYTES_PER_QP, "kernel": "mpc_qp_kernel", "kernel_ms": kernel_ms,
                         "fp64": {"achieved_tflops": tflops, "peak_tflops": FP64_PEAK_TFLOPS, "frac": tflops / FP64_PEAK_TFLOPS,
                                  "algorithmic_flops_per_qp": flops, "peak_source": "nominal 148 SM x 64 lanes x 2 x 1.965 GHz (tools/ubench_fp64: 64.0 lanes/clk/SM measured)"},
                         "note": "compute path: ~2e6 FP64 flops and 3.1 KB per QP; the FP64 fraction is the meaningful one (SURVEY.md 8d)"},
            "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"first {m} records, {cores} threads; oracle/mpc_oracle.c (dense condensed model + Goldfarb-Idnani); the reference has no MPC code"},
            "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": n * MPC_REC_DTYPE.itemsize, "d2h_bytes_per_step": n * MPC_OUT_DTYPE.itemsize,
                    "steps": e2e_steps, "ok": bool((pin_out.array["status"] == 0).all()), "api": "qpb_mpc_batch_host (pinned host buffers, staged H2D/kernel/D2H ring)"},
            "gpu_launches": int(tot_launches), "clocks": clocks}), flush=True)
    pin_in.free()
    pin_out.free()
    solver.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_tick(args):
    """--workload tick: the whole controller tick per robot on the device (SURVEY.md 8f ranks 1, 3, 4 around the hot path):
    CoMState/JointState messages -> adapter -> foothold planner + swing-foot reference -> balance QP + J^T f + swing-leg PD
    -> JointTorqueCmd, five launches on one stream, 1 048 576 mixed-contact robots per GPU."""
    import torch

    from quadruped_control_b200 import lib
    from quadruped_control_b200.records import (COM_MSG_DTYPE, JOINT_MSG_DTYPE, PLAN_DTYPE, SWING_DTYPE, TORQUE_CMD_DTYPE,
                                                default_joint_gains, default_plan_params)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    n = 1048576
    params = default_params(MU)
    solver = lib.BalanceSolver(params, device=local_rank)
    S = states.generate_states(n, 20260103, lo=rank * n, profile="light", masks="mixed")
    rng = np.random.default_rng(7 + rank)
    com = np.zeros(n, dtype=COM_MSG_DTYPE)
    from scipy.spatial.transform import Rotation

    com["orientation"] = Rotation.from_matrix(S["Rwb"].reshape(n, 3, 3)).as_quat()
    com["position"], com["linear"], com["angular"] = S["x"], S["xdot"], S["w"]
    js = np.zeros(n, dtype=JOINT_MSG_DTYPE)
    js["position"] = S["q"].reshape(n, 4, 3).transpose(0, 2, 1).reshape(n, 12)
    js["velocity"] = rng.normal(0, 1.0, size=(n, 12))
    plan = np.zeros(n, dtype=PLAN_DTYPE)
    plan["phase"] = rng.uniform(0.83, 1.0, size=(n, 4))
    plan["replan"] = 1

    def pin(a):
        return torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory()

    h_com, h_js, h_plan = pin(com), pin(js), pin(plan)
    d_com, d_js, d_S = h_com.to(dev), h_js.to(dev), torch.from_numpy(S.view(np.uint8).reshape(-1)).to(dev)
    d_plan0 = h_plan.to(dev)
    d_plan = d_plan0.clone()
    d_sw = torch.zeros(n * SWING_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(n * OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_cmd = torch.zeros(n * TORQUE_CMD_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    h_cmd = torch.empty(n * TORQUE_CMD_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    h_plan_out = torch.empty(n * PLAN_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream

    def tick():
        solver.adapt_inputs(d_com, d_js, d_S, d_sw, n, stream=st)
        solver.plan(d_S, d_plan, d_sw, n, stream=st)
        solver.tick_packed(d_S, d_sw, d_out, n, stream=st)
        solver.torque_cmd(d_S, d_out, d_cmd, n, stream=st)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        tick()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    launches0 = solver.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        tick()
    ev1.record(stream)
    barrier()
    t_end = time.perf_counter()
    launches = solver.launches - launches0
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    out = d_out.cpu().numpy().view(OUT_DTYPE)
    failed = int((out["status"] != 0).sum())
    elapsed_ms, (tot_launches, tot_failed) = reduce_report(ev0.elapsed_time(ev1), [launches, failed], dist, dev)
    total = world * n * args.steps / (elapsed_ms * 1e-3)

    # end to end: messages and plan records up from pinned host memory, torque commands and plan records back
    e2e_steps = max(3, min(args.steps, 10))

    def tick_e2e():
        d_com.copy_(h_com, non_blocking=True)
        d_js.copy_(h_js, non_blocking=True)
        d_plan.copy_(h_plan, non_blocking=True)
        tick()
        h_cmd.copy_(d_cmd, non_blocking=True)
        h_plan_out.copy_(d_plan, non_blocking=True)
        torch.cuda.synchronize()

    tick_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tick_e2e()
    e2e_local = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms, _ = reduce_report(e2e_local, [0], dist, dev)
    e2e = world * n * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        import oracle

        cores = os.cpu_count() or 1
        m = 65536
        S_ref, plan_ref = S[:m].copy(), plan[:m].copy()
        sw_ref = np.zeros(m, dtype=SWING_DTYPE)
        gains, pp = default_joint_gains(), default_plan_params()
        t0 = time.perf_counter()
        oracle.adapt_inputs(params, com[:4096], js[:4096], S_ref[:4096], sw_ref[:4096])
        t_adapt = (time.perf_counter() - t0) / 4096  # python loop over a C call: an upper bound, reported separately
        t0 = time.perf_counter()
        oracle.plan_batch(pp, S_ref, plan_ref, sw_ref)
        t_plan = (time.perf_counter() - t0) / m
        t0 = time.perf_counter()
        ref = oracle.tick_batch(params, gains, S_ref, sw_ref, cores)
        t_tick = (time.perf_counter() - t0) / m
        cpu_rate = 1.0 / (t_plan / 1.0 + t_tick)  # planner single-threaded + tick on all cores, adapter excluded
        ok = np.isfinite(ref["tau"]).all(axis=1) & (np.abs(ref["tau"]).max(axis=1) < 1e3)
        err = float((np.abs(out["grf_body"][:m] - ref["grf_body"]).max(axis=1) / np.maximum(np.abs(ref["grf_body"]).max(axis=1), 1.0))[ok].max())
        bytes_per_robot = 104 + 192 + 240 + 240 + 512 + 288 + 256 + 112  # messages, plan in/out, state, swing, result, command
        peak, peak_src = measured_peak_hbm()
        kernel_ms = elapsed_ms / args.steps
        achieved = bytes_per_robot * n / (kernel_ms * 1e-3) / 1e9
        print(json.dumps({
            "metric": "control ticks/sec (messages -> torque commands, batched)", "value": total, "unit": "ticks/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "tick: 1048576 mixed 2/3/4-foot robots per GPU, light profile; adapter + planner (footholds planned on the first tick, swing-foot references every tick; the end-to-end leg re-plans every tick) + balance QP + swing PD + torque command",
                       "robots_per_step_per_gpu": n, "parallelism": f"batch-sharded x{world}", "l2": "inputs larger than L2 (1.9 GB of records per step)"},
            "max_rel_grf_err_vs_oracle": err, "failed_qps": int(tot_failed),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_robot": bytes_per_robot, "kernel": "adapt + plan + balance_qp_kernel16 + swing + torque_cmd",
                         "kernel_ms": kernel_ms, "note": "five launches; the balance kernel is about three quarters of the step and is FP64-issue bound"},
            "cpu_baseline": {"value": cpu_rate, "unit": "ticks/s", "cores": cores, "kind": "port",
                             "sample": f"first {m} robots: oracle planner (1 thread) + oracle tick ({cores} threads); message adapter not included ({t_adapt * 1e6:.1f} us per robot through a python loop)"},
            "e2e": {"value": e2e, "unit": "ticks/s", "h2d_bytes_per_step": n * (104 + 192 + 240), "d2h_bytes_per_step": n * (112 + 240), "steps": e2e_steps,
                    "api": "qpb_adapt_inputs_batch + qpb_plan_batch + qpb_tick_batch_packed + qpb_torque_cmd_batch around pinned-memory copies"},
            "gpu_launches": int(tot_launches), "clocks": clocks}), flush=True)
    solver.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false",
                    help="skip the cfg3 / cfg5 / cfg4 block timed after the headline region (default: on for the cfg2 line)")
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["cfg4", "tick"], default="cfg2")
    ap.add_argument("--profile", choices=("default", "light", "stress"), default=None,
                    help="disturbance profile of the synthetic states (SURVEY.md 8d); default: the workload's own")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE line, the JSON: while the run is going on, file descriptor 1 points at stderr, so that
    # whatever a library prints there (NCCL's version banner at communicator creation) cannot get in front of it
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.workload == "tick":
        if args.impl == "reference":
            raise SystemExit("--workload tick has no reference arm (use the default workload)")
        run_tick(args)
    elif args.workload == "cfg4":
        (run_mpc_reference if args.impl == "reference" else run_mpc)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
