"""CPU tests of the benchmark's host-side helpers (no GPU, no timing): the active-row histogram that every bench line
carries, the ncu figures quoted beside the roofline, and the shape of the workload table."""
import json
import os

import numpy as np

import bench
import oracle
from quadruped_control_b200 import default_params, states

NCPU = os.cpu_count() or 1


def test_active_rows_histogram_matches_the_oracle_working_sets(built):
    p = default_params(0.6)
    S = states.generate_states(1024, 20260103, masks="mixed")
    out = oracle.control_batch(p, S, NCPU)
    hist = bench.active_rows_histogram(S, out, p.mu, p.fzmin, p.fzmax)
    assert sum(hist) == int((out["status"] == 0).sum()) and len(hist) <= 13  # 12 variables: at most 12 independent rows
    # a stance-pose robot at rest needs no row at all; two-foot stances can never have more than 6 active rows
    S0 = states.stance_state(default_params(0.8))
    assert bench.active_rows_histogram(S0, oracle.control_batch(default_params(0.8), S0), 0.8, 10.0, 120.0) == [1]
    two = S[S["contact"].astype(bool).sum(axis=1) == 2]
    h2 = bench.active_rows_histogram(two, oracle.control_batch(p, two, NCPU), p.mu, p.fzmin, p.fzmax)
    assert len(h2) <= 7
    # light profile: the bimodal histogram SURVEY App. E describes (a large share of unconstrained solutions)
    L = states.generate_states(1024, 20260102, profile="light")
    hl = bench.active_rows_histogram(L, oracle.control_batch(p, L, NCPU), p.mu, p.fzmin, p.fzmax)
    assert hl[0] > 0.2 * sum(hl)


def test_ncu_summary_feeds_the_roofline_fields():
    with open(os.path.join(bench.ROOT, "profiles", "ncu_summary.json")) as f:
        summary = json.load(f)
    for wl in ("cfg2", "cfg3", "cfg4"):
        assert bench.ncu_traffic_per_launch(wl) == summary[wl]["dram_bytes_per_launch"] > 0
        ex = bench.ncu_explanatory(wl)
        assert 0 < ex["fp64_pipe_pct"] < 100 and 0 < ex["issue_slots_busy_pct"] < 100
        assert os.path.exists(os.path.join(bench.ROOT, ex["source"]))
    assert bench.ncu_traffic_per_launch("no-such-workload") is None and bench.ncu_explanatory("no-such-workload") is None
    # the three launches of a config-2 step read the 512-byte records once from DRAM and the scratch they hand each other
    # (which mostly stays in the 126-MB L2 at this size): between 1x and 3.5x the input, never more
    assert 1.0 <= summary["cfg2"]["dram_read_bytes"] / (65536 * 512) < 3.5
    for src in summary["cfg3"]["sources"]:
        assert os.path.exists(os.path.join(bench.ROOT, src))


def test_workload_table_is_the_baseline_configs():
    assert bench.WORKLOADS["cfg2"][:3] == (65536, 20260102, "all4")
    assert bench.WORKLOADS["cfg3"][:3] == (1048576, 20260103, "mixed")
    assert bench.WORKLOADS["cfg5"][:3] == (1048576, 20260105, "mixed")  # per GPU; 8 388 608 over 8
    assert bench.MPC_N == 65536 and bench.ALGO_BYTES_PER_QP == 680
    peak, src = bench.measured_peak_hbm()
    assert 3000 < peak < 9000 and ("measured" in src or "fallback" in src)


def test_reference_arm_contract_under_torchrun(built):
    """`bench.py --impl reference` launched as the driver launches it for N > 1: rank 0 alone runs the CPU path and prints
    exactly one JSON line with the contract's keys; the other rank exits 0 without work.  Needs no GPU."""
    import socket
    import subprocess
    import sys

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(bench.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=bench.ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "QP/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 2 and d["steps"] == 1 and d["warmup"] == 3 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("cfg2")
