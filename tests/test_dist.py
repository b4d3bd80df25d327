"""world_size-2 gloo test (CPU) of the host-side multi-GPU logic: shard ownership, the max/sum report
reduction that bench.py uses, and the optional all-gather of the result records."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quadruped_control_b200 import states
from quadruped_control_b200.sharding import gather_outputs, reduce_report, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    S = states.generate_states(hi - lo, 20260105, lo=lo, masks="mixed")
    checksum = int(np.frombuffer(S.tobytes(), dtype=np.uint64).sum(dtype=np.uint64) & np.uint64(0xFFFFFFFF))
    elapsed, (count, csum) = reduce_report(10.0 + rank, [hi - lo, checksum], dist)
    # equal shards of result-sized records (200-B wire records), gathered in rank order
    m = n_total // world
    mine = torch.from_numpy(np.frombuffer(states.generate_states(m, 7, lo=rank * m).tobytes(), dtype=np.uint8)[:m * 200].copy())
    whole = gather_outputs(mine, dist)
    if rank == 0:
        q.put((elapsed, count, csum, whole.numpy().tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_report():
    n_total, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    elapsed, count, csum, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert elapsed == 11.0  # max over ranks
    assert count == n_total
    full = states.generate_states(n_total, 20260105, masks="mixed")
    lo, hi = shard_range(n_total, 0, 2)
    parts = [full[lo:hi], full[hi:]]
    want = sum(int(np.frombuffer(p.tobytes(), dtype=np.uint64).sum(dtype=np.uint64) & np.uint64(0xFFFFFFFF)) for p in parts)
    assert csum == want
    m = n_total // world
    assert gathered == b"".join(states.generate_states(m, 7, lo=r * m).tobytes()[:m * 200] for r in range(world))


def test_single_process_report_is_identity():
    assert reduce_report(3.5, [1, 2], None) == (3.5, [1.0, 2.0])
    t = torch.arange(5)
    assert gather_outputs(t, None) is t
