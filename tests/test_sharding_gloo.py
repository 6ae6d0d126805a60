"""world_size-2 gloo test of the N>1 path's host logic (no GPU)."""
import importlib
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sh = importlib.import_module("go-curdleproofs_b200.sharding")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sh.shard_bounds(total, world, rank)
    local = [1 if (i % 8) else 0 for i in range(lo, hi)]  # every 8th proof is a mutated (rejected) one
    full = sh.gather_verdicts(local, total, world, rank)
    q.put((rank, lo, hi, full))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [4096, 13, 1])
def test_shard_and_gather_world2(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [1 if (i % 8) else 0 for i in range(total)]
    covered = []
    for rank, lo, hi, full in sorted(res):
        assert full == want
        covered.extend(range(lo, hi))
    assert covered == list(range(total))


def test_shard_bounds_properties(pkg):
    sh = importlib.import_module("go-curdleproofs_b200.sharding")
    for total in (0, 1, 7, 4096):
        for world in (1, 2, 4, 8):
            spans = [sh.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
