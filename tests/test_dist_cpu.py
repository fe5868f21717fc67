"""Host-side multi-process logic on CPU: world_size 2, gloo backend (no GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chore_b200 import dist as cd


def test_shard_range_covers_everything_without_overlap():
    for total in (0, 1, 127, 128, 129, 20000, 256 ** 3, 1000003):
        for w in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(w):
                s, c = cd.shard_range(total, r, w)
                assert s == pos and c >= 0
                assert s % 128 == 0 or c == 0 or s == total
                pos += c
            assert pos == total
    assert cd.shard_images(32, 3, 8) == [12, 13, 14, 15]
    assert sum(len(cd.shard_images(5, r, 4)) for r in range(4)) == 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, count = cd.shard_range(total)
        # stand-in for "evaluate my slab": the value of point i is f(i)
        idx = torch.arange(start, start + count, dtype=torch.float32)
        local = torch.stack([idx * 2.0, idx + 0.5])                     # (n_out=2, count)
        full = cd.gather_results(local, dim=1)
        ref = torch.arange(total, dtype=torch.float32)
        ok = torch.equal(full, torch.stack([ref * 2.0, ref + 0.5]))
        # per-image summaries: rank r owns images shard_images(5)
        mine = cd.shard_images(5)
        summ = torch.tensor([[float(i), float(i) ** 2] for i in mine]).reshape(-1, 2)
        allsum = cd.gather_results(summ, dim=0)
        ok = ok and torch.equal(allsum, torch.tensor([[float(i), float(i) ** 2] for i in range(5)]))
        out_q.put((rank, bool(ok), cd.world()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_results_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1000, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res
    assert all(r[2][1] == 2 for r in res)
