"""CPU, world_size=2, gloo: the N>1 host logic -- slice assignment and the optional final gather
of (n, fp, ier).  The fit itself has no collective (independent curves)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fitpack_b200.dist import gather_fit_summary, shard_range


def test_shard_range_covers_batch_exactly():
    for nbatch in (0, 1, 127, 128, 129, 1000, 1 << 20, (1 << 20) + 77):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            for r in range(world):
                b0, b1 = shard_range(nbatch, r, world)
                assert b0 == min(seen, nbatch) and b1 >= b0
                assert b0 % 128 == 0 or b0 == nbatch
                seen = b1
            assert seen == nbatch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nbatch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the "fit" of this rank's slice is faked with a deterministic function of the curve id:
        # this test is about who owns which curve and how the summary is reassembled
        b0, b1 = shard_range(nbatch, rank, world)
        ids = torch.arange(b0, b1)
        n = (8 + ids % 50).to(torch.int32)
        ier = torch.where(ids % 7 == 0, -2, 0).to(torch.int32)
        fp = (ids.to(torch.float64) * 0.125 + 2.56)
        n_all, fp_all, ier_all = gather_fit_summary(n, fp, ier, nbatch=nbatch)
        ids = torch.arange(nbatch)
        ok = bool(torch.equal(n_all, (8 + ids % 50).to(torch.int32))
                  and torch.equal(ier_all, torch.where(ids % 7 == 0, -2, 0).to(torch.int32))
                  and torch.equal(fp_all, ids.to(torch.float64) * 0.125 + 2.56))
        q.put((rank, ok, b0, b1))
    finally:
        dist.destroy_process_group()


def _run(world, nbatch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nbatch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_two_rank_gather_over_gloo():
    res = _run(2, 1000)
    assert all(ok for _, ok, _, _ in res)
    assert res[0][2:] == (0, 512) and res[1][2:] == (512, 1000)


def test_two_rank_gather_with_an_empty_slice():
    res = _run(2, 100)   # 100 curves: rank 0 takes them all (128-aligned), rank 1 is empty
    assert all(ok for _, ok, _, _ in res)
    assert res[1][2] == res[1][3] == 100
