"""world_size-2 checks of the multi-process host logic on CPU (gloo): shard split, checksum all-gather, max-over-ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from detectinblur_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_indices(17, rank, world)
    sums = sharding.gather_checksums(0xF000000000000001 + rank)
    slow = sharding.max_over_ranks(1.0 + rank)
    out.put((rank, mine, sums, slow))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, i0, s0, m0), (r1, i1, s1, m1) = res
    # DistributedSampler split: strided, padded to equal length by wrapping
    assert i0 == [0, 2, 4, 6, 8, 10, 12, 14, 16] and i1 == [1, 3, 5, 7, 9, 11, 13, 15, 0]
    assert s0 == s1 == [0xF000000000000001, 0xF000000000000002]
    assert m0 == m1 == 2.0


def test_shard_indices_single_process():
    assert sharding.shard_indices(8, 0, 1) == list(range(8))
    assert sharding.shard_indices(8, 3, 4) == [3, 7]
    assert sharding.shard_indices(10, 1, 4, drop_last=True) == [1, 5]
    assert sharding.gather_checksums(5) == [5]
    got = [sorted(sharding.shard_indices(64, r, 8)) for r in range(8)]
    assert sorted(sum(got, [])) == list(range(64))
