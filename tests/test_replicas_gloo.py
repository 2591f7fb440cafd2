"""The N>1 host logic (replica sharding + max-time / sum-work aggregation) on CPU with gloo, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from faster_qwen3_tts.replicas import aggregate, rtf, shard_requests


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_requests(7, rank, world)
    dist.barrier()
    t, w = aggregate([100.0 + 50.0 * rank, 10.0], [128.0 * len(mine), 1.0])
    q.put((rank, mine, t, w))
    dist.destroy_process_group()


def test_two_replicas_aggregate_like_bench():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, t0, w0), (r1, m1, t1, w1) = out
    assert sorted(m0 + m1) == list(range(7)) and not set(m0) & set(m1)
    assert t0 == t1 == [150.0, 10.0]          # max over ranks
    assert w0 == w1 == [128.0 * 7, 2.0]       # sum over ranks
    assert abs(rtf(w0[0], t0[0]) - 128 * 7 * 0.08 / 0.15) < 1e-9


def test_single_process_is_identity():
    assert aggregate([3.0], [4.0]) == ([3.0], [4.0])
    assert shard_requests(5, 0, 1) == [0, 1, 2, 3, 4]
