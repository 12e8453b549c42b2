"""N > 1 path on CPU: two gloo ranks shard a batch by buffer (no data-path collective), index their shards with
the host walker, and reduce the bench counters the way bench.py does (max of times, sum of work)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import draco_sharp_b200 as D
    from draco_sharp_b200 import synth_gen as G
    # every rank generates and owns its own buffers (distinct seeds), as bench.py does per GPU
    spec = G.make_spec(2000, seed=bench.rank_seed(rank), scheme=1, colors=1)
    arena, offs, lens, sums, schemes, used = G.synth_batch(spec, 8, n_threads=2)
    bufs = [arena[int(o): int(o + l)] for o, l in zip(offs, lens)]
    bt = D.index_only(bufs)
    pts = bt.points
    first = bytes(bufs[0][:64])
    ms, (tot_pts, tot_out) = bench.reduce_over_ranks(dist, "cpu", 10.0 + rank, [pts, bt.out_bytes])
    # static sharding of ONE shared list of buffers by rank::world covers it exactly once
    owned = list(range(rank, 21, world))
    t = torch.zeros(21, dtype=torch.int64)
    t[owned] = 1
    dist.all_reduce(t)
    out_q.put((rank, pts, ms, tot_pts, tot_out, first, bool((t == 1).all())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (r0, p0, ms0, tp0, to0, f0, c0), (r1, p1, ms1, tp1, to1, f1, c1) = res
    assert p0 == p1 == 8 * 2000
    assert ms0 == ms1 == 11.0                  # max over ranks
    assert tp0 == tp1 == 2 * 8 * 2000          # whole-job work: sum over ranks
    assert to0 == to1 and to0 > 0
    assert f0 != f1                            # distinct data per rank
    assert c0 and c1                           # rank::world sharding is a partition
