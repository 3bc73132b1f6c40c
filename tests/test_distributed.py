"""N > 1 host logic on CPU: sweeps shard across ranks with no data-path collective; the only
collectives are the benchmark's barrier / MAX-of-time / SUM-of-units (gloo, world_size 2)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from pp_b200.pipeline import shard_range
    lo, hi = shard_range(n_items, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi))
    ms, units = bench.reduce_over_ranks(ms_local=10.0 * (rank + 1), units_local=float(hi - lo),
                                        device=torch.device("cpu"))
    q.put((rank, gathered, ms, units))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_reduce_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_items = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, gathered, ms, units in res:
        assert gathered == [(0, 4), (4, 7)]              # contiguous, disjoint, covering
        assert ms == 20.0                                 # MAX over ranks
        assert units == float(n_items)                    # SUM over ranks


def test_shard_range_properties():
    from pp_b200.pipeline import shard_range
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= (n + w - 1) // w
