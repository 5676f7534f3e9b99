"""CPU, world_size 2 over gloo: the N>1 host logic — stream sharding and the one-time broadcast
of the table blob rank 0 designs (the only collective of the path)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    import importlib
    pkg = load_package()
    sh = importlib.import_module("sdrjfm_b200.sharding")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = pkg.design_tables(input_filter_hz=165000, audio_lp_hz=15000).blob
    blob = sh.broadcast_tables(mine if rank == 0 else None, dist)
    T = pkg.Tables(blob)
    ok = bool(np.array_equal(blob, mine)) and int(T.hdr["fm_rate"]) == 192000 and len(T.fmband1) == 25
    lo, hi = sh.stream_range(256 + 3, world, rank)
    q.put((rank, ok, lo, hi, int(blob.size)))
    dist.destroy_process_group()


def test_table_broadcast_and_sharding_world2(pkg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=60) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 130, 130, 259)
    assert res[0][4] == res[1][4] > 1_000_000


def test_stream_range_covers_everything(pkg):
    import importlib
    sh = importlib.import_module("sdrjfm_b200.sharding")
    for total in (1, 7, 256, 259):
        for world in (1, 2, 4, 8):
            spans = [sh.stream_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
