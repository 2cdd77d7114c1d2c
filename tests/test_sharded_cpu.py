"""Model-parallel row sharding at world size 2 over gloo, on the CPU parameter-server path:
owner routing, the count/key/row exchanges and the inverse permutation.  The GPU path swaps in the routing and
scatter kernels and NCCL; the sequencing code (hugectr_backend_b200/sharded.py) is the same."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rows, dim, seed, ret, max_batch=1 << 16, skewed=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import hugectr_backend_b200 as hb
    from hugectr_backend_b200.sharded import ShardedLookup
    from oracle import hps_oracle as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        hps = hb.HPS(num_partitions=4, num_threads=2)
        hps.add_model(hb.ModelParams("dlrm", max_batch, [dim], [1], [0.25], use_gpu_embedding_cache=False))
        hps.load_table_procedural_shard("dlrm", 0, rows, seed, rank, world)
        owned = hps.table_rows("dlrm", 0)
        ref = O.NumpyTable(dim, 0.25)
        ref.fill_procedural(rows, seed)
        sl = ShardedLookup(hps, "dlrm", 0, dim, device=-1)
        ok = True
        rng = np.random.default_rng(100 + rank)
        for n in ((max_batch, max_batch) if skewed else (1, 257, 5000, 0)):
            keys = rng.integers(-5, rows + 5, size=n)  # a few keys no shard owns -> default value
            if skewed:  # every rank asks only for keys rank 0 owns: rank 0 receives world x max_batch keys, twice what one
                pool = np.arange(rows)  # request of its session may hold — it must serve them in pieces, not fail mid-exchange
                pool = pool[O.owner(pool, world) == 0]
                keys = pool[rng.integers(0, len(pool), size=n)]
            out = sl.lookup(torch.from_numpy(keys))
            ok &= np.array_equal(out.numpy(), ref.lookup(keys))
            own = O.owner(keys, world)
            ok &= np.array_equal(sl.last["send_counts"], np.bincount(own, minlength=world)) if n else True
        total = torch.tensor([owned])
        dist.all_reduce(total)
        ret[rank] = (bool(ok), int(owned), int(total[0]))
    finally:
        dist.destroy_process_group()


def test_sharded_lookup_world2_gloo():
    import torch.multiprocessing as mp

    rows, dim, seed, world = 20000, 16, 9, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, dim, seed, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret[0][0] and ret[1][0]
    # shards partition the table: balanced by the hash, nothing lost or duplicated
    assert ret[0][2] == rows and abs(ret[0][1] - rows / 2) < 0.05 * rows


def test_skewed_ownership_is_served_in_pieces():
    import torch.multiprocessing as mp

    rows, dim, seed, world = 20000, 16, 9, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, dim, seed, ret, 512, True)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret[0][0] and ret[1][0]


def test_owner_batch_matches_scalar_and_oracle():
    import hugectr_backend_b200 as hb
    from hugectr_backend_b200 import hps as H
    from oracle import hps_oracle as O

    keys = np.random.default_rng(3).integers(-(1 << 62), 1 << 62, size=5000)
    for shards in (1, 2, 3, 8, 64):
        b = H.owner_batch(keys, shards)
        assert np.array_equal(b, O.owner(keys, shards)) and b.max() < shards
        assert all(H.owner(int(k), shards) == int(o) for k, o in zip(keys[:50], b[:50]))
