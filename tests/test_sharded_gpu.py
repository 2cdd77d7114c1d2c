"""Model-parallel rows over 2+ GPUs of one box (config C4 shape, scaled down): NCCL all-to-all key routing
(mode "nccl") and the fused exchange over NVLink peer memory through CUDA IPC (mode "p2p", hpsx_shard_group_*).
Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_sharded_gpu.py -m gpu`."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rows, dim, seed, pagelock, mode, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import hugectr_backend_b200 as hb
    from hugectr_backend_b200.sharded import ShardedLookup
    from oracle import hps_oracle as O

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        hps = hb.HPS(num_partitions=8, num_threads=4)
        hps.add_model(hb.ModelParams("dlrm", 1 << 17, [dim], [1], [0.25], cache_size_percentage=0.6,
                                     hit_rate_threshold=1.0, deployed_devices=[rank], enable_pagelock=pagelock))
        hps.load_table_procedural_shard("dlrm", 0, rows, seed, rank, world)
        hps.create_embedding_cache("dlrm")
        ref = O.NumpyTable(dim, 0.25)
        ref.fill_procedural(rows, seed)
        sl = ShardedLookup(hps, "dlrm", 0, dim, device=rank, mode=mode)
        ok = True
        rng = np.random.default_rng(100 + rank)
        for n in (1, 4097, 60000, 0, 60000):
            keys = rng.integers(-5, rows + 5, size=n)
            out = sl.lookup(torch.from_numpy(keys).cuda())
            ok &= bool(np.array_equal(out.cpu().numpy(), ref.lookup(keys)))
            if n:
                ok &= bool(np.array_equal(sl.last["send_counts"], np.bincount(O.owner(keys, world), minlength=world)))
        sl.close()
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
@pytest.mark.parametrize("pagelock", [False, True])
def test_sharded_lookup_nccl(cuda_device, pagelock, mode):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    rows, dim, seed = 200_000, 128, 21
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, dim, seed, pagelock, mode, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert all(ret[r] for r in range(world))


def _tier_worker(rank, world, port, rows, dim, seed, ret):
    """One replica per GPU (the reference's multi-GPU mode) + the NVLink tier between them over CUDA IPC."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import hugectr_backend_b200 as hb
    from oracle import hps_oracle as O

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        hps = hb.HPS(num_partitions=8, num_threads=4)
        hps.add_model(hb.ModelParams("dcn", 1 << 16, [dim], [1], [0.25], cache_size_percentage=0.1,
                                     hit_rate_threshold=1.0, deployed_devices=[rank], enable_pagelock=True))
        # every replica holds the whole table, but with its OWN values: a returned row names the rank that served it
        hps.load_table_procedural("dcn", 0, rows, seed + rank)
        hps.create_embedding_cache("dcn")
        refs = []
        for r in range(world):
            t = O.NumpyTable(dim, 0.25)
            t.fill_procedural(rows, seed + r)
            refs.append(t)

        def gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        info = hps.peer_tier_connect_distributed("dcn", rank, rank, world, 1, gather)
        ok = info["committed"] == 1 and info["index_entries_in_tier"] == rows
        s = hps.session("dcn", rank)
        rng = np.random.default_rng(200 + rank)
        warm = rows // 10
        for n in (1, 4097, 60000, 60000):
            keys = rng.integers(warm, rows + 5, size=n)
            out = torch.full((n, dim), float("nan"), device="cuda")
            s.lookup([keys], [out], [n])
            own = O.owner(keys, world)
            exp = refs[rank].lookup(keys)  # default vector for keys past the table
            for r in range(world):
                sel = (own == r) & (keys < rows)
                exp[sel] = refs[r].lookup(keys[sel])
            ok &= bool(np.array_equal(out.cpu().numpy(), exp))
        st = s.stats()
        ok &= st.tier_bytes > 0 and st.h2d_bytes == (1 + 4097 + 60000 + 60000) * 8
        del s
        dist.barrier()  # nobody reads a shard any more
        hps.peer_tier_detach("dcn", rank)
        dist.barrier()  # every rank has unmapped its peers before any shard is freed with its cache
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_peer_tier_over_cuda_ipc(cuda_device):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    rows, dim, seed = 300_000, 128, 77
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_tier_worker, args=(r, world, port, rows, dim, seed, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert all(ret[r] for r in range(world))


def test_one_server_all_gpus_with_the_tier(tmp_path, cuda_device):
    """The deployment `bench.py --gpus N` measures: ONE backend process, the model on every GPU of the box with
    "hpsx_peer_tier": true, one instance per GPU.  The table is loaded from reference-format sparse files; every GPU's
    responses are bit-exact, and the misses were served from the tier (nothing but keys crossed PCIe is not observable
    through Triton, so the engine-level twin of this test — tests/test_peer_tier_gpu.py — checks the byte counters)."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fake_triton as FT
    from oracle import hps_oracle as O
    from test_triton_backend_cpu import model_entry, ps_json, write_tables

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    devs = list(range(world))
    dirs, tables = write_tables(str(tmp_path), [(60000, 32)])
    entry = model_entry("dcn", dirs, [32], [26], gpucache=True, max_batch=512, hit_rate_threshold=1.0, gpucacheper=0.1,
                        enable_pagelock=True, hpsx_peer_tier=True, devices=devs, workers=1)
    ps = ps_json(str(tmp_path / "ps.json"), [entry])
    ref = O.NumpyTable(32, 0.0)
    ref.insert(*tables[0])
    rng = np.random.default_rng(9)
    with FT.Backend(ps) as be:
        m = be.model("dcn", FT.model_config("dcn", gpus=devs, max_batch_size=512))
        insts = [m.instance(name=f"dcn_{d}", kind=FT.KIND_GPU, device=d) for d in devs]
        for it in range(3):
            for d, inst in enumerate(insts):
                n = 512 * 26
                keys = rng.choice(tables[0][0], size=n)
                keys[::101] = -5  # in no database: default vector
                out = torch.full((n * 32,), float("nan"), device=f"cuda:{d}")
                r = inst.infer(keys, np.array([[n]], dtype=np.int32), gpu_out=out, out_device=d)
                assert r.error_code is None, r.error_message
                assert r.params["DeviceID"] == d
                assert np.array_equal(out.cpu().numpy(), ref.lookup(keys).ravel()), (it, d)
        for inst in insts:
            inst.close()
        m.close()
