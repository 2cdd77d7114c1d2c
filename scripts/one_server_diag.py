"""Where does the time of the one-server Triton arm go?  One backend, the DCN model on two GPUs with the tier, one
instance per GPU: each instance alone, then both at once (C++ threads in the harness)."""
import sys, os, json, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import fake_triton as FT
import bench

dim, slots, batch, rows = 128, 26, 65536, 10_000_000
n = batch * slots
m = bench._ps_model("dcn", rows, bench.SEED, dim, slots, batch, 0, gpucacheper=0.2)
m["deployed_device_list"] = [0, 1]
m["hpsx_peer_tier"] = os.environ.get("DIAG_TIER", "1") == "1"
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "ps.json")
    json.dump({"supportlonglong": True, "volatile_db": {"type": "parallel_hash_map", "num_partitions": 16}, "models": [m]}, open(path, "w"))
    with FT.Backend(path) as be:
        model = be.model("dcn", FT.model_config("dcn", gpus=[0, 1], max_batch_size=batch))
        insts = [model.instance(name=f"dcn_{d}", kind=FT.KIND_GPU, device=d) for d in (0, 1)]
        numkeys = np.array([[n]], dtype=np.int32)
        rng = np.random.default_rng(5)
        prepared, keep = [], []
        for d, inst in enumerate(insts):
            out = torch.empty(n * dim, device=f"cuda:{d}")
            reqs = []
            for _ in range(12):
                k = rng.integers(2_000_000, rows, size=n, dtype=np.int64)
                hot = rng.random(n) < 0.87
                k[hot] = rng.integers(0, 2_000_000, size=int(hot.sum()))
                reqs.append(torch.from_numpy(k).pin_memory())
            keep.append((out, reqs))
            prepared.append(inst.prepare([dict(keys=t.numpy(), numkeys=numkeys, gpu_out=out, out_device=d) for t in reqs]))
        FT.run_sequences_parallel(prepared, 0, 12)
        for d in (0, 1):
            torch.cuda.synchronize(d)
        for d in (0, 1):
            t = FT.run_sequences_parallel([prepared[d]], 2, 12)
            st = insts[d].stats()
            print(f"instance {d} alone: {t / 10 * 1e3:.3f} ms/step; last compute window {(st['compute_end'] - st['compute_start']) / 1e6:.3f} ms, "
                  f"exec window {(st['exec_end'] - st['exec_start']) / 1e6:.3f} ms", flush=True)
        t = FT.run_sequences_parallel(prepared, 2, 12)
        for d in (0, 1):
            st = insts[d].stats()
            print(f"both: {t / 10 * 1e3:.3f} ms/step; instance {d} last compute window {(st['compute_end'] - st['compute_start']) / 1e6:.3f} ms", flush=True)
        for p in prepared: p.close()
        for i in insts: i.close()
        model.close()
