"""Repro harness for the tier-only (synthetic_device) table at scale: which of {rows, keys per request} breaks it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hugectr_backend_b200 as hb
import bench

def run(rows, n, dim=128, seed=0xB2000000 + 44, host_keys=True):
    hps = hb.HPS(num_partitions=8)
    batch = max(1, (n + 25) // 26)
    hps.add_model(hb.ModelParams("mp", batch, [dim], [26], [0.0], hit_rate_threshold=1.0, cache_size_percentage=0.0,
                                 enable_pagelock=True, embedding_cache_type="static", peer_tier=True,
                                 sparse_files=[f"synthetic_device:rows={rows},seed={seed}"]))
    hps.create_embedding_cache("mp")
    info = hps.peer_tier_info("mp", 0)
    s = hps.session("mp", 0)
    rng = np.random.default_rng(1)
    keys = rng.integers(0, rows, size=n, dtype=np.int64)
    out = torch.full((n, dim), float("nan"), device="cuda")
    if host_keys:
        s.lookup([keys], [out], [n])
    else:
        s.lookup_device_keys([torch.from_numpy(keys).cuda()], [out], [n])
    torch.cuda.synchronize()
    d_keys = torch.from_numpy(keys).cuda()
    j = torch.arange(dim, device="cuda", dtype=torch.int64)
    bad = 0
    for b0 in range(0, n, 1 << 18):
        k = d_keys[b0:b0 + (1 << 18)]
        r = bench._splitmix64_torch(torch, k[:, None] * 131 + j[None, :] + seed)
        bits = ((r >> 41) & ((1 << 23) - 1)) | 0x3F800000
        expect = bits.to(torch.int32).view(torch.float32) - 1.5
        bad += int((out[b0:b0 + (1 << 18)] != expect).any(dim=1).sum())
    st = s.stats()
    print(f"rows={rows} n={n} host_keys={host_keys}: wrong rows {bad}, own_rows {info['own_rows']}, in index {info['index_entries_in_tier']}, "
          f"default_filled {st.default_filled}", flush=True)
    if bad:
        wrong = (out != out).any(dim=1).sum().item()
        zero = (out == 0).all(dim=1).sum().item()
        print(f"   NaN rows {wrong}, all-zero rows {zero}", flush=True)
    del s, hps

for rows, n in [(150_000, 8192), (150_000, 1_703_936), (50_000_000, 8192), (50_000_000, 1_703_936), (125_000_000, 1_703_936)]:
    run(rows, n)
run(50_000_000, 1_703_936, host_keys=False)
