#!/usr/bin/env python
"""Times one probe+gather kernel variant on fixed requests against a STATIC cache (no inserts, state never changes).
usage: HPSX_PIPE_CFG=4x4 python scripts/probe_variants.py pipe [hit_fraction]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hugectr_backend_b200 as hb

variant = sys.argv[1] if len(sys.argv) > 1 else "ldg"
hit = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows, dim, n, seed = 4_000_000, 128, 65536 * 26, 0xB2000002
os.environ["HPSX_DIRECT_PULL"] = "1"
hps = hb.HPS(num_partitions=16)
hps.add_model(hb.ModelParams("m", 65536, [dim], [26], [0.0], cache_size_percentage=0.5, embedding_cache_type="dynamic",
                             hit_rate_threshold=1.0, cache_load_factor=0.5))
hps.load_table_procedural("m", 0, rows, seed)
hps.create_embedding_cache("m")
hot = hps.cache_keys("m", 0, 0)
rng = np.random.default_rng(1)
reqs = []
for _ in range(6):
    k = hot[rng.integers(0, len(hot), size=n)]
    if hit < 1.0:
        cold = rng.random(n) >= hit
        k[cold] = rng.integers(rows, 2 * rows, size=int(cold.sum()))  # keys in no table: misses that are never inserted
    reqs.append(torch.from_numpy(k).cuda())
out = torch.empty((n, dim), device="cuda")
s = hps.session("m", 0)
s.set_probe_variant(variant)
for i in range(3):
    s.lookup_device_keys([reqs[i]], [out], [n])
s.reset_stats()
for i in range(12):
    s.lookup_device_keys([reqs[i % 6]], [out], [n])
st = s.stats()
ms = st.probe_kernel_ms / st.probe_kernel_launches
alg = (st.hits * (8 + 8 * dim) + st.misses * (8 + 4 * dim)) / st.lookups
print(f"{variant:5s} cfg={os.environ.get('HPSX_PIPE_CFG', '-'):5s} hit={st.hits / st.keys:.3f} kernel {ms * 1e3:7.1f} us  {alg / ms / 1e6:7.1f} GB/s "
      f"({alg / ms / 1e6 / 6537.6:.3f} of measured peak)")
