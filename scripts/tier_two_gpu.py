"""One process, two GPUs: kernel time (CUDA events, session stats) and wall time of tier lookups, one GPU at a time and
both at once — separates the gather kernels from the host path of the one-server bench arms."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hugectr_backend_b200 as hb

dim, slots, batch = 128, 26, 65536
n = batch * slots

def timed(sessions, reqs, outs, reps, label):
    for s in sessions:
        s.reset_stats()
    def work(i):
        torch.cuda.set_device(i)
        for r in range(reps):
            sessions[i].lookup_device_keys([reqs[i][r % len(reqs[i])]], [outs[i]], [n])
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(sessions))]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    wall = (time.perf_counter() - t0) / reps * 1e3
    for i, s in enumerate(sessions):
        st = s.stats()
        print(f"  {label} GPU{i}: wall {wall:.3f} ms/step, probe {st.probe_kernel_ms / max(1, st.probe_kernel_launches):.3f} ms, "
              f"pull/gather {st.pull_kernel_ms / reps:.3f} ms, tier {st.tier_bytes / reps / 1e6:.1f} MB/step, misses {st.misses // reps}", flush=True)

def run(name, params_kw, load, rows):
    hps = hb.HPS(num_partitions=16)
    hps.add_model(hb.ModelParams(name, batch, [dim], [slots], [0.0], hit_rate_threshold=1.0, deployed_devices=[0, 1],
                                 enable_pagelock=True, peer_tier=True, **params_kw))
    load(hps)
    hps.create_embedding_cache(name)
    print(name, hps.peer_tier_info(name, 0), flush=True)
    sess = [hps.session(name, d) for d in (0, 1)]
    rng = np.random.default_rng(3)
    reqs, outs = [], []
    for d in (0, 1):
        with torch.cuda.device(d):
            reqs.append([torch.from_numpy(rng.integers(0, rows, size=n, dtype=np.int64)).cuda() for _ in range(6)])
            outs.append(torch.empty((n, dim), device="cuda"))
    for d in (0, 1):
        torch.cuda.synchronize(d)
    timed(sess, reqs, outs, 2, "warm-up")
    timed(sess[:1], reqs[:1], outs[:1], 6, "GPU0 alone")
    timed(sess, reqs, outs, 6, "both")
    del sess, hps

rows_mp = 2 * 20_000_000
run("mp", dict(cache_size_percentage=0.0, embedding_cache_type="static",
               sparse_files=[f"synthetic_device:rows={rows_mp},seed=7"]), lambda h: None, rows_mp)
rows = 10_000_000
run("dcn", dict(cache_size_percentage=0.2), lambda h: h.load_table_procedural("dcn", 0, rows, 11), rows)
