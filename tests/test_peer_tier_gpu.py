"""NVLink tier (include/hpsx.h hpsx_cache_peer_tier_*): the host table sharded over the GPUs' HBM, cache misses read
from the owner's shard by the unchanged direct-pull kernels.  One GPU is enough to test the logic: two parameter
servers stand for two ranks of one box (their shards then live on the same device, the "peer" pointers are plain
device pointers).  The servers hold DIFFERENT values for the same keys, so every returned row proves which shard
served it: owner(key) == 0 -> rank 0's values, owner(key) == 1 -> rank 1's.  Multi-GPU (real NVLink, CUDA IPC):
tests/test_sharded_gpu.py::test_peer_tier_*.

Reference behaviour the tier must keep: docs/hierarchical_parameter_server.md:244-246 (a key in no database gets the
default vector), hps_backend/src/model_state.cpp:124-178 (database reload while serving).
"""
import numpy as np
import pytest

import hugectr_backend_b200 as hb
from oracle import hps_oracle as O

pytestmark = pytest.mark.gpu

SEEDS = (0xB2000010, 0xB2000011)


def _torch():
    import torch

    return torch


def make_rank(rows, dim, seed, *, cache_pct=0.1, name="m", default=0.25, extra_rows=0):
    hps = hb.HPS(num_partitions=8)
    hps.add_model(hb.ModelParams(name, 8192, [dim], [1], [default], hit_rate_threshold=1.0, cache_size_percentage=cache_pct,
                                 enable_pagelock=True))
    hps.load_table_procedural(name, 0, rows + extra_rows, seed)
    hps.create_embedding_cache(name)
    ref = O.NumpyTable(dim, default)
    ref.fill_procedural(rows + extra_rows, seed)
    return hps, ref


def owners(keys, world):
    out = np.empty(len(keys), dtype=np.uint32)
    L = hb.lib()
    k = np.ascontiguousarray(keys, dtype=np.int64)
    assert L.hpsx_owner_batch(k.ctypes.data, len(k), world, out.ctypes.data) == 0
    return out


@pytest.mark.parametrize("dim", [128, 32, 20])
def test_two_ranks_on_one_device_serve_each_others_misses(cuda_device, dim):
    torch = _torch()
    rows, n = 60000, 8192
    WARM = 6000  # cache_pct 0.1
    ranks = [make_rank(rows, dim, SEEDS[r]) for r in range(2)]
    for r, (hps, _) in enumerate(ranks):
        hps.peer_tier_build("m", 0, r, 2)
    for r, (hps, _) in enumerate(ranks):
        p = 1 - r
        hps.peer_tier_attach_local("m", 0, p, ranks[p][0], "m", 0)
        hps.peer_tier_commit("m", 0)
        info = hps.peer_tier_info("m", 0)
        assert info["world"] == 2 and info["rank"] == r and info["committed"] == 1
        assert info["index_entries_in_tier"] == rows  # every key of the table is served from a shard
        assert abs(info["own_rows"] - rows / 2) < 8 * np.sqrt(rows / 2) + 8
    refs = [ref for _, ref in ranks]
    rng = np.random.default_rng(5)
    for r, (hps, _) in enumerate(ranks):
        s = hps.session("m", 0)
        for it in range(3):
            # keys past the warm-up range (those were cached from the rank's OWN table at cache creation);
            # some keys are in no database -> default vector
            keys = rng.integers(WARM, rows + 500, size=n)
            keys[::97] = -7
            out = torch.full((n, dim), float("nan"), device="cuda")
            s.lookup([keys], [out], [n])
            own = owners(keys, 2)
            exp = refs[r].lookup(keys)
            other = (own == 1 - r) & (keys < rows) & (keys >= 0)
            exp[other] = refs[1 - r].lookup(keys[other])
            got = out.cpu().numpy()
            assert np.array_equal(got, exp), f"rank {r} request {it}"
        st = s.stats()
        assert st.misses > 0 and st.tier_bytes > 0 and st.default_filled > 0
        # nothing crossed the host link except the keys themselves
        assert st.h2d_bytes == 3 * n * 8
        del s
    for hps, _ in ranks:
        hps.peer_tier_detach("m", 0)
    # after the detach the index holds host addresses again: every rank answers with its own values
    for r, (hps, ref) in enumerate(ranks):
        s = hps.session("m", 0)
        keys = rng.integers(WARM, rows, size=n)
        out = torch.full((n, dim), float("nan"), device="cuda")
        s.lookup([keys], [out], [n])
        # rows cached while the tier was attached keep the values they were cached with (a detach does not
        # invalidate the cache), so compare the keys this rank owns: same values either way
        st = s.stats()
        assert st.tier_bytes == 0
        got = out.cpu().numpy()
        own = owners(keys, 2)
        mine = own == r
        assert np.array_equal(got[mine], ref.lookup(keys[mine]))
        del s


def test_tier_pull_with_heavy_key_duplication(cuda_device):
    """Quads of the tier pull whose entries share buckets (the same missing key many times in one request) are claimed
    one by one; nothing may hang and every occurrence gets its row."""
    torch = _torch()
    rows, dim, n = 60000, 128, 8192
    hps, ref = make_rank(rows, dim, SEEDS[0], cache_pct=0.05)
    hps.peer_tier_build("m", 0, 0, 1)
    hps.peer_tier_commit("m", 0)
    s = hps.session("m", 0)
    rng = np.random.default_rng(29)
    for distinct in (1, 3, 40, 500):
        pool = rng.integers(rows // 2, rows, size=distinct)
        keys = pool[rng.integers(0, distinct, size=n)]
        out = torch.full((n, dim), float("nan"), device="cuda")
        s.lookup([keys], [out], [n])
        assert np.array_equal(out.cpu().numpy(), ref.lookup(keys))
    # cache pressure: far more distinct missing keys than slots, several rounds (evictions under the quad claims)
    for _ in range(6):
        keys = rng.integers(rows // 20, rows, size=n)
        out = torch.full((n, dim), float("nan"), device="cuda")
        s.lookup([keys], [out], [n])
        assert np.array_equal(out.cpu().numpy(), ref.lookup(keys))
    resident = hps.cache_keys("m", 0, 0)
    assert len(np.unique(resident)) == len(resident)  # no key was placed twice


def test_world_one_tier_is_the_whole_table_in_hbm(cuda_device):
    torch = _torch()
    rows, dim, n = 30000, 128, 4096
    hps, ref = make_rank(rows, dim, SEEDS[0], cache_pct=0.05)
    hps.peer_tier_build("m", 0, 0, 1)
    hps.peer_tier_commit("m", 0)
    info = hps.peer_tier_info("m", 0)
    assert info["own_rows"] == rows and info["index_entries_in_tier"] == rows
    s = hps.session("m", 0)
    rng = np.random.default_rng(11)
    for _ in range(4):
        keys = rng.integers(0, rows, size=n)
        out = torch.full((n, dim), float("nan"), device="cuda")
        s.lookup([keys], [out], [n])
        assert np.array_equal(out.cpu().numpy(), ref.lookup(keys))
    st = s.stats()
    assert st.h2d_bytes == 4 * n * 8 and st.tier_bytes == (st.misses - st.default_filled) * dim * 4


def test_commit_needs_every_shard_and_build_needs_pagelock(cuda_device):
    rows, dim = 5000, 32
    hps, _ = make_rank(rows, dim, SEEDS[0])
    hps.peer_tier_build("m", 0, 0, 2)
    with pytest.raises(Exception, match="not attached"):
        hps.peer_tier_commit("m", 0)
    hps.peer_tier_detach("m", 0)
    plain = hb.HPS(num_partitions=4)
    plain.add_model(hb.ModelParams("p", 256, [dim], [1], [0.0], cache_size_percentage=0.5, enable_pagelock=False))
    plain.load_table_procedural("p", 0, rows, 1)
    plain.create_embedding_cache("p")
    with pytest.raises(Exception, match="enable_pagelock"):
        plain.peer_tier_build("p", 0, 0, 2)


def test_batched_and_large_requests_through_the_tier(cuda_device):
    """The merged miss list of a batch of requests and a chunked large request use the same index."""
    torch = _torch()
    rows, dim = 400000, 128
    ranks = [make_rank(rows, dim, SEEDS[r], cache_pct=0.1) for r in range(2)]
    for r, (hps, _) in enumerate(ranks):
        hps.peer_tier_build("m", 0, r, 2)
    for r, (hps, _) in enumerate(ranks):
        hps.peer_tier_attach_local("m", 0, 1 - r, ranks[1 - r][0], "m", 0)
        hps.peer_tier_commit("m", 0)
    refs = [ref for _, ref in ranks]
    hps = ranks[0][0]
    s = hps.session("m", 0)
    rng = np.random.default_rng(3)

    def exp_for(keys):
        own = owners(keys, 2)
        e = refs[0].lookup(keys)
        e[own == 1] = refs[1].lookup(keys[own == 1])
        return e

    reqs = []
    for r in range(6):
        n = int(rng.integers(100, 1500))
        keys = rng.integers(rows // 10, rows, size=n)  # past the warm-up range
        out = torch.full((n, dim), float("nan"), device="cuda")
        reqs.append(([keys], [out], [n]))
    s.lookup_batch(reqs)
    for keys_l, out_l, _ in reqs:
        assert np.array_equal(out_l[0].cpu().numpy(), exp_for(keys_l[0]))
    del s, ranks


# ------------------------------------------------------------------------------------------------
# tables that live in the tier only ("synthetic_device:" — model-parallel rows without a host copy, BASELINE configs[3])
# ------------------------------------------------------------------------------------------------
def make_device_table_rank(rows, dim, seed, *, default=0.25, max_batch=8192):
    hps = hb.HPS(num_partitions=8)
    hps.add_model(hb.ModelParams("mp", max_batch, [dim], [1], [default], hit_rate_threshold=1.0, cache_size_percentage=0.0,
                                 enable_pagelock=True, embedding_cache_type="static", peer_tier=True,
                                 sparse_files=[f"synthetic_device:rows={rows},seed={seed}"]))
    return hps


@pytest.mark.parametrize("dim", [128, 32, 20])
def test_device_table_single_rank(cuda_device, dim):
    """One device: create_embedding_cache builds the world-1 tier by itself (the table has no other copy)."""
    torch = _torch()
    rows, n, seed = 150_000, 8192, 0xB2000020
    hps = make_device_table_rank(rows, dim, seed)
    hps.create_embedding_cache("mp")
    assert hps.table_rows("mp", 0) == rows
    info = hps.peer_tier_info("mp", 0)
    assert info["world"] == 1 and info["committed"] == 1 and info["own_rows"] == rows and info["index_entries_in_tier"] == rows
    ref = O.NumpyTable(dim, 0.25)
    ref.fill_procedural(rows, seed)
    s = hps.session("mp", 0)
    rng = np.random.default_rng(17)
    for n_req in (n, 1, 4097, 3):
        keys = rng.integers(-3, rows + 300, size=n_req)
        out = torch.full((n_req, dim), float("nan"), device="cuda")
        s.lookup([keys], [out], [n_req])
        assert np.array_equal(out.cpu().numpy(), ref.lookup(keys))
        d_keys = torch.from_numpy(keys).cuda()
        out2 = torch.full((n_req, dim), float("nan"), device="cuda")
        s.lookup_device_keys([d_keys], [out2], [n_req])
        assert torch.equal(out, out2)
    st = s.stats()
    assert st.hits == 0 and st.default_filled > 0 and st.tier_bytes == (st.misses - st.default_filled) * dim * 4
    # host output (Triton may hand back CPU memory)
    keys = rng.integers(0, rows, size=1000)
    h_out = np.full((1000, dim), np.nan, dtype=np.float32)
    s.lookup_ex([keys], [h_out], [1000], key_memory="host", vector_memory="host")
    assert np.array_equal(h_out, ref.lookup(keys))


def test_device_table_two_ranks_on_one_device(cuda_device):
    """Two ranks (two servers on one device): every rank generates only the rows it owns, maps the other's shard and
    builds its index from both; each serves the whole key range."""
    torch = _torch()
    rows, dim, n, seed = 300_000, 128, 8192, 0xB2000021
    ranks = [make_device_table_rank(rows, dim, seed) for _ in range(2)]
    for r, hps in enumerate(ranks):
        # deployed on one device the cache creation builds a world-1 tier; rebuild it as rank r of 2
        hps.create_embedding_cache("mp")
        hps.peer_tier_build("mp", 0, r, 2)
    own = [hps.peer_tier_info("mp", 0)["own_rows"] for hps in ranks]
    assert sum(own) == rows and all(abs(o - rows / 2) < 8 * np.sqrt(rows / 2) + 8 for o in own)
    for r, hps in enumerate(ranks):
        hps.peer_tier_attach_local("mp", 0, 1 - r, ranks[1 - r], "mp", 0)
        hps.peer_tier_commit("mp", 0)
        assert hps.peer_tier_info("mp", 0)["index_entries_in_tier"] == rows
    ref = O.NumpyTable(dim, 0.25)
    ref.fill_procedural(rows, seed)
    rng = np.random.default_rng(23)
    for hps in ranks:
        s = hps.session("mp", 0)
        keys = rng.integers(-3, rows + 300, size=n)
        out = torch.full((n, dim), float("nan"), device="cuda")
        s.lookup([keys], [out], [n])
        assert np.array_equal(out.cpu().numpy(), ref.lookup(keys))
        del s
    for hps in ranks:
        hps.peer_tier_detach("mp", 0)
