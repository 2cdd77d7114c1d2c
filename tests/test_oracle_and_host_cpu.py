"""CPU-side tests: the oracle against the committed golden vectors (tests/golden, generator committed
beside them), the two oracle implementations against each other, the engine's host parameter server
(the `gpucache: false` path of libhpsx.so) against the oracle, ps.json parsing, and the C-ABI surface.
Nothing here needs a GPU; nothing here calls a GPU entry point.
"""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import hugectr_backend_b200 as hb
from hugectr_backend_b200 import _native as N
from oracle import hps_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(GOLDEN, "hps_kat.npz"))


def tables_from(kat, impl):
    t0 = impl(1, float(kat["wdl_defaults"][0]))
    t0.insert(kat["wdl_k0"], kat["wdl_v0"])
    t1 = impl(16, float(kat["wdl_defaults"][1]))
    t1.insert(kat["wdl_k1"], kat["wdl_v1"])
    return [t0, t1]


# ------------------------------------------------------------------------------------------------
# oracle vs golden vectors
# ------------------------------------------------------------------------------------------------
def test_synthetic_rows_match_golden_bits(kat):
    seed = int(kat["synth_seed"])
    rows = O.synth_rows(kat["synth_keys"], 8, seed)
    assert np.array_equal(rows.view(np.uint32), kat["synth_bits_dim8"])
    assert rows.min() >= -0.5 and rows.max() < 0.5
    L = O.c_lib()
    for i, k in enumerate(kat["synth_keys"]):
        for j in range(8):
            c = np.float32(L.hps_oracle_synth_value(int(k), j, seed))
            assert c.view(np.uint32) == kat["synth_bits_dim8"][i, j]


def test_owner_hash_matches_golden(kat):
    assert np.array_equal(O.owner(kat["owner_keys"], 8), kat["owner_8"])
    assert np.array_equal(O.owner(kat["owner_keys"], 3), kat["owner_3"])
    for k, o8, o3 in zip(kat["owner_keys"], kat["owner_8"], kat["owner_3"]):
        assert O.c_owner(int(k), 8) == o8 and O.c_owner(int(k), 3) == o3
        assert hb.hps.owner(int(k), 8) == o8 and hb.hps.owner(int(k), 3) == o3  # the engine's host-side router


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_wdl_sample_request_matches_golden(kat, impl):
    """10 samples, tables d=[1,16] -> OUTPUT0 [4180] (Hierarchical_Parameter_Server_Deployment.ipynb:738-753,793-795)."""
    if impl == "numpy":
        out = O.request(tables_from(kat, O.NumpyTable), kat["wdl_KEYS"], kat["wdl_NUMKEYS"].ravel())
    else:
        out = O.c_request(tables_from(kat, O.CTable), kat["wdl_KEYS"].ravel(), kat["wdl_NUMKEYS"].ravel(), threads=3)
    assert out.shape == (4180,) and np.array_equal(out, kat["wdl_OUTPUT0"])


@pytest.mark.parametrize("impl", [O.NumpyTable, O.CTable])
def test_ensemble_sample_duplicates_defaults_and_pooling_match_golden(kat, impl):
    t = impl(16, 1.0)
    t.insert(kat["ens_keys"], kat["ens_vecs"])
    keys = kat["ens_KEYS"].ravel()
    assert np.array_equal(t.lookup(keys).ravel(), kat["ens_OUTPUT0"])
    if impl is O.NumpyTable:
        s, m = O.pooled(t, keys, 64, 3, "sum"), O.pooled(t, keys, 64, 3, "mean")
    else:
        s, m = t.pooled(keys, 64, 3, "sum"), t.pooled(keys, 64, 3, "mean")
    assert np.array_equal(s, kat["ens_pooled_sum"]) and np.array_equal(m, kat["ens_pooled_mean"])


def test_unique_matches_golden(kat):
    keys = kat["ens_KEYS"].ravel()
    for u, inv in (O.unique_first_occurrence(keys), O.c_unique(keys)):
        assert np.array_equal(u, kat["ens_unique"]) and np.array_equal(inv, kat["ens_inverse"])
        assert np.array_equal(u[inv], keys)


def test_sparse_dir_written_by_the_reference_samples_writer(kat):
    """tests/golden/sparse_d4 was produced by the sample's struct.pack writer (01_model_training.ipynb:498-505)."""
    d = os.path.join(GOLDEN, "sparse_d4")
    w = kat["sparse_d4_weights"]
    for t in (O.NumpyTable(4, -1.0), O.CTable(4, -1.0)):
        assert t.load_dir(d) == 37
        keys = np.array([0, 36, 5, 37, -1, 5], dtype=np.int64)
        got = t.lookup(keys)
        assert np.array_equal(got[0], w[0]) and np.array_equal(got[1], w[36]) and np.array_equal(got[2], w[5])
        assert np.all(got[3] == -1.0) and np.all(got[4] == -1.0) and np.array_equal(got[5], w[5])


def test_numpy_and_c_oracle_agree_on_random_tables():
    rng = np.random.default_rng(42)
    for dim, rows, n, parts in [(1, 10, 100, 1), (7, 1000, 5000, 3), (128, 5000, 4096, 8), (32, 0, 64, 8)]:
        keys = rng.choice(1 << 40, size=rows, replace=False).astype(np.int64) - (1 << 39)
        vecs = rng.standard_normal((rows, dim)).astype(np.float32)
        a, b = O.NumpyTable(dim, 2.5), O.CTable(dim, 2.5, num_partitions=parts)
        a.insert(keys, vecs)
        b.insert(keys, vecs)
        q = np.concatenate([rng.choice(keys, size=n // 2), rng.integers(-(1 << 41), 1 << 41, size=n - n // 2)]) if rows else \
            rng.integers(-100, 100, size=n)
        assert np.array_equal(a.lookup(q), b.lookup(q, threads=4))
    a, b = O.NumpyTable(16, 0.0), O.CTable(16, 0.0)
    a.fill_procedural(1000, 5)
    b.fill_procedural(1000, 5, threads=4)
    q = rng.integers(-10, 1100, size=3000)
    assert np.array_equal(a.lookup(q), b.lookup(q, threads=2))


# ------------------------------------------------------------------------------------------------
# engine: host parameter server (CPU path of libhpsx.so) vs oracle and golden
# ------------------------------------------------------------------------------------------------
def test_abi_library_exports_every_declared_symbol():
    L = hb.lib()
    assert L.hpsx_abi_version() == 5
    header = open(os.path.join(ROOT, "include", "hpsx.h")).read()
    declared = set(re.findall(r"\b(hpsx_[a-z0-9_]+)\s*\(", header))
    declared -= {"hpsx_status"}
    assert declared == set(N.SYMBOLS), declared ^ set(N.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.hpsx_device_count() >= 0  # 0 on a CPU box, never an error


def test_engine_cpu_parameter_server_matches_golden(kat):
    hps = hb.HPS(num_partitions=4, num_threads=4)
    hps.add_model(hb.ModelParams("wdl", 64, [1, 16], [2, 26], [float(x) for x in kat["wdl_defaults"]],
                                 use_gpu_embedding_cache=False))
    hps.load_table("wdl", 0, kat["wdl_k0"], kat["wdl_v0"])
    hps.load_table("wdl", 1, kat["wdl_k1"], kat["wdl_v1"])
    assert hps.table_rows("wdl", 0) == 40 and hps.table_rows("wdl", 1) == 300
    keys, nk = kat["wdl_KEYS"].ravel(), kat["wdl_NUMKEYS"].ravel()
    # CPU lookup session: vectors land in host memory (hps.cc:640-642)
    s = hps.session("wdl", device=-1)
    out0 = np.empty((nk[0], 1), dtype=np.float32)
    out1 = np.empty((nk[1], 16), dtype=np.float32)
    s.lookup([keys[:nk[0]], keys[nk[0]:]], [out0, out1], [int(nk[0]), int(nk[1])])
    assert np.array_equal(np.concatenate([out0.ravel(), out1.ravel()]), kat["wdl_OUTPUT0"])
    st = s.stats()
    assert st.keys == 280 and st.default_filled == 2 and st.kernel_launches == 0
    assert np.array_equal(hps.lookup(keys[nk[0]:], "wdl", 1).ravel(), kat["wdl_OUTPUT0"][20:])


def test_engine_loads_reference_format_sparse_files_from_ps_json(tmp_path, kat):
    d = os.path.join(GOLDEN, "sparse_d4")
    ps = {"supportlonglong": "true",
          "volatile_db": {"type": "hash_map", "num_partitions": "3", "allocation_rate": 1 << 20},
          "models": [{"model": "m", "sparse_files": [d], "num_of_worker_buffer_in_pool": "2",
                      "embedding_vecsize_per_table": ["4"], "maxnum_catfeature_query_per_table_per_sample": [3],
                      "default_value_for_each_table": [-1.0], "deployed_device_list": [0], "max_batch_size": 16,
                      "gpucache": False}]}
    path = tmp_path / "ps.json"
    path.write_text(json.dumps(ps))
    hps = hb.HPS(ps_json=str(path))
    assert hps.model_names() == ["m"] and hps.table_rows("m", 0) == 37
    w = kat["sparse_d4_weights"]
    got = hps.lookup(np.array([3, 36, 40, 3]), "m", 0, dim=4)
    assert np.array_equal(got, np.stack([w[3], w[36], np.full(4, -1.0, np.float32), w[3]]))


def test_engine_host_fetch_matches_oracle_at_scale():
    rows, dim, n = 200_000, 32, 300_000
    hps = hb.HPS(num_partitions=8)
    hps.add_model(hb.ModelParams("m", 1 << 20, [dim], [1], [0.5], use_gpu_embedding_cache=False))
    hps.load_table_procedural("m", 0, rows, 77)
    ref = O.NumpyTable(dim, 0.5)
    ref.fill_procedural(rows, 77)
    q = np.random.default_rng(1).integers(-1000, rows + 1000, size=n)
    assert np.array_equal(hps.lookup(q, "m", 0), ref.lookup(q))
    # overwrite some rows: later inserts win (docs/architecture.md:185-218 loader semantics)
    k = np.array([5, 6, 5], dtype=np.int64)
    v = np.arange(3 * dim, dtype=np.float32).reshape(3, dim)
    hps.load_table("m", 0, k, v)
    got = hps.lookup(np.array([5, 6]), "m", 0)
    assert np.array_equal(got[0], v[2]) and np.array_equal(got[1], v[1])


def test_gpu_entry_points_fail_loudly_without_a_gpu():
    if hb.lib().hpsx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    hps = hb.HPS()
    hps.add_model(hb.ModelParams("g", 8, [4], [1], [0.0], use_gpu_embedding_cache=True))
    hps.load_table_procedural("g", 0, 10, 1)
    with pytest.raises(hb.HpsxError) as e:
        hps.create_embedding_cache("g")  # no silent CPU fallback for the GPU path
    assert e.value.code == N.ERR_CUDA
    with pytest.raises(hb.HpsxError):
        hps.session("g", 0)


# ------------------------------------------------------------------------------------------------
# ps.json surface (hps_backend/src/backend.cpp:102-526)
# ------------------------------------------------------------------------------------------------
def _model(**over):
    m = {"model": "m", "sparse_files": [os.path.join(GOLDEN, "sparse_d4")], "num_of_worker_buffer_in_pool": 1,
         "embedding_vecsize_per_table": [4], "maxnum_catfeature_query_per_table_per_sample": [3],
         "default_value_for_each_table": [0.0], "deployed_device_list": [0], "max_batch_size": 16, "gpucache": False}
    m.update(over)
    return {k: v for k, v in m.items() if v is not None}


def _load(tmp_path, cfg):
    p = tmp_path / "ps.json"
    p.write_text(json.dumps(cfg) if not isinstance(cfg, str) else cfg)
    return hb.HPS(ps_json=str(p))


@pytest.mark.parametrize("missing", ["model", "max_batch_size", "sparse_files", "gpucache", "num_of_worker_buffer_in_pool",
                                     "deployed_device_list", "default_value_for_each_table",
                                     "maxnum_catfeature_query_per_table_per_sample", "embedding_vecsize_per_table"])
def test_ps_json_mandatory_keys(tmp_path, missing):
    cfg = {"supportlonglong": True, "models": [_model(**{missing: None})]}
    with pytest.raises(hb.HpsxError) as e:
        _load(tmp_path, cfg)
    assert e.value.code == N.ERR_INVALID_ARG and missing in e.value.message


def test_ps_json_gpucache_needs_threshold_and_percentage(tmp_path):
    for key in ("hit_rate_threshold", "gpucacheper"):
        extra = {"gpucache": True, "hit_rate_threshold": 0.9, "gpucacheper": 0.5, "init_ec": False}
        extra[key] = None
        with pytest.raises(hb.HpsxError) as e:
            _load(tmp_path, {"supportlonglong": True, "models": [_model(**extra)]})
        assert key in e.value.message


def test_ps_json_enum_aliases_and_unsupported_backends(tmp_path):
    for alias in ("parallel_hash_map", "Parallel Hash-Map", "hash_map", "hashmap", "parallel_hashmap"):
        hps = _load(tmp_path, {"supportlonglong": True, "volatile_db": {"type": alias}, "models": [_model()]})
        assert hps.model_names() == ["m"]
    for vt in ("redis_cluster", "redis"):
        with pytest.raises(hb.HpsxError) as e:
            _load(tmp_path, {"supportlonglong": True, "volatile_db": {"type": vt}, "models": [_model()]})
        assert e.value.code == N.ERR_UNSUPPORTED
    with pytest.raises(hb.HpsxError) as e:
        _load(tmp_path, {"supportlonglong": True, "persistent_db": {"type": "rocks_db", "path": "/tmp/x"}, "models": [_model()]})
    assert e.value.code == N.ERR_UNSUPPORTED
    with pytest.raises(hb.HpsxError):
        _load(tmp_path, {"supportlonglong": True, "volatile_db": {"type": "bogus"}, "models": [_model()]})


def test_ps_json_malformed_and_mismatched(tmp_path):
    with pytest.raises(hb.HpsxError) as e:
        _load(tmp_path, "{ not json")
    assert e.value.code == N.ERR_INVALID_ARG
    with pytest.raises(hb.HpsxError):
        hb.HPS(ps_json=str(tmp_path / "nope.json"))
    with pytest.raises(hb.HpsxError):  # per-table lists differ in length
        _load(tmp_path, {"supportlonglong": True, "models": [_model(embedding_vecsize_per_table=[4, 8])]})
    with pytest.raises(hb.HpsxError) as e:  # vector file is not a whole number of rows of this size
        _load(tmp_path, {"supportlonglong": True, "models": [_model(embedding_vecsize_per_table=[5])]})
    assert e.value.code == N.ERR_IO
    with pytest.raises(hb.HpsxError) as e:
        _load(tmp_path, {"supportlonglong": True, "models": [_model(sparse_files=["/no/such/dir"])]})
    assert e.value.code == N.ERR_IO


def test_model_params_round_trip_through_the_abi(tmp_path):
    cfg = {"supportlonglong": True,
           "models": [_model(model="rt", max_batch_size=99, default_value_for_each_table=[-2.5], gpucache=False,
                             num_of_worker_buffer_in_pool=3, deployed_device_list=[1, 0], embedding_table_names=["emb"],
                             enable_pagelock=True, hpsx_peer_tier=True, hpsx_pull_grid_ctas=296, hpsx_request_chunks=2)]}
    hps = _load(tmp_path, cfg)
    p = N.ModelParamsC()
    N.check(hb.lib().hpsx_ps_get_model_params(hps._h, b"rt", ctypes.byref(p)))
    assert p.model_name == b"rt" and p.max_batch_size == 99 and p.num_tables == 1
    assert p.embedding_vecsize_per_table[0] == 4 and p.maxnum_catfeature_query_per_table_per_sample[0] == 3
    assert p.default_value_for_each_table[0] == -2.5 and p.use_gpu_embedding_cache == 0
    assert p.number_of_worker_buffers_in_pool == 3 and p.num_deployed_devices == 2 and p.deployed_devices[1] == 0
    assert p.table_names[0] == b"emb" and p.enable_pagelock == 1
    assert p.sparse_files[0].decode().endswith("sparse_d4")
    # engine extensions of ps.json: the NVLink tier switch and the pull tuning
    assert p.peer_tier == 1 and p.pull_grid_ctas == 296 and p.request_chunks == 2


def test_tier_only_table_spec_is_validated_on_the_host():
    """sparse_files entry "synthetic_device:rows=N,seed=S" (a table that lives in the NVLink tier only, include/hpsx.h):
    needs gpucache + enable_pagelock + hpsx_peer_tier; has no host rows (the CPU path answers with the default vector)."""
    hps = hb.HPS(num_partitions=2)
    with pytest.raises(Exception, match="NVLink tier"):
        hps.add_model(hb.ModelParams("a", 64, [8], [1], [0.5], enable_pagelock=True,
                                     sparse_files=["synthetic_device:rows=100,seed=3"]))
    with pytest.raises(Exception, match="malformed"):
        hps.add_model(hb.ModelParams("b", 64, [8], [1], [0.5], enable_pagelock=True, peer_tier=True,
                                     sparse_files=["synthetic_device:rows=abc"]))
    hps.add_model(hb.ModelParams("c", 64, [8], [1], [0.5], enable_pagelock=True, peer_tier=True, embedding_cache_type="static",
                                 cache_size_percentage=0.0, sparse_files=["synthetic_device:rows=100,seed=3"]))
    assert hps.table_rows("c", 0) == 100
    assert np.array_equal(hps.lookup(np.arange(5), "c", 0), np.full((5, 8), 0.5, dtype=np.float32))
    if hb.lib().hpsx_device_count() == 0:
        with pytest.raises(Exception):  # no CUDA device: the GPU entry points fail, nothing falls back to the CPU
            hps.create_embedding_cache("c")
