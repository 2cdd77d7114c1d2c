"""Property tests (hypothesis) of the host side of the path and of the two oracle implementations: arbitrary key
sets — duplicates, negative keys, INT64_MIN (the cache's empty marker), INT64_MAX — against a plain Python dict.
The reference ships no tests for this path; these pin the lookup contract of
docs/hierarchical_parameter_server.md:67-78,244-246 (a key in no database gets the table's default vector)."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import hugectr_backend_b200 as hb
from oracle import hps_oracle as O

I64 = st.integers(min_value=-(2 ** 63), max_value=2 ** 63 - 1)
SPECIAL = st.sampled_from([-(2 ** 63), 2 ** 63 - 1, 0, -1, 1])
KEYS = st.lists(st.one_of(I64, SPECIAL, st.integers(min_value=-50, max_value=50)), min_size=0, max_size=200)


def dict_lookup(table: dict, keys, dim, default):
    out = np.full((len(keys), dim), default, dtype=np.float32)
    for i, k in enumerate(keys):
        if k in table:
            out[i] = table[k]
    return out


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(loaded=KEYS, queried=KEYS, dim=st.sampled_from([1, 3, 8, 16]), default=st.floats(-4, 4, width=32),
       partitions=st.sampled_from([1, 3, 8]))
def test_host_parameter_server_and_both_oracles_equal_a_dict(loaded, queried, dim, default, partitions):
    rng = np.random.default_rng(len(loaded) * 131 + len(queried))
    lk = np.array(loaded, dtype=np.int64)
    vecs = rng.standard_normal((len(lk), dim)).astype(np.float32)
    table = {}
    for k, v in zip(lk.tolist(), vecs):  # a key loaded twice keeps its last vector (insert or overwrite)
        table[k] = v
    q = np.array(queried + loaded[:5], dtype=np.int64)
    want = dict_lookup(table, q.tolist(), dim, np.float32(default))

    nt = O.NumpyTable(dim, float(np.float32(default)))
    ct = O.CTable(dim, float(np.float32(default)), num_partitions=partitions)
    if len(lk):
        nt.insert(lk, vecs)
        ct.insert(lk, vecs)
    assert np.array_equal(nt.lookup(q), want)
    assert np.array_equal(ct.lookup(q, threads=2), want)

    hps = hb.HPS(num_partitions=partitions, num_threads=2)
    hps.add_model(hb.ModelParams("p", 64, [dim], [4], [float(np.float32(default))], use_gpu_embedding_cache=False))
    if len(lk):
        hps.load_table("p", 0, lk, vecs)
    assert hps.table_rows("p", 0) == len(table)
    assert np.array_equal(hps.lookup(q, "p", 0), want)
    hps.close()


@settings(max_examples=100, deadline=None)
@given(keys=st.lists(I64, min_size=1, max_size=64), shards=st.integers(min_value=1, max_value=64))
def test_owner_is_a_function_of_the_key_and_in_range(keys, shards):
    k = np.array(keys, dtype=np.int64)
    a = hb.hps.owner_batch(k, shards)
    assert a.max() < shards
    assert np.array_equal(a, O.owner(k, shards))
    assert np.array_equal(a, hb.hps.owner_batch(k[::-1].copy(), shards)[::-1])


@settings(max_examples=60, deadline=None)
@given(keys=st.lists(st.integers(min_value=-5, max_value=40), min_size=0, max_size=300))
def test_unique_first_occurrence_inverse_reconstructs_the_keys(keys):
    k = np.array(keys, dtype=np.int64)
    uniq, inv = O.unique_first_occurrence(k)
    assert len(set(uniq.tolist())) == len(uniq)
    assert np.array_equal(uniq[inv], k) if len(k) else len(uniq) == 0
    cu, cinv = O.c_unique(k)
    assert np.array_equal(cu[cinv], k) if len(k) else len(cu) == 0
    assert sorted(cu.tolist()) == sorted(uniq.tolist())
