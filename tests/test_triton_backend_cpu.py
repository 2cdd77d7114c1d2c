"""The Triton `hps` backend shell (libtriton_hps.so) driven through the fake-Triton harness on the CPU
parameter-server path (`gpucache: false`, the reference's configs[0] / ps_cpu.json CI case:
/root/reference/test/triton_server.sh:45-52).  No GPU needed: this is the drop-in boundary's contract —
lifecycle, validation errors, wire format, response parameters, statistics, ownership
(reference hps_backend/src/hps.cc:57-788, src/model_state.cpp:180-432).
"""
import json
import os

import numpy as np
import pytest

from oracle import hps_oracle as O
import fake_triton as FT  # tests/fake_triton (tests/ is on sys.path, see conftest.py)


def write_tables(root, spec, seed=11):
    """spec: [(rows, dim)] -> sparse model dirs <root>/t<i>/{key,emb_vector} + oracle tables (keys are 3*i+1)."""
    rng = np.random.default_rng(seed)
    dirs, tables = [], []
    for i, (rows, dim) in enumerate(spec):
        keys = (np.arange(rows, dtype=np.int64) * 3 + 1)
        vecs = rng.standard_normal((rows, dim)).astype(np.float32)
        d = os.path.join(root, f"t{i}")
        O.write_sparse_dir(d, keys, vecs)
        dirs.append(d)
        tables.append((keys, vecs))
    return dirs, tables


def ps_json(path, models, volatile=None):
    cfg = {"supportlonglong": True,
           "volatile_db": volatile or {"type": "parallel_hash_map", "num_partitions": 8, "initial_cache_rate": 1.0},
           "models": models}
    with open(path, "w") as f:
        json.dump(cfg, f)
    return path


def model_entry(name, dirs, dims, maxq, *, gpucache=False, max_batch=64, defaults=None, workers=2, devices=(0,), **extra):
    """One `models[]` entry with the sample's key set (Hierarchical_Parameter_Server_Deployment.ipynb:282-315)."""
    m = {"model": name, "sparse_files": dirs, "num_of_worker_buffer_in_pool": workers,
         "num_of_refresher_buffer_in_pool": 1, "embedding_table_names": [f"sparse_embedding{i}" for i in range(len(dirs))],
         "embedding_vecsize_per_table": dims, "maxnum_catfeature_query_per_table_per_sample": maxq,
         "default_value_for_each_table": defaults or [0.0] * len(dirs), "deployed_device_list": list(devices),
         "max_batch_size": max_batch, "cache_refresh_percentage_per_iteration": 0.2, "hit_rate_threshold": 0.9,
         "gpucacheper": 0.5, "gpucache": gpucache}
    m.update(extra)
    return m


@pytest.fixture()
def wdl(tmp_path):
    """The sample's Wide&Deep model: two tables, dims [1, 16], 2 + 26 keys per sample."""
    dirs, tables = write_tables(str(tmp_path), [(500, 1), (2000, 16)])
    ps = ps_json(str(tmp_path / "ps.json"), [model_entry("wdl", dirs, [1, 16], [2, 26], defaults=[0.0, 1.5])])
    ref = []
    for (keys, vecs), default in zip(tables, [0.0, 1.5]):
        t = O.NumpyTable(vecs.shape[1], default)
        t.insert(keys, vecs)
        ref.append(t)
    return ps, ref, tables


def wdl_request(tables, samples, rng):
    """KEYS table-major: samples*2 keys of table 0 then samples*26 of table 1 (docs/architecture.md:220-230)."""
    k0 = rng.choice(tables[0][0], size=samples * 2)
    k1 = rng.choice(tables[1][0], size=samples * 26)
    return np.concatenate([k0, k1]), np.array([[samples * 2, samples * 26]], dtype=np.int32)


def test_exports_only_the_seven_entry_points():
    import subprocess
    FT.lib()
    out = subprocess.check_output(["nm", "-D", "--defined-only", FT.BACKEND_SO], text=True)
    names = sorted(line.split()[-1] for line in out.splitlines() if line.strip())
    assert names == sorted(["TRITONBACKEND_Initialize", "TRITONBACKEND_Finalize", "TRITONBACKEND_ModelInitialize",
                            "TRITONBACKEND_ModelFinalize", "TRITONBACKEND_ModelInstanceInitialize",
                            "TRITONBACKEND_ModelInstanceFinalize", "TRITONBACKEND_ModelInstanceExecute"])


def test_wdl_request_matches_oracle_cpu_path(wdl):
    ps, ref, tables = wdl
    errors0 = FT.live_errors()
    with FT.Backend(ps) as be:
        model = be.model("wdl", FT.model_config("wdl", kind="KIND_CPU", max_batch_size=64))
        inst = model.instance(kind=FT.KIND_CPU)
        rng = np.random.default_rng(0)
        keys, numkeys = wdl_request(tables, 10, rng)
        r = inst.infer(keys, numkeys)
        assert r.error_code is None, r.error_message
        assert r.shape == [4180] and r.byte_size == 4180 * 4  # 10*2*1 + 10*26*16 (Deployment.ipynb:793-795)
        assert r.memory_type == FT.MEM_CPU and r.output_name == "OUTPUT0"
        assert np.array_equal(r.data, O.request(ref, keys, numkeys.ravel()))
        assert r.params == {"NumSample": 10, "DeviceID": 0}
        assert r.sent == 1 and r.flags == FT.RESPONSE_COMPLETE_FINAL and r.released == 1
        st = inst.stats()
        assert st["ok_requests"] == 1 and st["failed_requests"] == 0 and st["batch_reports"] == 1
        assert st["last_batch_size"] == 10  # the reference always reports 0 here
        assert st["exec_start"] <= st["compute_start"] <= st["compute_end"] <= st["exec_end"]
        inst.close()
        model.close()
    assert FT.live_errors() == errors0 and FT.live_messages() == 0


def test_absent_keys_duplicates_and_empty_table_slice(wdl):
    ps, ref, tables = wdl
    with FT.Backend(ps) as be:
        model = be.model("wdl", FT.model_config("wdl", kind="KIND_CPU"))
        inst = model.instance(kind=FT.KIND_CPU)
        # table 0 gets no keys at all, table 1 gets duplicates and keys that were never loaded (-> default 1.5)
        k1 = np.array([1, 1, 1, 4, 0, 2, 10**12, -5, 4], dtype=np.int64)
        r = inst.infer(k1, np.array([[0, len(k1)]], dtype=np.int32))
        assert r.error_code is None, r.error_message
        expect = O.request(ref, k1, [0, len(k1)])
        assert r.shape == [len(k1) * 16] and np.array_equal(r.data, expect)
        got = r.data.reshape(len(k1), 16)
        assert np.all(got[4] == 1.5) and np.all(got[6] == 1.5) and np.array_equal(got[0], got[1])
        assert r.params["NumSample"] == 0
        # NUMKEYS may name fewer tables than the model has: only table 0 here
        k0 = np.array([1, 4, 7, 2], dtype=np.int64)
        r = inst.infer(k0, np.array([[4]], dtype=np.int32))
        assert r.error_code is None and np.array_equal(r.data, O.request(ref[:1], k0, [4]))
        # a request with zero keys is served with an empty tensor
        r = inst.infer(np.empty(0, dtype=np.int64), np.array([[0, 0]], dtype=np.int32))
        assert r.error_code is None and r.shape == [0] and r.data.size == 0
        inst.close()
        model.close()


def test_multi_buffer_keys_and_many_requests_per_execute(wdl):
    ps, ref, tables = wdl
    with FT.Backend(ps) as be:
        model = be.model("wdl", FT.model_config("wdl", kind="KIND_CPU"))
        inst = model.instance(kind=FT.KIND_CPU)
        rng = np.random.default_rng(5)
        reqs, expect = [], []
        for i, samples in enumerate([1, 7, 64, 3]):
            keys, numkeys = wdl_request(tables, samples, rng)
            reqs.append(dict(keys=keys, numkeys=numkeys, key_buffers=1 + i, request_id=f"r{i}"))
            expect.append(O.request(ref, keys, numkeys.ravel()))
        out = inst.infer_many(reqs)
        for r, e, samples in zip(out, expect, [1, 7, 64, 3]):
            assert r.error_code is None, r.error_message
            assert np.array_equal(r.data, e) and r.params["NumSample"] == samples and r.released == 1 and r.sent == 1
        st = inst.stats()
        assert st["ok_requests"] == 4 and st["batch_reports"] == 1 and st["last_batch_size"] == 75
        inst.close()
        model.close()


def test_per_request_errors_do_not_poison_the_batch(wdl):
    ps, ref, tables = wdl
    errors0 = FT.live_errors()
    with FT.Backend(ps) as be:
        model = be.model("wdl", FT.model_config("wdl", kind="KIND_CPU"))
        inst = model.instance(kind=FT.KIND_CPU)
        rng = np.random.default_rng(9)
        good_keys, good_nk = wdl_request(tables, 4, rng)
        big_keys, big_nk = wdl_request(tables, 65, rng)  # max_batch_size is 64
        bad_sum = dict(keys=good_keys, numkeys=np.array([[8, 100]], dtype=np.int32))
        neg = dict(keys=good_keys, numkeys=np.array([[-1, len(good_keys) + 1]], dtype=np.int32))
        too_many_tables = dict(keys=good_keys, numkeys=np.array([[8, 100, 4]], dtype=np.int32))
        wrong_name = dict(keys=good_keys, numkeys=good_nk, input_names=("CATCOLUMN", "NUMKEYS"))
        wrong_type = dict(keys=good_keys, numkeys=good_nk, keys_dtype=FT.TYPE_INT32)
        no_buffer = dict(keys=good_keys, numkeys=good_nk, fail_output_buffer=True)
        out = inst.infer_many([dict(keys=good_keys, numkeys=good_nk), dict(keys=big_keys, numkeys=big_nk), bad_sum, neg,
                               too_many_tables, wrong_name, wrong_type, no_buffer, dict(keys=good_keys, numkeys=good_nk)])
        ok = O.request(ref, good_keys, good_nk.ravel())
        assert out[0].error_code is None and np.array_equal(out[0].data, ok)
        assert out[8].error_code is None and np.array_equal(out[8].data, ok)
        # oversize batch: UNSUPPORTED with the reference's message (hps.cc:576-582) — and no crash (Appendix B.2)
        assert out[1].error_code == FT.ERR["UNSUPPORTED"]
        assert "The number of Input samples greater than max batch size" in out[1].error_message
        for r in out[2:5]:
            assert r.error_code == FT.ERR["INVALID_ARG"] and "NUMKEYS" in r.error_message
        assert out[5].error_code == FT.ERR["INVALID_ARG"] and "expected input name as KEYS and NUMKEYS" in out[5].error_message
        assert out[6].error_code == FT.ERR["INVALID_ARG"] and "TYPE_INT64" in out[6].error_message
        assert out[7].error_code == FT.ERR["INTERNAL"]
        for r in out:  # exactly one FINAL response and one release per request, whatever happened
            assert r.sent == 1 and r.flags == FT.RESPONSE_COMPLETE_FINAL and r.released == 1
        st = inst.stats()
        assert st["ok_requests"] == 2 and st["failed_requests"] == 7 and st["last_batch_size"] == 8
        inst.close()
        model.close()
    assert FT.live_errors() == errors0  # every TRITONSERVER_Error the backend created was deleted


def test_request_without_requested_output_still_gets_a_response(wdl):
    ps, ref, tables = wdl
    with FT.Backend(ps) as be:
        model = be.model("wdl", FT.model_config("wdl", kind="KIND_CPU"))
        inst = model.instance(kind=FT.KIND_CPU)
        keys, nk = wdl_request(tables, 2, np.random.default_rng(1))
        r = inst.infer(keys, nk, requested_output=None)
        assert r.error_code is None and r.shape is None and r.sent == 1 and r.released == 1
        assert r.params == {"NumSample": 2, "DeviceID": 0}
        inst.close()
        model.close()


def test_api_version_gate(wdl):
    ps, _, _ = wdl
    # server major differs, or server minor older than the backend's -> UNSUPPORTED (hps.cc:77-82)
    for version in [(2, 10), (1, 9), (0, 99)]:
        with pytest.raises(FT.TritonError) as e:
            FT.Backend(ps, api_version=version)
        assert e.value.code == FT.ERR["UNSUPPORTED"]
        assert "Triton backend API version does not support this backend" in e.value.message
    be = FT.Backend(ps, api_version=(1, 17))  # newer minor is fine
    be.close()


def test_backend_config_without_ps_path_is_rejected(tmp_path):
    with pytest.raises(FT.TritonError) as e:
        FT.Backend(None)
    assert e.value.code == FT.ERR["INVALID_ARG"] and "ps" in e.value.message
    with pytest.raises(FT.TritonError) as e:
        FT.Backend(str(tmp_path / "missing.json"))
    assert e.value.code == FT.ERR["INVALID_ARG"]


@pytest.mark.parametrize("mutate,code,needle", [
    (lambda c: c["input"].pop(), "INVALID_ARG", "expect 2 input, got 1"),
    (lambda c: c["input"][0].update(name="CATCOLUMN"), "INVALID_ARG", "expected input name as KEYS and NUMKEYS, but got CATCOLUMN"),
    (lambda c: c["input"][0].update(data_type="TYPE_INT32"), "INVALID_ARG", "expected KEYS input datatype as TYPE_INT64, got TYPE_INT32"),
    (lambda c: c["input"][1].update(data_type="TYPE_FP32"), "INVALID_ARG", "expected NUMKEYS input datatype as TYPE_INT32, got TYPE_FP32"),
    (lambda c: c["input"][0].update(dims=[26]), "INVALID_ARG", "expected input shape equal -1, got [26]"),
    (lambda c: c["output"].append(dict(c["output"][0])), "INVALID_ARG", "expect 1 output, got 2"),
    (lambda c: c["output"][0].update(data_type="TYPE_FP16"), "INVALID_ARG", "output datatype as TYPE_FP32, got TYPE_FP16"),
    (lambda c: c["output"][0].update(dims=[4180]), "INVALID_ARG", "output shape equal -1, got [4180]"),
    (lambda c: c.update(instance_group=[]), "INVALID_ARG", "expect at least one instance in instance group"),
    (lambda c: c["instance_group"][0].update(count=3), "INVALID_ARG", "num_of_worker_buffer_in_pool"),
    (lambda c: c.update(max_batch_size=-1), "INVALID_ARG", "max_batch_size"),
])
def test_model_config_validation_errors(wdl, mutate, code, needle):
    """Same checks and (where the reference's text is not itself wrong) the same wording as
    ModelState::ValidateModelConfig / ParseModelConfig (src/model_state.cpp:180-371)."""
    ps, _, _ = wdl
    cfg = FT.model_config("wdl", kind="KIND_CPU")
    mutate(cfg)
    with FT.Backend(ps) as be:
        with pytest.raises(FT.TritonError) as e:
            be.model("wdl", cfg)
        assert e.value.code == FT.ERR[code] and needle in e.value.message, e.value.message


def test_two_dim_inputs_and_string_encoded_numbers(wdl):
    """max_batch_size 0 with explicit [-1,-1] dims (02_model_inference_hps_tf_ensemble.ipynb:155-165) and
    protobuf-JSON style string numbers are both accepted."""
    ps, ref, tables = wdl
    cfg = FT.model_config("wdl", kind="KIND_CPU", two_dims=True)
    cfg["input"][0]["dims"] = ["-1", "-1"]
    cfg["instance_group"][0]["count"] = "2"
    cfg["parameters"] = {"refresh_interval": {"string_value": "0"}, "refresh_delay": {"string_value": "0"},
                         "freeze_sparse": {"string_value": "false"}}
    with FT.Backend(ps) as be:
        model = be.model("wdl", cfg)
        inst = model.instance(kind=FT.KIND_CPU)
        keys, nk = wdl_request(tables, 3, np.random.default_rng(2))
        r = inst.infer(keys, nk)
        assert r.error_code is None and np.array_equal(r.data, O.request(ref, keys, nk.ravel()))
        inst.close()
        model.close()


def test_unknown_model_and_online_deployment(tmp_path):
    """A model missing from ps.json is rejected; once ps.json gains it, loading it re-reads the file
    (online deployment: hps.cc:207-219, README.md:161-169) without touching the models already served."""
    dirs_a, tabs_a = write_tables(str(tmp_path / "a"), [(100, 8)])
    dirs_b, tabs_b = write_tables(str(tmp_path / "b"), [(50, 4)], seed=3)
    ps = ps_json(str(tmp_path / "ps.json"), [model_entry("a", dirs_a, [8], [3])])
    with FT.Backend(ps) as be:
        ma = be.model("a", FT.model_config("a", kind="KIND_CPU"))
        with pytest.raises(FT.TritonError) as e:
            be.model("b", FT.model_config("b", kind="KIND_CPU"))
        assert e.value.code == FT.ERR["INVALID_ARG"] and "has been added to the Parameter Server json" in e.value.message
        ps_json(ps, [model_entry("a", dirs_a, [8], [3]), model_entry("b", dirs_b, [4], [2], defaults=[-1.0])])
        mb = be.model("b", FT.model_config("b", kind="KIND_CPU"), version=2)
        ia, ib = ma.instance(kind=FT.KIND_CPU), mb.instance(kind=FT.KIND_CPU)
        ka = np.array([1, 4, 298, 5], dtype=np.int64)
        ra = ia.infer(ka, np.array([[4]], dtype=np.int32))
        ta = O.NumpyTable(8, 0.0)
        ta.insert(*tabs_a[0])
        assert ra.error_code is None and np.array_equal(ra.data, O.request([ta], ka, [4]))
        kb = np.array([1, 148, 3, 151], dtype=np.int64)
        rb = ib.infer(kb, np.array([[4]], dtype=np.int32))
        tb = O.NumpyTable(4, -1.0)
        tb.insert(*tabs_b[0])
        assert rb.error_code is None and np.array_equal(rb.data, O.request([tb], kb, [4]))
        for x in (ia, ib, ma, mb):
            x.close()


def test_gpucache_model_needs_gpu_instances_and_deployed_devices(tmp_path):
    """Config errors of GPU-cache models surface at ModelInitialize (model_state.cpp:287-290, 395-402).  On a box
    without a GPU the backend itself cannot start such a model: creating the cache must fail loudly, not fall back."""
    dirs, _ = write_tables(str(tmp_path), [(100, 8)])
    ps = ps_json(str(tmp_path / "ps.json"), [model_entry("g", dirs, [8], [3], gpucache=True, init_ec=False, devices=[0])])
    import hugectr_backend_b200 as hb
    have_gpu = hb.lib().hpsx_device_count() > 0
    with FT.Backend(ps) as be:
        with pytest.raises(FT.TritonError) as e:
            be.model("g", FT.model_config("g", kind="KIND_CPU"))
        assert e.value.code == FT.ERR["INVALID_ARG"] and "expect GPU kind instance in instance group" in e.value.message
        with pytest.raises(FT.TritonError) as e:
            be.model("g", FT.model_config("g", gpus=[3]))
        assert e.value.code == FT.ERR["INVALID_ARG"]
        assert "Please confirm that device 3 is added to 'deployed_device_list'" in e.value.message
        if not have_gpu:
            with pytest.raises(FT.TritonError) as e:
                be.model("g", FT.model_config("g", gpus=[0]))
            assert "CUDA" in e.value.message  # no silent CPU fallback for a gpucache model


@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_opt_in_slot_pooling_matches_golden(tmp_path, mode):
    """North-star stage a8 behind the Triton boundary: config.pbtxt parameters hps_pooling / hps_pooling_hotness turn
    OUTPUT0 into the slot-wise reduced tensor.  Golden: the ensemble sample's request (64 samples x 3 keys, dim 16,
    default 1.0 for absent keys), reduced longhand in tests/golden/make_golden.py."""
    kat = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hps_kat.npz"))
    d = str(tmp_path / "ens")
    O.write_sparse_dir(d, kat["ens_keys"], kat["ens_vecs"])
    ps = ps_json(str(tmp_path / "ps.json"), [model_entry("ens", [d], [16], [3], defaults=[1.0], max_batch=64)])
    with FT.Backend(ps) as be:
        cfg = FT.model_config("ens", kind="KIND_CPU", parameters={"hps_pooling": mode, "hps_pooling_hotness": "3"})
        model = be.model("ens", cfg)
        inst = model.instance(kind=FT.KIND_CPU)
        r = inst.infer(kat["ens_KEYS"], kat["ens_NUMKEYS"])
        assert r.error_code is None, r.error_message
        assert r.shape == [64 * 16] and r.params["NumSample"] == 64
        assert np.array_equal(r.data.reshape(64, 16), kat[f"ens_pooled_{mode}"])
        # a key count that is not a multiple of the hotness is a request error
        r = inst.infer(kat["ens_KEYS"].ravel()[:10], np.array([[10]], dtype=np.int32))
        assert r.error_code == FT.ERR["INVALID_ARG"] and "pooling hotness" in r.error_message
        inst.close()
        model.close()
    with FT.Backend(ps) as be:
        with pytest.raises(FT.TritonError) as e:
            be.model("ens", FT.model_config("ens", kind="KIND_CPU", parameters={"hps_pooling": "max"}))
        assert e.value.code == FT.ERR["INVALID_ARG"]
        # without the parameters the reference's un-pooled contract applies
        model = be.model("ens", FT.model_config("ens", kind="KIND_CPU"))
        inst = model.instance(kind=FT.KIND_CPU)
        r = inst.infer(kat["ens_KEYS"], kat["ens_NUMKEYS"])
        assert r.shape == [3072] and np.array_equal(r.data, kat["ens_OUTPUT0"])
        inst.close()
        model.close()


@pytest.mark.parametrize("freeze", [False, True])
def test_new_version_reloads_the_database_unless_frozen(tmp_path, freeze):
    """Loading version 2 of a model that is already served re-reads its sparse files in the background
    (update_database_per_model, hps_backend/src/model_state.cpp:124-143,413-420) unless `freeze_sparse`."""
    root = str(tmp_path)
    dirs, tabs = write_tables(root, [(200, 8)], seed=1)
    keys_v1, vecs_v1 = tabs[0]
    ps = ps_json(str(tmp_path / "ps.json"), [model_entry("m", dirs, [8], [4], defaults=[-2.0])])
    params = {"freeze_sparse": "true"} if freeze else None
    with FT.Backend(ps) as be:
        m1 = be.model("m", FT.model_config("m", kind="KIND_CPU", parameters=params), version=1)
        i1 = m1.instance(kind=FT.KIND_CPU)
        q = np.array([1, 4, 598, 601, 1000], dtype=np.int64)  # 601 and 1000 are not in version 1
        nk = np.array([[len(q)]], dtype=np.int32)
        t1 = O.NumpyTable(8, -2.0)
        t1.insert(keys_v1, vecs_v1)
        assert np.array_equal(i1.infer(q, nk).data, O.request([t1], q, [len(q)]))
        # the trainer dumps a new model: every vector changes and key 601 appears
        keys_v2 = np.concatenate([keys_v1, np.array([601], dtype=np.int64)])
        vecs_v2 = np.random.default_rng(2).standard_normal((len(keys_v2), 8)).astype(np.float32)
        O.write_sparse_dir(dirs[0], keys_v2, vecs_v2)
        t2 = O.NumpyTable(8, -2.0)
        t2.insert(keys_v2, vecs_v2)
        m2 = be.model("m", FT.model_config("m", kind="KIND_CPU", parameters=params), version=2)
        i2 = m2.instance(kind=FT.KIND_CPU)
        i2.close()
        m2.close()  # ModelFinalize joins the background reload
        want = t1 if freeze else t2
        assert np.array_equal(i1.infer(q, nk).data, O.request([want], q, [len(q)]))
        i1.close()
        m1.close()


def test_periodic_refresh_thread_lifecycle_on_a_cpu_model(tmp_path):
    """`refresh_interval` starts the model's refresh timer (model_state.cpp:422-427); on a CPU model there is no cache to
    refresh, the thread just ticks beside the lookups and ModelFinalize stops and joins it promptly."""
    import time
    dirs, tabs = write_tables(str(tmp_path), [(120, 4)], seed=5)
    ps = ps_json(str(tmp_path / "ps.json"), [model_entry("m", dirs, [4], [3], defaults=[0.5])])
    t = O.NumpyTable(4, 0.5)
    t.insert(*tabs[0])
    with FT.Backend(ps) as be:
        m = be.model("m", FT.model_config("m", kind="KIND_CPU", parameters={"refresh_interval": "0.02", "refresh_delay": "0"}))
        inst = m.instance(kind=FT.KIND_CPU)
        q = np.array([1, 4, 7, 100000], dtype=np.int64)
        for _ in range(10):
            r = inst.infer(q, np.array([[4]], dtype=np.int32))
            assert r.error_code is None and np.array_equal(r.data, O.request([t], q, [4]))
            time.sleep(0.01)
        inst.close()
        t0 = time.perf_counter()
        m.close()
        assert time.perf_counter() - t0 < 1.0
