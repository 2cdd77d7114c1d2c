"""libtriton_hps.so end to end on the GPU, driven through the fake-Triton harness:
TRITONBACKEND_ModelInstanceExecute -> hpsx C ABI -> sm_100a kernels writing straight into the "Triton"
output buffer.  Parity against the CPU oracle is bit-exact (the path only copies fp32 rows).
Reference flow: hps_backend/src/hps.cc:348-788, src/model_instance_state.cpp:176-197.
"""
import json
import os

import numpy as np
import pytest

import fake_triton as FT  # tests/fake_triton
from oracle import hps_oracle as O
from test_triton_backend_cpu import model_entry, ps_json, wdl_request, write_tables

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["staged", "direct"])
def wdl_gpu(request, tmp_path, cuda_device):
    dirs, tables = write_tables(str(tmp_path), [(500, 1), (20000, 16)])
    entry = model_entry("wdl", dirs, [1, 16], [2, 26], gpucache=True, defaults=[0.0, 1.5], max_batch=1024,
                        hit_rate_threshold=1.0, gpucacheper=0.3, enable_pagelock=(request.param == "direct"))
    ps = ps_json(str(tmp_path / "ps.json"), [entry])
    ref = []
    for (keys, vecs), default in zip(tables, [0.0, 1.5]):
        t = O.NumpyTable(vecs.shape[1], default)
        t.insert(keys, vecs)
        ref.append(t)
    return ps, ref, tables


def test_gpu_output_buffer_is_written_in_place(wdl_gpu):
    import torch
    ps, ref, tables = wdl_gpu
    errors0 = FT.live_errors()
    with FT.Backend(ps) as be:
        model = be.model("wdl", FT.model_config("wdl", gpus=[0], max_batch_size=1024))
        inst = model.instance(kind=FT.KIND_GPU, device=0)
        rng = np.random.default_rng(0)
        for samples in (10, 1, 1024, 333):
            keys, numkeys = wdl_request(tables, samples, rng)
            keys[::7] = 10**15 + np.arange(len(keys[::7]))  # keys in no table -> default vector
            n_out = samples * 2 * 1 + samples * 26 * 16
            out = torch.full((n_out + 64,), float("nan"), device="cuda")
            r = inst.infer(keys, numkeys, gpu_out=out)
            assert r.error_code is None, r.error_message
            assert r.memory_type == FT.MEM_GPU and r.device_ptr == out.data_ptr() and r.shape == [n_out]
            got = out.cpu().numpy()
            assert np.array_equal(got[:n_out], O.request(ref, keys, numkeys.ravel()))
            assert np.isnan(got[n_out:]).all()  # nothing written past the tensor
            assert r.params == {"NumSample": samples, "DeviceID": 0} and r.sent == 1 and r.released == 1
        inst.close()
        model.close()
    assert FT.live_errors() == errors0 and FT.live_messages() == 0


def test_triton_may_hand_back_cpu_memory_for_the_output(wdl_gpu):
    """OutputBuffer's memory type is only a preference (hps.cc:638-648): with a CPU buffer the rows are gathered on
    the GPU and copied D2H (hps.cc:681-685)."""
    ps, ref, tables = wdl_gpu
    with FT.Backend(ps) as be:
        model = be.model("wdl", FT.model_config("wdl", gpus=[0]))
        inst = model.instance(kind=FT.KIND_GPU, device=0)
        keys, numkeys = wdl_request(tables, 77, np.random.default_rng(4))
        r = inst.infer(keys, numkeys)  # no GPU buffer offered -> harness answers with CPU memory
        assert r.error_code is None, r.error_message
        assert r.memory_type == FT.MEM_CPU
        assert np.array_equal(r.data, O.request(ref, keys, numkeys.ravel()))
        inst.close()
        model.close()


def test_inputs_already_in_gpu_memory_and_split_buffers(wdl_gpu):
    import torch
    ps, ref, tables = wdl_gpu
    with FT.Backend(ps) as be:
        model = be.model("wdl", FT.model_config("wdl", gpus=[0]))
        inst = model.instance(kind=FT.KIND_GPU, device=0)
        keys, numkeys = wdl_request(tables, 50, np.random.default_rng(8))
        expect = O.request(ref, keys, numkeys.ravel())
        out = torch.empty(len(expect), device="cuda")
        d_keys = torch.from_numpy(keys).cuda()
        d_nk = torch.from_numpy(numkeys).cuda()
        r = inst.infer(keys, numkeys, gpu_out=out, keys_device_ptr=d_keys.data_ptr(), numkeys_device_ptr=d_nk.data_ptr())
        assert r.error_code is None, r.error_message
        assert np.array_equal(out.cpu().numpy(), expect)
        out.zero_()
        r = inst.infer(keys, numkeys, gpu_out=out, key_buffers=5)
        assert r.error_code is None and np.array_equal(out.cpu().numpy(), expect)
        inst.close()
        model.close()


def test_two_instances_share_one_cache_and_errors_stay_per_request(wdl_gpu):
    import threading
    import torch
    ps, ref, tables = wdl_gpu
    with FT.Backend(ps) as be:
        cfg = FT.model_config("wdl", gpus=[0], count=2)
        model = be.model("wdl", cfg)
        insts = [model.instance(name=f"wdl_0_{i}", kind=FT.KIND_GPU, device=0) for i in range(2)]
        results = {}

        def work(i):
            rng = np.random.default_rng(100 + i)
            ok = True
            for _ in range(20):
                keys, numkeys = wdl_request(tables, 64, rng)
                out = torch.empty(64 * 2 + 64 * 26 * 16, device="cuda")
                r = insts[i].infer(keys, numkeys, gpu_out=out)
                ok &= r.error_code is None and np.array_equal(out.cpu().numpy(), O.request(ref, keys, numkeys.ravel()))
            results[i] = ok

        threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
        [t.start() for t in threads]
        [t.join() for t in threads]
        assert results == {0: True, 1: True}
        # an oversize request fails alone; the instance keeps serving
        big_keys, big_nk = wdl_request(tables, 1025, np.random.default_rng(1))
        r = insts[0].infer(big_keys, big_nk)
        assert r.error_code == FT.ERR["UNSUPPORTED"]
        keys, numkeys = wdl_request(tables, 3, np.random.default_rng(2))
        r = insts[0].infer(keys, numkeys)
        assert r.error_code is None and np.array_equal(r.data, O.request(ref, keys, numkeys.ravel()))
        for i in insts:
            i.close()
        model.close()


def test_instance_on_undeployed_device_is_rejected(wdl_gpu):
    ps, _, _ = wdl_gpu
    with FT.Backend(ps) as be:
        model = be.model("wdl", FT.model_config("wdl", gpus=[0]))
        with pytest.raises(FT.TritonError) as e:
            model.instance(kind=FT.KIND_GPU, device=5)
        assert e.value.code == FT.ERR["INVALID_ARG"]
        with pytest.raises(FT.TritonError):
            model.instance(kind=FT.KIND_CPU, device=0)
        model.close()


@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_opt_in_slot_pooling_on_gpu(wdl_gpu, mode):
    import torch
    ps, ref, tables = wdl_gpu
    with FT.Backend(ps) as be:
        cfg = FT.model_config("wdl", gpus=[0], parameters={"hps_pooling": mode, "hps_pooling_hotness": "2,13"})
        model = be.model("wdl", cfg)
        inst = model.instance(kind=FT.KIND_GPU, device=0)
        samples = 200
        keys, numkeys = wdl_request(tables, samples, np.random.default_rng(6))
        n0, n1 = int(numkeys[0, 0]), int(numkeys[0, 1])
        expect = np.concatenate([O.pooled(ref[0], keys[:n0], n0 // 2, 2, mode).ravel(),
                                 O.pooled(ref[1], keys[n0:], n1 // 13, 13, mode).ravel()])
        out = torch.full((len(expect),), float("nan"), device="cuda")
        r = inst.infer(keys, numkeys, gpu_out=out)
        assert r.error_code is None, r.error_message
        assert r.shape == [len(expect)]
        np.testing.assert_allclose(out.cpu().numpy(), expect, rtol=0, atol=1e-5)  # north-star tolerance on fp32 slot sums
        r = inst.infer(keys, numkeys)  # CPU output buffer
        np.testing.assert_allclose(r.data, expect, rtol=0, atol=1e-5)
        inst.close()
        model.close()


def test_version_reload_and_periodic_refresh_on_gpu(tmp_path, cuda_device):
    """f3 through the Triton boundary: version 2 of a served model reloads the sparse files and refreshes the HBM
    cache in the background (model_state.cpp:124-143,413-420); `refresh_interval` keeps a periodic refresh thread
    alive beside lookups (model_state.cpp:145-178,422-427) and ModelFinalize stops it."""
    import time

    import torch
    dirs, tabs = write_tables(str(tmp_path), [(3000, 16)], seed=4)
    keys1, vecs1 = tabs[0]
    entry = model_entry("m", dirs, [16], [8], gpucache=True, defaults=[0.25], max_batch=256, hit_rate_threshold=1.0,
                        gpucacheper=1.0, enable_pagelock=True)
    ps = ps_json(str(tmp_path / "ps.json"), [entry])
    with FT.Backend(ps) as be:
        cfg = FT.model_config("m", gpus=[0], max_batch_size=256, parameters={"refresh_interval": "0.05"})
        m1 = be.model("m", cfg, version=1)
        i1 = m1.instance(kind=FT.KIND_GPU, device=0)
        rng = np.random.default_rng(1)
        q = rng.choice(keys1, size=2048)
        nk = np.array([[len(q)]], dtype=np.int32)
        out = torch.empty(len(q) * 16, device="cuda")
        t1 = O.NumpyTable(16, 0.25)
        t1.insert(keys1, vecs1)
        for _ in range(20):  # lookups race the periodic refresh; values never change because the database did not
            r = i1.infer(q, nk, gpu_out=out)
            assert r.error_code is None, r.error_message
            assert np.array_equal(out.cpu().numpy(), O.request([t1], q, [len(q)]))
            time.sleep(0.01)
        vecs2 = np.random.default_rng(9).standard_normal(vecs1.shape).astype(np.float32)
        O.write_sparse_dir(dirs[0], keys1, vecs2)
        t2 = O.NumpyTable(16, 0.25)
        t2.insert(keys1, vecs2)
        m2 = be.model("m", cfg, version=2)
        i2 = m2.instance(kind=FT.KIND_GPU, device=0)
        i2.close()
        m2.close()  # joins the background reload + refresh
        r = i1.infer(q, nk, gpu_out=out)
        assert r.error_code is None, r.error_message
        assert np.array_equal(out.cpu().numpy(), O.request([t2], q, [len(q)]))
        i1.close()
        m1.close()


def test_many_requests_per_execute_are_served_in_one_pass(wdl_gpu):
    """f4 cross-request batching: one Execute call with several requests (the reference serves them one by one,
    hps.cc:392-406).  GPU-output requests share one engine pass; a bad request in the middle, a request whose
    output lands in CPU memory and a batch that exceeds one request's key budget all still get exact answers."""
    import torch
    ps, ref, tables = wdl_gpu
    with FT.Backend(ps) as be:
        model = be.model("wdl", FT.model_config("wdl", gpus=[0], max_batch_size=1024))
        inst = model.instance(kind=FT.KIND_GPU, device=0)
        rng = np.random.default_rng(5)
        for sizes in ([7, 1, 64, 33, 2], [600, 500, 400], [1] * 20):
            reqs, outs, want = [], [], []
            for i, samples in enumerate(sizes):
                keys, numkeys = wdl_request(tables, samples, rng)
                keys[::5] = 10**14 + np.arange(len(keys[::5]))
                want.append(O.request(ref, keys, numkeys.ravel()))
                if i == 1 and len(sizes) == 5:  # this one gets a CPU output buffer: it cannot join the fused pass
                    reqs.append(dict(keys=keys, numkeys=numkeys))
                    outs.append(None)
                else:
                    out = torch.full((len(want[-1]),), float("nan"), device="cuda")
                    reqs.append(dict(keys=keys, numkeys=numkeys, gpu_out=out))
                    outs.append(out)
            if len(sizes) == 5:  # an invalid request in the middle fails alone
                bad_keys, bad_nk = wdl_request(tables, 3, rng)
                reqs.insert(3, dict(keys=bad_keys, numkeys=bad_nk + 1))
                outs.insert(3, "bad")
                want.insert(3, None)
            rs = inst.infer_many(reqs)
            assert len(rs) == len(reqs)
            for r, out, w in zip(rs, outs, want):
                if w is None:
                    assert r.error_code == FT.ERR["INVALID_ARG"] and r.sent == 1 and r.released == 1
                    continue
                assert r.error_code is None, r.error_message
                got = r.data if out is None else out.cpu().numpy()
                assert np.array_equal(got, w)
                assert r.sent == 1 and r.released == 1
        st = inst.stats()
        assert st["failed_requests"] == 1 and st["ok_requests"] == 5 + 3 + 20
        inst.close()
        model.close()


def test_two_models_with_two_instances_each_run_concurrently(tmp_path, cuda_device):
    """configs[4] shape (multi-model HPS): a DCN-like and a Wide&Deep-like model on one parameter server, two
    instances per model sharing that model's cache (include/model_state.hpp:76-84), four Triton worker threads
    issuing mixed batch sizes at once.  Every response is bit-exact."""
    import threading

    import torch
    dirs_a, tabs_a = write_tables(str(tmp_path / "dcn"), [(30000, 32)], seed=21)
    dirs_b, tabs_b = write_tables(str(tmp_path / "wdl"), [(400, 1), (15000, 16)], seed=22)
    ps = ps_json(str(tmp_path / "ps.json"), [
        model_entry("dcn", dirs_a, [32], [26], gpucache=True, max_batch=512, hit_rate_threshold=1.0, gpucacheper=0.2,
                    enable_pagelock=True, workers=2),
        model_entry("wdl", dirs_b, [1, 16], [2, 26], gpucache=True, defaults=[0.0, 1.5], max_batch=512,
                    hit_rate_threshold=1.0, gpucacheper=0.3, workers=2)])
    ref_a = [O.NumpyTable(32, 0.0)]
    ref_a[0].insert(*tabs_a[0])
    ref_b = []
    for (k, v), d in zip(tabs_b, [0.0, 1.5]):
        t = O.NumpyTable(v.shape[1], d)
        t.insert(k, v)
        ref_b.append(t)
    with FT.Backend(ps) as be:
        ma = be.model("dcn", FT.model_config("dcn", gpus=[0], count=2, max_batch_size=512))
        mb = be.model("wdl", FT.model_config("wdl", gpus=[0], count=2, max_batch_size=512))
        insts = [ma.instance(name="dcn_0_0", kind=FT.KIND_GPU, device=0), ma.instance(name="dcn_0_1", kind=FT.KIND_GPU, device=0),
                 mb.instance(name="wdl_0_0", kind=FT.KIND_GPU, device=0), mb.instance(name="wdl_0_1", kind=FT.KIND_GPU, device=0)]
        results = {}

        def work(i):
            rng = np.random.default_rng(50 + i)
            ok = True
            for it in range(25):
                samples = int(rng.choice([1, 8, 64, 512]))
                if i < 2:
                    keys = rng.choice(tabs_a[0][0], size=samples * 26)
                    numkeys = np.array([[samples * 26]], dtype=np.int32)
                    want = O.request(ref_a, keys, numkeys.ravel())
                else:
                    keys, numkeys = wdl_request(tabs_b, samples, rng)
                    want = O.request(ref_b, keys, numkeys.ravel())
                out = torch.full((len(want),), float("nan"), device="cuda")
                r = insts[i].infer(keys, numkeys, gpu_out=out)
                ok &= r.error_code is None and np.array_equal(out.cpu().numpy(), want)
            results[i] = ok

        threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
        [t.start() for t in threads]
        [t.join(120) for t in threads]
        assert results == {0: True, 1: True, 2: True, 3: True}
        for x in insts:
            x.close()
        ma.close()
        mb.close()


def test_tier_only_model_through_execute(tmp_path, cuda_device):
    """BASELINE configs[3] through the boundary, scaled down: a table given as "synthetic_device:rows=…" lives in the
    NVLink tier only (generated on the device, no host copy, no local cache) and is served by the same
    TRITONBACKEND_ModelInstanceExecute call as any other model — here with the one GPU of the test box as a world of 1."""
    import torch
    rows, dim, seed = 200_000, 128, 0xB2000031
    entry = model_entry("dlrm", [f"synthetic_device:rows={rows},seed={seed}"], [dim], [26], gpucache=True, max_batch=256,
                        hit_rate_threshold=1.0, gpucacheper=0.0, enable_pagelock=True, hpsx_peer_tier=True,
                        embedding_cache_type="static", workers=2)
    ps = ps_json(str(tmp_path / "ps.json"), [entry])
    ref = O.NumpyTable(dim, 0.0)
    ref.fill_procedural(rows, seed)
    rng = np.random.default_rng(31)
    with FT.Backend(ps) as be:
        m = be.model("dlrm", FT.model_config("dlrm", gpus=[0], count=2, max_batch_size=256))
        insts = [m.instance(name=f"dlrm_0_{j}", kind=FT.KIND_GPU, device=0) for j in range(2)]
        for samples in (256, 1, 77):
            for inst in insts:
                n = samples * 26
                keys = rng.integers(-3, rows + 50, size=n)  # a few keys no shard holds -> default vector
                out = torch.full((n * dim,), float("nan"), device="cuda")
                r = inst.infer(keys, np.array([[n]], dtype=np.int32), gpu_out=out)
                assert r.error_code is None, r.error_message
                assert r.params["NumSample"] == samples
                assert np.array_equal(out.cpu().numpy(), ref.lookup(keys).ravel())
        # host output buffer (Triton may hand back CPU memory, hps.cc:682-690)
        n = 64 * 26
        keys = rng.integers(0, rows, size=n)
        h_out = torch.full((n * dim,), float("nan"))
        r = insts[0].infer(keys, np.array([[n]], dtype=np.int32), cpu_out=h_out)
        assert r.error_code is None, r.error_message
        assert np.array_equal(h_out.numpy(), ref.lookup(keys).ravel())
        for inst in insts:
            inst.close()
        m.close()


def test_peer_tier_key_with_one_deployed_device_serves_from_the_host(tmp_path, cuda_device):
    """"hpsx_peer_tier": true needs >= 2 deployed devices to have anything to shard over; with one device the model is
    served as usual (misses from the page-locked host table) — the key must not break a single-GPU deployment."""
    import torch
    dirs, tables = write_tables(str(tmp_path), [(20000, 16)])
    entry = model_entry("m", dirs, [16], [26], gpucache=True, max_batch=256, hit_rate_threshold=1.0, gpucacheper=0.1,
                        enable_pagelock=True, hpsx_peer_tier=True)
    ps = ps_json(str(tmp_path / "ps.json"), [entry])
    ref = O.NumpyTable(16, 0.0)
    ref.insert(*tables[0])
    rng = np.random.default_rng(4)
    with FT.Backend(ps) as be:
        m = be.model("m", FT.model_config("m", gpus=[0], max_batch_size=256))
        inst = m.instance(kind=FT.KIND_GPU, device=0)
        n = 256 * 26
        keys = rng.choice(tables[0][0], size=n)
        out = torch.full((n * 16,), float("nan"), device="cuda")
        r = inst.infer(keys, np.array([[n]], dtype=np.int32), gpu_out=out)
        assert r.error_code is None, r.error_message
        assert np.array_equal(out.cpu().numpy(), ref.lookup(keys).ravel())
        inst.close()
        m.close()
