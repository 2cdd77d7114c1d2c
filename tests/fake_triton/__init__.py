"""Python face of tests/fake_triton/fake_triton.cpp — TEST INFRASTRUCTURE.

Plays the Triton server for libtriton_hps.so: loads the backend with ``--backend-config=hps,ps=<ps.json>``
semantics, loads models from a ``config.pbtxt``-as-JSON dict, creates instances and sends
``KEYS``/``NUMKEYS`` requests through ``TRITONBACKEND_ModelInstanceExecute``.
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_HERE))
HARNESS_SO = os.path.join(_HERE, "libfake_triton.so")
BACKEND_SO = os.path.join(ROOT, "hugectr_backend_b200", "lib", "libtriton_hps.so")

# Triton enum values (include/triton_compat.h)
TYPE_INT32, TYPE_INT64, TYPE_FP32 = 8, 9, 11
MEM_CPU, MEM_CPU_PINNED, MEM_GPU = 0, 1, 2
KIND_CPU, KIND_GPU = 1, 2
ERR = {"UNKNOWN": 0, "INTERNAL": 1, "NOT_FOUND": 2, "INVALID_ARG": 3, "UNAVAILABLE": 4, "UNSUPPORTED": 5,
       "ALREADY_EXISTS": 6}
RESPONSE_COMPLETE_FINAL = 1

_vp, _cp, _int = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int
_u32, _u64, _i64 = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64
_vpp = ctypes.POINTER(ctypes.c_void_p)

_SIGS = {
    "ft_last_error": (_cp, []),
    "ft_last_error_code": (_int, []),
    "ft_live_errors": (ctypes.c_long, []),
    "ft_live_messages": (ctypes.c_long, []),
    "ft_set_api_version": (None, [_u32, _u32]),
    "ft_log_count": (ctypes.c_long, [_int]),
    "ft_last_log": (_cp, [_int]),
    "ft_backend_load": (_int, [_cp, _cp, _cp, _cp, _vpp]),
    "ft_backend_unload": (_int, [_vp]),
    "ft_backend_has_state": (_int, [_vp]),
    "ft_model_load": (_int, [_vp, _cp, _u64, _cp, _cp, _vpp]),
    "ft_model_unload": (_int, [_vp]),
    "ft_instance_create": (_int, [_vp, _cp, _int, _int, _vpp]),
    "ft_instance_destroy": (_int, [_vp]),
    "ft_instance_stats": (None, [_vp, ctypes.POINTER(_u64)]),
    "ft_request_new": (_vp, [_cp, _u64]),
    "ft_request_delete": (None, [_vp]),
    "ft_request_add_input": (_int, [_vp, _cp, _int, ctypes.POINTER(_i64), _u32, _u64]),
    "ft_request_input_append_buffer": (None, [_vp, _int, _vp, _u64, _int, _i64]),
    "ft_request_add_requested_output": (None, [_vp, _cp]),
    "ft_request_set_gpu_output": (None, [_vp, _vp, _u64, _i64]),
    "ft_request_set_cpu_output": (None, [_vp, _vp, _u64, _int]),
    "ft_request_force_output_memory": (None, [_vp, _int]),
    "ft_request_fail_output_buffer": (None, [_vp, _int]),
    "ft_execute": (_int, [_vp, _vpp, _u32]),
    "ft_execute_repeat": (_int, [_vp, _vpp, _u32, _u32, ctypes.POINTER(_u64)]),
    "ft_execute_sequence": (_int, [_vp, _vpp, _u32, ctypes.POINTER(_u64)]),
    "ft_execute_sequences_parallel": (_int, [_vpp, _u32, _vpp, ctypes.POINTER(_u32), ctypes.POINTER(_u64)]),
    "ft_request_released": (_int, [_vp]),
    "ft_request_response_count": (_int, [_vp]),
    "ft_response_sent": (_int, [_vp]),
    "ft_response_flags": (_u32, [_vp]),
    "ft_response_error_code": (_int, [_vp]),
    "ft_response_error_message": (_cp, [_vp]),
    "ft_response_output_count": (_int, [_vp]),
    "ft_response_output": (_int, [_vp, _int, ctypes.POINTER(_cp), ctypes.POINTER(_int),
                                  ctypes.POINTER(ctypes.POINTER(_i64)), ctypes.POINTER(_u32), _vpp,
                                  ctypes.POINTER(_u64), ctypes.POINTER(_int), ctypes.POINTER(_i64)]),
    "ft_response_int_param": (_int, [_vp, _cp, ctypes.POINTER(_i64)]),
}

_lib = None


def build() -> None:
    subprocess.check_call(["make", "-C", ROOT, "tests/fake_triton/libfake_triton.so",
                           "hugectr_backend_b200/lib/libtriton_hps.so"], stdout=subprocess.DEVNULL)


def lib() -> ctypes.CDLL:
    """The harness, loaded RTLD_GLOBAL so that the backend's TRITONSERVER_*/TRITONBACKEND_* imports bind to it
    (Triton itself exports them from the server executable)."""
    global _lib
    if _lib is None:
        if not (os.path.exists(HARNESS_SO) and os.path.exists(BACKEND_SO)):
            build()
        L = ctypes.CDLL(HARNESS_SO, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class TritonError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"triton error {code}: {message}")
        self.code = code - 100 if code >= 100 else None  # TRITONSERVER_Error_Code, None for harness failures
        self.message = message


def _check(rc: int) -> None:
    if rc != 0:
        raise TritonError(rc, lib().ft_last_error().decode())


def model_config(name: str, *, max_batch_size: int = 0, gpus: Sequence[int] = (0,), count: int = 1,
                 kind: str = "KIND_GPU", output: str = "OUTPUT0", parameters: Optional[Dict[str, str]] = None,
                 two_dims: bool = False) -> dict:
    """config.pbtxt of an hps model as the JSON Triton hands to the backend
    (hps_backend/samples/Hierarchical_Parameter_Server_Deployment.ipynb:215-250;
    hps-triton-ensemble/02_model_inference_hps_tf_ensemble.ipynb:152-186)."""
    dims = [-1, -1] if two_dims else [-1]
    cfg = {
        "name": name, "backend": "hps", "max_batch_size": max_batch_size,
        "input": [{"name": "KEYS", "data_type": "TYPE_INT64", "dims": dims},
                  {"name": "NUMKEYS", "data_type": "TYPE_INT32", "dims": dims}],
        "output": [{"name": output, "data_type": "TYPE_FP32", "dims": [-1]}],
        "instance_group": [{"name": f"{name}_0", "count": count, "kind": kind, "gpus": list(gpus)}],
    }
    if kind != "KIND_GPU":
        cfg["instance_group"][0].pop("gpus")
    if parameters:
        cfg["parameters"] = {k: {"string_value": v} for k, v in parameters.items()}
    return cfg


@dataclass
class Response:
    error_code: Optional[int]  # TRITONSERVER_Error_Code of an error response, None on success
    error_message: str
    sent: int
    flags: int
    released: int
    output_name: Optional[str] = None
    shape: Optional[List[int]] = None
    memory_type: Optional[int] = None
    memory_type_id: Optional[int] = None
    data: Optional[np.ndarray] = None  # host copy of a CPU output buffer (None for GPU buffers)
    device_ptr: Optional[int] = None
    byte_size: int = 0
    params: Dict[str, int] = field(default_factory=dict)


class Instance:
    def __init__(self, model: "Model", name: str, kind: int, device: int):
        self._L = lib()
        self.model = model
        h = ctypes.c_void_p()
        _check(self._L.ft_instance_create(model._h, name.encode(), kind, device, ctypes.byref(h)))
        self._h = h
        self.device = device
        self._caller_cpu_out = False

    def close(self) -> None:
        if self._h:
            _check(self._L.ft_instance_destroy(self._h))
            self._h = None

    def stats(self) -> dict:
        a = (_u64 * 8)()
        self._L.ft_instance_stats(self._h, a)
        keys = ["ok_requests", "failed_requests", "batch_reports", "last_batch_size", "exec_start", "compute_start",
                "compute_end", "exec_end"]
        return dict(zip(keys, [int(x) for x in a]))

    def _make_request(self, keys, numkeys, *, numkeys_shape=None, keys_shape=None, gpu_out=None, out_device=0,
                      key_buffers: int = 1, keys_device_ptr: Optional[int] = None, numkeys_device_ptr: Optional[int] = None,
                      cpu_out=None, cpu_out_pinned: bool = False,
                      requested_output: Optional[str] = "OUTPUT0", force_output_memory: Optional[int] = None,
                      fail_output_buffer: bool = False, input_names=("KEYS", "NUMKEYS"), keys_dtype=TYPE_INT64,
                      numkeys_dtype=TYPE_INT32, request_id: str = "req", keep: Optional[list] = None):
        L = self._L
        keys = np.ascontiguousarray(keys, dtype=np.int64).ravel()
        numkeys = np.ascontiguousarray(numkeys, dtype=np.int32)
        if numkeys_shape is None:
            numkeys_shape = [1, numkeys.size]
        if keys_shape is None:
            keys_shape = [1, keys.size]
        keep.extend([keys, numkeys])
        r = L.ft_request_new(request_id.encode(), 0)
        ks = (_i64 * len(keys_shape))(*keys_shape)
        ki = L.ft_request_add_input(r, input_names[0].encode(), keys_dtype, ks, len(keys_shape), keys.nbytes)
        if keys_device_ptr is not None:
            L.ft_request_input_append_buffer(r, ki, keys_device_ptr, keys.nbytes, MEM_GPU, out_device)
        else:
            # optionally split KEYS into several buffers, as Triton does for batched/concatenated inputs
            bounds = np.linspace(0, keys.size, key_buffers + 1).astype(np.int64)
            for b in range(key_buffers):
                part = keys[bounds[b]:bounds[b + 1]]
                L.ft_request_input_append_buffer(r, ki, part.ctypes.data, part.nbytes, MEM_CPU, 0)
        ns = (_i64 * len(numkeys_shape))(*numkeys_shape)
        ni = L.ft_request_add_input(r, input_names[1].encode(), numkeys_dtype, ns, len(numkeys_shape), numkeys.nbytes)
        if numkeys_device_ptr is not None:
            L.ft_request_input_append_buffer(r, ni, numkeys_device_ptr, numkeys.nbytes, MEM_GPU, out_device)
        else:
            L.ft_request_input_append_buffer(r, ni, numkeys.ctypes.data, numkeys.nbytes, MEM_CPU, 0)
        if requested_output is not None:
            L.ft_request_add_requested_output(r, requested_output.encode())
        if gpu_out is not None:
            L.ft_request_set_gpu_output(r, gpu_out.data_ptr(), gpu_out.numel() * gpu_out.element_size(), out_device)
        if cpu_out is not None:
            # a caller-owned host buffer (torch CPU tensor, pinned or not) that the harness hands out for a CPU output
            L.ft_request_set_cpu_output(r, cpu_out.data_ptr(), cpu_out.numel() * cpu_out.element_size(),
                                        MEM_CPU_PINNED if cpu_out_pinned else MEM_CPU)
            keep.append(cpu_out)
        if force_output_memory is not None:
            L.ft_request_force_output_memory(r, force_output_memory)
        if fail_output_buffer:
            L.ft_request_fail_output_buffer(r, 1)
        return r

    def _collect(self, r) -> Response:
        L = self._L
        code = L.ft_response_error_code(r)
        resp = Response(error_code=None if code < 0 else code, error_message=L.ft_response_error_message(r).decode(),
                        sent=L.ft_response_sent(r), flags=L.ft_response_flags(r), released=L.ft_request_released(r))
        if L.ft_response_output_count(r) > 0:
            name, dtype, shape, dims = _cp(), _int(), ctypes.POINTER(_i64)(), _u32()
            buf, nbytes, mt, mt_id = ctypes.c_void_p(), _u64(), _int(), _i64()
            assert L.ft_response_output(r, 0, ctypes.byref(name), ctypes.byref(dtype), ctypes.byref(shape),
                                        ctypes.byref(dims), ctypes.byref(buf), ctypes.byref(nbytes), ctypes.byref(mt),
                                        ctypes.byref(mt_id)) == 0
            resp.output_name = name.value.decode()
            resp.shape = [int(shape[i]) for i in range(dims.value)]
            resp.memory_type, resp.memory_type_id = mt.value, mt_id.value
            resp.byte_size = nbytes.value
            assert dtype.value == TYPE_FP32
            if mt.value == MEM_GPU:
                resp.device_ptr = buf.value
            elif self._caller_cpu_out:
                pass  # the rows are in the caller's own buffer: no copy
            elif nbytes.value:
                resp.data = np.ctypeslib.as_array(ctypes.cast(buf, ctypes.POINTER(ctypes.c_float)),
                                                  shape=(nbytes.value // 4,)).copy()
            else:
                resp.data = np.empty(0, dtype=np.float32)
        for p in ("NumSample", "DeviceID", "CacheHits", "CacheMisses"):
            v = _i64()
            if L.ft_response_int_param(r, p.encode(), ctypes.byref(v)):
                resp.params[p] = int(v.value)
        return resp

    def infer_many(self, requests: Sequence[dict]) -> List[Response]:
        """One TRITONBACKEND_ModelInstanceExecute call carrying len(requests) requests."""
        keep: list = []
        self._caller_cpu_out = any(kw.get("cpu_out") is not None for kw in requests)
        handles = [self._make_request(keep=keep, **kw) for kw in requests]
        arr = (ctypes.c_void_p * len(handles))(*handles)
        try:
            _check(self._L.ft_execute(self._h, arr, len(handles)))
            return [self._collect(h) for h in handles]
        finally:
            for h in handles:
                self._L.ft_request_delete(h)

    def prepare(self, requests: Sequence[dict]) -> "Prepared":
        """Requests built once and executed many times (benchmarks): see Prepared.run."""
        return Prepared(self, requests)

    def infer(self, keys, numkeys, **kw) -> Response:
        return self.infer_many([dict(keys=keys, numkeys=numkeys, **kw)])[0]


class Prepared:
    """A fixed set of requests for one instance; run(repeat) executes them `repeat` times, ONE Execute call per pass
    carrying all of them, and returns the seconds spent inside the backend (timed in C, no Python in the loop)."""

    def __init__(self, inst: Instance, requests: Sequence[dict]):
        self.inst = inst
        self._keep: list = []
        inst._caller_cpu_out = any(kw.get("cpu_out") is not None for kw in requests)
        self._handles = [inst._make_request(keep=self._keep, **kw) for kw in requests]
        self._arr = (ctypes.c_void_p * len(self._handles))(*self._handles)

    def run(self, repeat: int = 1) -> float:
        ns = _u64()
        _check(self.inst._L.ft_execute_repeat(self.inst._h, self._arr, len(self._handles), repeat, ctypes.byref(ns)))
        return ns.value / 1e9

    def run_sequence(self) -> float:
        """One Execute call PER request, in order (a stream of distinct requests); seconds, timed in C."""
        ns = _u64()
        _check(self.inst._L.ft_execute_sequence(self.inst._h, self._arr, len(self._handles), ctypes.byref(ns)))
        return ns.value / 1e9

    def responses(self) -> List[Response]:
        return [self.inst._collect(h) for h in self._handles]

    def close(self) -> None:
        for h in self._handles:
            self.inst._L.ft_request_delete(h)
        self._handles = []

    def __del__(self):
        self.close()


def run_sequences_parallel(prepared: Sequence[Prepared], lo: int = 0, hi: Optional[int] = None) -> float:
    """Requests [lo, hi) of every Prepared as one stream per instance, all instances at once, one C++ thread each (a
    server process driving several GPUs; no Python in the timed region).  Returns the wall seconds of the whole pass."""
    L = prepared[0].inst._L
    k = len(prepared)
    insts = (ctypes.c_void_p * k)(*[p.inst._h for p in prepared])
    handles, counts = [], []
    for p in prepared:
        part = p._handles[lo:hi]
        handles += part
        counts.append(len(part))
    arr = (ctypes.c_void_p * len(handles))(*handles)
    cnt = (_u32 * k)(*counts)
    ns = _u64()
    _check(L.ft_execute_sequences_parallel(insts, k, arr, cnt, ctypes.byref(ns)))
    return ns.value / 1e9


class Model:
    def __init__(self, backend: "Backend", name: str, config: dict, version: int = 1, repo: str = "/models"):
        self._L = lib()
        self.backend = backend
        self.name = name
        h = ctypes.c_void_p()
        _check(self._L.ft_model_load(backend._h, name.encode(), version, json.dumps(config).encode(),
                                     f"{repo}/{name}".encode(), ctypes.byref(h)))
        self._h = h

    def instance(self, name: Optional[str] = None, kind: int = KIND_GPU, device: int = 0) -> Instance:
        return Instance(self, name or f"{self.name}_0", kind, device)

    def close(self) -> None:
        if self._h:
            _check(self._L.ft_model_unload(self._h))
            self._h = None


class Backend:
    """`tritonserver --backend-config=hps,ps=<ps_json>` (hps_backend/README.md:105-109)."""

    def __init__(self, ps_json: Optional[str], api_version=(1, 10), extra_cmdline: Optional[dict] = None,
                 so_path: str = BACKEND_SO):
        self._L = lib()
        self._L.ft_set_api_version(*api_version)
        cmdline = {"auto-complete-config": "false", "backend-directory": "/opt/tritonserver/backends",
                   "min-compute-capability": "6.000000", "default-max-batch-size": "4"}
        if ps_json is not None:
            cmdline["ps"] = ps_json
        cmdline.update(extra_cmdline or {})
        h = ctypes.c_void_p()
        _check(self._L.ft_backend_load(so_path.encode(), b"hps", json.dumps({"cmdline": cmdline}).encode(),
                                       os.path.dirname(so_path).encode(), ctypes.byref(h)))
        self._h = h

    def model(self, name: str, config: dict, version: int = 1) -> Model:
        return Model(self, name, config, version)

    def close(self) -> None:
        if self._h:
            _check(self._L.ft_backend_unload(self._h))
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def live_errors() -> int:
    return int(lib().ft_live_errors())


def live_messages() -> int:
    return int(lib().ft_live_messages())


def log_count(level: int) -> int:
    return int(lib().ft_log_count(level))


def last_log(level: int) -> str:
    return lib().ft_last_log(level).decode()
