"""Host-side mirror of the interfaces the reference glue uses for the lookup path.

``HPS``            ~ ``HugeCTR::HierParameterServerBase``  (reference: hps_backend/src/backend.cpp:68-71)
``LookupSession``  ~ ``HugeCTR::LookupSessionBase``        (hps_backend/src/model_instance_state.cpp:170-195)

Every method is one call into ``libhpsx.so``; arrays are passed by address (numpy for host memory,
torch CUDA tensors or raw integer addresses for device memory).  Nothing is computed in Python.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N


def _addr(x) -> int:
    """Address of a numpy array, a torch tensor or a raw pointer."""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr()
    raise TypeError(f"cannot take the address of {type(x)!r}")


PROBE_VARIANTS = {"ldg": 0, "tma": 1, "v8": 4}


@dataclass
class ModelParams:
    """~ ``HugeCTR::InferenceParams`` as filled from ps.json (hps_backend/src/backend.cpp:318-523)."""
    model_name: str
    max_batch_size: int
    embedding_vecsize_per_table: Sequence[int]
    maxnum_catfeature_query_per_table_per_sample: Sequence[int]
    default_value_for_each_table: Optional[Sequence[float]] = None
    sparse_files: Optional[Sequence[str]] = None
    use_gpu_embedding_cache: bool = True
    hit_rate_threshold: float = 1.0
    cache_size_percentage: float = 1.0
    number_of_worker_buffers_in_pool: int = 1
    deployed_devices: Sequence[int] = field(default_factory=lambda: [0])
    embedding_cache_type: str = "dynamic"
    cache_load_factor: float = 0.0
    enable_pagelock: bool = False
    # engine extensions (include/hpsx.h hpsx_model_params)
    split_lock: bool = True
    request_chunks: int = 0
    pull_grid_ctas: int = 0
    probe_variant: Optional[str] = None  # "v8" (default), "ldg", "tma"
    peer_tier: bool = False  # NVLink tier over the deployed devices of this process (needs enable_pagelock)


@dataclass
class SessionStats:
    lookups: int
    keys: int
    hits: int
    misses: int
    inserted: int
    default_filled: int
    h2d_bytes: int
    d2h_bytes: int
    kernel_launches: int
    probe_kernel_ms: float
    probe_kernel_launches: int
    probe_kernel_keys: int
    insert_kernel_ms: float
    host_gather_ms: float
    pull_kernel_ms: float = 0.0
    tier_bytes: int = 0


class HPS:
    """The parameter server: host database + per-(model, device) HBM embedding caches."""

    def __init__(self, ps_json: Optional[str] = None, num_partitions: int = 0, num_threads: int = 0,
                 allocation_rate: int = 0, pull_window_mb: int = 0):
        self._L = N.lib()
        h = ctypes.c_void_p()
        if ps_json is not None:
            N.check(self._L.hpsx_ps_create_from_json(ps_json.encode(), ctypes.byref(h)))
        else:
            vp = N.VolatileParamsC(num_partitions, allocation_rate, 1.0, num_threads, pull_window_mb << 20)
            N.check(self._L.hpsx_ps_create(ctypes.byref(vp), ctypes.byref(h)))
        self._h = h
        self._dims = {}

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.hpsx_ps_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- models -------------------------------------------------------------------------------
    def model_names(self) -> List[str]:
        n = ctypes.c_size_t()
        N.check(self._L.hpsx_ps_num_models(self._h, ctypes.byref(n)))
        out = []
        for i in range(n.value):
            s = ctypes.c_char_p()
            N.check(self._L.hpsx_ps_model_name(self._h, i, ctypes.byref(s)))
            out.append(s.value.decode())
        return out

    def has_model(self, name: str) -> bool:
        return bool(self._L.hpsx_ps_has_model(self._h, name.encode()))

    def add_model(self, p: ModelParams) -> None:
        T = len(p.embedding_vecsize_per_table)
        defaults = list(p.default_value_for_each_table) if p.default_value_for_each_table is not None else [0.0] * T
        c = N.ModelParamsC()
        c.model_name = p.model_name.encode()
        c.max_batch_size = p.max_batch_size
        c.num_tables = T
        keep = []
        if p.sparse_files:
            arr = (ctypes.c_char_p * T)(*[s.encode() if s else None for s in p.sparse_files])
            keep.append(arr)
            c.sparse_files = arr
        vec = (ctypes.c_size_t * T)(*[int(v) for v in p.embedding_vecsize_per_table])
        mq = (ctypes.c_size_t * T)(*[int(v) for v in p.maxnum_catfeature_query_per_table_per_sample])
        dv = (ctypes.c_float * T)(*[float(v) for v in defaults])
        dev = (ctypes.c_int * len(p.deployed_devices))(*[int(d) for d in p.deployed_devices])
        c.embedding_vecsize_per_table = vec
        c.maxnum_catfeature_query_per_table_per_sample = mq
        c.default_value_for_each_table = dv
        c.use_gpu_embedding_cache = 1 if p.use_gpu_embedding_cache else 0
        c.hit_rate_threshold = p.hit_rate_threshold
        c.cache_size_percentage = p.cache_size_percentage
        c.number_of_worker_buffers_in_pool = p.number_of_worker_buffers_in_pool
        c.deployed_devices = dev
        c.num_deployed_devices = len(p.deployed_devices)
        c.embedding_cache_type = 1 if p.embedding_cache_type.lower() == "static" else 0
        c.cache_load_factor = p.cache_load_factor
        c.enable_pagelock = 1 if p.enable_pagelock else 0
        c.split_lock = 0 if p.split_lock else -1
        c.request_chunks = int(p.request_chunks)
        c.pull_grid_ctas = int(p.pull_grid_ctas)
        c.probe_variant_set = 0 if p.probe_variant is None else 1
        c.probe_variant = PROBE_VARIANTS[p.probe_variant] if p.probe_variant is not None else 0
        c.peer_tier = 1 if p.peer_tier else 0
        N.check(self._L.hpsx_ps_add_model(self._h, ctypes.byref(c)))
        self._dims[p.model_name] = [int(v) for v in p.embedding_vecsize_per_table]

    def load_table(self, model: str, table: int, keys: np.ndarray, vectors: np.ndarray) -> None:
        keys = np.ascontiguousarray(keys, dtype=np.int64).ravel()
        vectors = np.ascontiguousarray(vectors, dtype=np.float32)
        N.check(self._L.hpsx_ps_load_table(self._h, model.encode(), table, _addr(keys), _addr(vectors), len(keys)))

    def load_table_procedural(self, model: str, table: int, rows: int, seed: int) -> None:
        N.check(self._L.hpsx_ps_load_table_procedural(self._h, model.encode(), table, rows, seed & ((1 << 64) - 1)))

    def load_table_procedural_shard(self, model: str, table: int, rows: int, seed: int, shard: int, num_shards: int) -> None:
        N.check(self._L.hpsx_ps_load_table_procedural_shard(self._h, model.encode(), table, rows,
                                                            seed & ((1 << 64) - 1), shard, num_shards))

    def table_rows(self, model: str, table: int) -> int:
        n = ctypes.c_size_t()
        N.check(self._L.hpsx_ps_table_rows(self._h, model.encode(), table, ctypes.byref(n)))
        return n.value

    def request_capacity(self, model: str, table: int) -> int:
        """Keys of `table` one request of the model may hold: max_batch_size x maxnum_catfeature_query_per_table_per_sample."""
        p = N.ModelParamsC()
        N.check(self._L.hpsx_ps_get_model_params(self._h, model.encode(), ctypes.byref(p)))
        return int(p.max_batch_size) * int(p.maxnum_catfeature_query_per_table_per_sample[table])

    def lookup(self, keys: np.ndarray, model: str, table: int, dim: Optional[int] = None) -> np.ndarray:
        """CPU parameter-server lookup (gpucache = false path).  ~ ``HPS.lookup(key, model_name, table_id)``."""
        keys = np.ascontiguousarray(keys, dtype=np.int64).ravel()
        if dim is None:
            dim = self._dims[model][table]
        out = np.empty((len(keys), dim), dtype=np.float32)
        N.check(self._L.hpsx_ps_lookup(self._h, model.encode(), table, _addr(keys), len(keys), _addr(out)))
        return out

    # -- embedding caches -----------------------------------------------------------------------
    def create_embedding_cache(self, model: str) -> None:
        N.check(self._L.hpsx_ps_create_embedding_cache_per_model(self._h, model.encode()))

    def update_database(self, model: str) -> None:
        """~ ``update_database_per_model``: re-read the model's sparse files into the host database."""
        N.check(self._L.hpsx_ps_update_database_per_model(self._h, model.encode()))

    def refresh_embedding_cache(self, model: str, device: int = 0) -> int:
        """~ ``refresh_embedding_cache(model, device)``: rewrite every cached row from the host database."""
        n = ctypes.c_size_t()
        N.check(self._L.hpsx_ps_refresh_embedding_cache(self._h, model.encode(), device, ctypes.byref(n)))
        return n.value

    def destroy_embedding_cache(self, model: str) -> None:
        N.check(self._L.hpsx_ps_destroy_embedding_cache_per_model(self._h, model.encode()))

    def _cache(self, model: str, device: int) -> ctypes.c_void_p:
        c = ctypes.c_void_p()
        N.check(self._L.hpsx_ps_get_embedding_cache(self._h, model.encode(), device, ctypes.byref(c)))
        return c

    def cache_capacity(self, model: str, device: int, table: int) -> int:
        n = ctypes.c_size_t()
        N.check(self._L.hpsx_cache_capacity(self._cache(model, device), table, ctypes.byref(n)))
        return n.value

    def cache_resident(self, model: str, device: int, table: int) -> int:
        n = ctypes.c_size_t()
        N.check(self._L.hpsx_cache_resident(self._cache(model, device), table, ctypes.byref(n)))
        return n.value

    def cache_keys(self, model: str, device: int, table: int) -> np.ndarray:
        cap = self.cache_capacity(model, device, table)
        out = np.empty(cap, dtype=np.int64)
        n = ctypes.c_size_t()
        N.check(self._L.hpsx_cache_dump_keys(self._cache(model, device), table, _addr(out), cap, ctypes.byref(n)))
        return out[: n.value].copy()

    # ---- NVLink tier (include/hpsx.h hpsx_cache_peer_tier_*) ----
    def peer_tier_connect_local(self, model: str) -> None:
        """Every cache of `model` in this process becomes one rank of a tier."""
        N.check(self._L.hpsx_ps_peer_tier_connect_local(self._h, model.encode()))

    def peer_tier_build(self, model: str, device: int, rank: int, world: int) -> None:
        N.check(self._L.hpsx_cache_peer_tier_build(self._cache(model, device), rank, world))

    def peer_tier_export(self, model: str, device: int, table: int):
        """-> (64-byte CUDA IPC handle, rows, capacity) of this rank's shard of `table`."""
        h = (ctypes.c_ubyte * 64)()
        rows, cap = ctypes.c_uint64(), ctypes.c_uint64()
        N.check(self._L.hpsx_cache_peer_tier_export(self._cache(model, device), table, h, ctypes.byref(rows), ctypes.byref(cap)))
        return bytes(h), int(rows.value), int(cap.value)

    def peer_tier_attach_ipc(self, model: str, device: int, table: int, peer: int, handle: bytes, rows: int, cap: int) -> None:
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        N.check(self._L.hpsx_cache_peer_tier_attach_ipc(self._cache(model, device), table, peer, buf, rows, cap))

    def peer_tier_attach_local(self, model: str, device: int, peer: int, peer_hps: "HPS", peer_model: str, peer_device: int) -> None:
        N.check(self._L.hpsx_cache_peer_tier_attach_local(self._cache(model, device), peer, peer_hps._cache(peer_model, peer_device)))

    def peer_tier_commit(self, model: str, device: int) -> None:
        N.check(self._L.hpsx_cache_peer_tier_commit(self._cache(model, device)))

    def peer_tier_detach(self, model: str, device: int) -> None:
        N.check(self._L.hpsx_cache_peer_tier_detach(self._cache(model, device)))

    def peer_tier_info(self, model: str, device: int) -> dict:
        info = N.PeerTierInfoC()
        N.check(self._L.hpsx_cache_peer_tier_info(self._cache(model, device), ctypes.byref(info)))
        return {f[0]: int(getattr(info, f[0])) for f in N.PeerTierInfoC._fields_}

    def peer_tier_connect_distributed(self, model: str, device: int, rank: int, world: int, num_tables: int, all_gather_object) -> dict:
        """One process per GPU: build this rank's shards, exchange the CUDA IPC handles with
        `all_gather_object(obj) -> list of world objs` (e.g. a torch.distributed wrapper), map the peers' shards and
        point the direct-pull index at them.  Collective over the ranks; call peer_tier_detach on every rank (and
        synchronise) before any rank destroys its cache."""
        self.peer_tier_build(model, device, rank, world)
        mine = [self.peer_tier_export(model, device, t) for t in range(num_tables)]
        everyone = all_gather_object(mine)
        for p, shards in enumerate(everyone):
            if p == rank:
                continue
            for t, (handle, rows, cap) in enumerate(shards):
                self.peer_tier_attach_ipc(model, device, t, p, handle, rows, cap)
        self.peer_tier_commit(model, device)
        return self.peer_tier_info(model, device)

    def drain_async(self, model: str, device: int) -> None:
        N.check(self._L.hpsx_cache_drain_async(self._cache(model, device)))

    def session(self, model: str, device: int = 0) -> "LookupSession":
        return LookupSession(self, model, device)


class LookupSession:
    """One lookup workspace + stream (one per Triton model instance)."""

    def __init__(self, hps: HPS, model: str, device: int):
        self._L = N.lib()
        self._hps = hps  # keep the server alive
        h = ctypes.c_void_p()
        N.check(self._L.hpsx_session_create(hps._h, model.encode(), device, ctypes.byref(h)))
        self._h = h
        d = ctypes.c_int()
        N.check(self._L.hpsx_session_device(h, ctypes.byref(d)))
        self.device = d.value

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.hpsx_session_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @property
    def stream(self) -> int:
        p = ctypes.c_void_p()
        N.check(self._L.hpsx_session_stream(self._h, ctypes.byref(p)))
        return p.value or 0

    def _arrays(self, keys_per_table, out_per_table, counts):
        T = len(counts)
        k = (ctypes.c_void_p * T)(*[_addr(x) for x in keys_per_table])
        o = (ctypes.c_void_p * T)(*[_addr(x) for x in out_per_table])
        n = (ctypes.c_size_t * T)(*[int(c) for c in counts])
        return k, o, n, T

    def lookup(self, h_keys_per_table, vectors_per_table, num_keys_per_table) -> None:
        """~ ``LookupSessionBase::lookup(h_keys_per_table, d_vectors_per_table, num_keys_per_table)``."""
        k, o, n, T = self._arrays(h_keys_per_table, vectors_per_table, num_keys_per_table)
        N.check(self._L.hpsx_session_lookup(self._h, k, o, n, T))

    def lookup_device_keys(self, d_keys_per_table, d_vectors_per_table, num_keys_per_table) -> None:
        k, o, n, T = self._arrays(d_keys_per_table, d_vectors_per_table, num_keys_per_table)
        N.check(self._L.hpsx_session_lookup_device_keys(self._h, k, o, n, T))

    def bind(self, keys_per_table, vectors_per_table, num_keys_per_table, device_keys: bool = False):
        """A lookup with its argument arrays built once: returns a callable that only makes the C call (benchmarks that
        drive several GPUs from Python threads keep the interpreter's share of a sub-millisecond step small)."""
        K, O_, Nn, T = self._arrays(keys_per_table, vectors_per_table, num_keys_per_table)
        fn = self._L.hpsx_session_lookup_device_keys if device_keys else self._L.hpsx_session_lookup
        h, keep = self._h, (keys_per_table, vectors_per_table)

        def call():
            rc = fn(h, K, O_, Nn, T)
            if rc != 0:
                N.check(rc)
            return keep is None

        return call

    def lookup_ex(self, keys_per_table, vectors_per_table, num_keys_per_table, key_memory: str = "host",
                  vector_memory: str = "device") -> None:
        """General form (what the Triton shell calls): keys and vectors each in "host" or "device" memory."""
        K, O_, Nn, T = self._arrays(keys_per_table, vectors_per_table, num_keys_per_table)
        N.check(self._L.hpsx_session_lookup_ex(self._h, K, 1 if key_memory == "device" else 0, O_,
                                               1 if vector_memory == "device" else 0, Nn, T))

    def lookup_batch(self, requests, device_keys: bool = False, device_vectors: bool = True) -> None:
        """``requests``: list of (keys_per_table, vectors_per_table, num_keys_per_table), served in one pass."""
        k, o, n = [], [], []
        for kp, op, cp in requests:
            k += [_addr(x) for x in kp]
            o += [_addr(x) for x in op]
            n += [int(c) for c in cp]
        K = (ctypes.c_void_p * len(k))(*k)
        O_ = (ctypes.c_void_p * len(o))(*o)
        Nn = (ctypes.c_size_t * len(n))(*n)
        N.check(self._L.hpsx_session_lookup_batch(self._h, len(requests), K, 1 if device_keys else 0, O_,
                                                  1 if device_vectors else 0, Nn))

    def lookup_bf16_mirror(self, table: int, keys, n: int, d_vectors, d_vectors_bf16, device_keys: bool = False) -> None:
        """Lookup of one table that also writes a bf16 copy of the vectors (feeds ``DenseMlp.forward_bf16``)."""
        N.check(self._L.hpsx_session_lookup_bf16_mirror(self._h, table, _addr(keys), 1 if device_keys else 0, n,
                                                        _addr(d_vectors), _addr(d_vectors_bf16)))

    def lookup_pooled(self, table: int, keys, num_bags: int, hotness: int, d_pooled, combiner: str = "sum",
                      device_keys: bool = False) -> None:
        comb = 1 if combiner == "mean" else 0
        fn = self._L.hpsx_session_lookup_pooled_device_keys if device_keys else self._L.hpsx_session_lookup_pooled
        N.check(fn(self._h, table, _addr(keys), num_bags, hotness, comb, _addr(d_pooled)))

    def stats(self) -> SessionStats:
        s = N.SessionStatsC()
        N.check(self._L.hpsx_session_get_stats(self._h, ctypes.byref(s)))
        return SessionStats(*[getattr(s, f[0]) for f in N.SessionStatsC._fields_])

    def reset_stats(self) -> None:
        N.check(self._L.hpsx_session_reset_stats(self._h))

    def set_insert_mode(self, mode: int) -> None:
        N.check(self._L.hpsx_session_set_insert_mode(self._h, mode))

    def set_debug(self, flags: int) -> None:
        N.check(self._L.hpsx_session_set_debug(self._h, flags))

    def set_probe_variant(self, variant: str) -> None:
        N.check(self._L.hpsx_session_set_probe_variant(self._h, PROBE_VARIANTS[variant]))


class _DeviceRows:
    """Zero-copy view of a device buffer owned by the engine (``__cuda_array_interface__``):
    ``torch.as_tensor(view, device="cuda")`` aliases it."""

    def __init__(self, ptr: int, rows: int, dim: int, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": (rows, dim), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


class ShardGroup:
    """~ ``hpsx_shard_group`` (include/hpsx.h): rank-local half of a model-parallel table.  The reference has no
    counterpart (replicas only, hps_backend/src/model_state.cpp:395-419)."""

    def __init__(self, session: LookupSession, table: int, rank: int, world: int, dim: int):
        self._L = N.lib()
        self._session = session
        self.rank, self.world, self.dim = int(rank), int(world), int(dim)
        h = ctypes.c_void_p()
        self.handle = np.zeros(64, dtype=np.uint8)
        N.check(self._L.hpsx_shard_group_create(session._h, table, rank, world, ctypes.byref(h), _addr(self.handle)))
        self._h = h

    def connect_ipc(self, all_handles: np.ndarray) -> None:
        """``all_handles``: uint8 [world, 64], row r = rank r's ``handle`` (other processes of the box)."""
        hs = np.ascontiguousarray(all_handles, dtype=np.uint8).reshape(self.world, 64)
        N.check(self._L.hpsx_shard_group_connect_ipc(self._h, _addr(hs)))

    def connect_local(self, groups: Sequence["ShardGroup"]) -> None:
        """All ranks live in this process (threads, one per GPU)."""
        arr = (ctypes.c_void_p * self.world)(*[g._h for g in groups])
        N.check(self._L.hpsx_shard_group_connect_local(self._h, arr))

    def lookup(self, d_keys, n: int):
        """Collective.  Returns a zero-copy [n, dim] view of this rank's output buffer (valid until the next lookup)."""
        out = ctypes.c_void_p()
        N.check(self._L.hpsx_shard_group_lookup(self._h, _addr(d_keys), n, ctypes.byref(out)))
        return _DeviceRows(out.value or 0, n, self.dim, self)

    @property
    def capacity(self) -> int:
        c = ctypes.c_size_t()
        N.check(self._L.hpsx_shard_group_capacity(self._h, ctypes.byref(c)))
        return c.value

    def lookup_ptr(self, d_keys, n: int) -> int:
        out = ctypes.c_void_p()
        N.check(self._L.hpsx_shard_group_lookup(self._h, _addr(d_keys), n, ctypes.byref(out)))
        return out.value or 0

    def stats(self) -> dict:
        s = N.ShardStatsC()
        N.check(self._L.hpsx_shard_group_get_stats(self._h, ctypes.byref(s)))
        return {"keys_sent_remote": s.keys_sent_remote, "keys_received": s.keys_received,
                "keys_received_remote": s.keys_received_remote, "misses": s.misses, "status": s.status,
                "sent": np.array(s.sent[:self.world], dtype=np.int64), "received": np.array(s.received[:self.world], dtype=np.int64)}

    def set_timeout_ms(self, ms: int) -> None:
        N.check(self._L.hpsx_shard_group_set_timeout_ms(self._h, int(ms)))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.hpsx_shard_group_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class DenseMlp:
    """~ the dense model of the reference's ensembles (hps-triton-ensemble/01_model_training.ipynb cells 7,11) on the
    tcgen05 tensor cores.  ``weights[l]``: float32 [out, in]; ``biases[l]``: float32 [out] or None."""

    def __init__(self, device: int, weights, biases=None, relu=None, precision: str = "bf16"):
        """precision "bf16" (default) or "tf32" (weights and activations stay fp32; ≤1e-3 of an fp32 model)."""
        self._L = N.lib()
        self.precision = precision
        L = len(weights)
        ws = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
        dims = [ws[0].shape[1]] + [w.shape[0] for w in ws]
        for l in range(L):
            if ws[l].shape[1] != dims[l]:
                raise ValueError(f"layer {l}: weight is {ws[l].shape}, expected [*, {dims[l]}]")
        bs = [None if (biases is None or biases[l] is None) else np.ascontiguousarray(biases[l], dtype=np.float32) for l in range(L)]
        self.dims = dims
        D = (ctypes.c_size_t * (L + 1))(*dims)
        W = (ctypes.c_void_p * L)(*[_addr(w) for w in ws])
        B = (ctypes.c_void_p * L)(*[_addr(b) for b in bs])
        R = (ctypes.c_int * L)(*[int(bool(r)) for r in (relu or [0] * L)])
        h = ctypes.c_void_p()
        N.check(self._L.hpsx_mlp_create_ex(device, L, D, W, B, R, {"bf16": 0, "tf32": 1}[precision], ctypes.byref(h)))
        self._h = h

    def forward(self, d_in, batch: int, d_out, stream: int = 0) -> None:
        N.check(self._L.hpsx_mlp_forward(self._h, _addr(d_in), batch, _addr(d_out), stream))

    def forward_bf16(self, d_in_bf16, batch: int, d_out, stream: int = 0) -> None:
        N.check(self._L.hpsx_mlp_forward_bf16(self._h, _addr(d_in_bf16), batch, _addr(d_out), stream))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.hpsx_mlp_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


# -- stand-alone device primitives ---------------------------------------------------------------
def unique(device: int, d_keys, n: int, d_unique, d_inverse, stream: int = 0) -> int:
    u = ctypes.c_size_t()
    N.check(N.lib().hpsx_unique(device, _addr(d_keys), n, _addr(d_unique), _addr(d_inverse), ctypes.byref(u), stream))
    return u.value


def owner(key: int, num_shards: int) -> int:
    return int(N.lib().hpsx_owner(int(key), int(num_shards)))


def owner_batch(keys: np.ndarray, num_shards: int) -> np.ndarray:
    keys = np.ascontiguousarray(keys, dtype=np.int64).ravel()
    out = np.empty(len(keys), dtype=np.uint32)
    N.check(N.lib().hpsx_owner_batch(_addr(keys), len(keys), num_shards, _addr(out)))
    return out


def route_keys(device: int, d_keys, n: int, num_shards: int, d_routed, d_perm, d_counts, stream: int = 0) -> np.ndarray:
    h = np.zeros(num_shards, dtype=np.uint32)
    N.check(N.lib().hpsx_route_keys(device, _addr(d_keys), n, num_shards, _addr(d_routed), _addr(d_perm),
                                    _addr(d_counts), _addr(h), stream))
    return h


def scatter_rows(device: int, d_rows, d_perm, n: int, dim: int, d_out, stream: int = 0) -> None:
    N.check(N.lib().hpsx_scatter_rows(device, _addr(d_rows), _addr(d_perm), n, dim, _addr(d_out), stream))


def gather_rows(device: int, d_table, d_idx, n: int, dim: int, d_out, stream: int = 0) -> None:
    """Measurement primitive: d_out[i] = d_table[d_idx[i]] (rows of 128 floats)."""
    N.check(N.lib().hpsx_gather_rows(device, _addr(d_table), _addr(d_idx), n, dim, _addr(d_out), stream))
