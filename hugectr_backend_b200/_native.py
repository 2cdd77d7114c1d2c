"""ctypes declarations of include/hpsx.h.  No fallback: a missing libhpsx.so is an error."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "lib", "libhpsx.so")


class HpsxError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"hpsx error {code}: {message}")
        self.code = code
        self.message = message


OK, ERR_INVALID_ARG, ERR_NOT_FOUND, ERR_UNSUPPORTED, ERR_CUDA, ERR_IO, ERR_INTERNAL = 0, -1, -2, -3, -4, -5, -6

c_size_p = ctypes.POINTER(ctypes.c_size_t)


class ModelParamsC(ctypes.Structure):
    _fields_ = [
        ("model_name", ctypes.c_char_p),
        ("max_batch_size", ctypes.c_size_t),
        ("num_tables", ctypes.c_size_t),
        ("sparse_files", ctypes.POINTER(ctypes.c_char_p)),
        ("table_names", ctypes.POINTER(ctypes.c_char_p)),
        ("embedding_vecsize_per_table", c_size_p),
        ("maxnum_catfeature_query_per_table_per_sample", c_size_p),
        ("default_value_for_each_table", ctypes.POINTER(ctypes.c_float)),
        ("use_gpu_embedding_cache", ctypes.c_int),
        ("hit_rate_threshold", ctypes.c_float),
        ("cache_size_percentage", ctypes.c_float),
        ("number_of_worker_buffers_in_pool", ctypes.c_size_t),
        ("deployed_devices", ctypes.POINTER(ctypes.c_int)),
        ("num_deployed_devices", ctypes.c_size_t),
        ("embedding_cache_type", ctypes.c_int),
        ("cache_load_factor", ctypes.c_float),
        ("enable_pagelock", ctypes.c_int),
        ("split_lock", ctypes.c_int),
        ("request_chunks", ctypes.c_int),
        ("pull_grid_ctas", ctypes.c_int),
        ("probe_variant", ctypes.c_int),
        ("probe_variant_set", ctypes.c_int),
        ("peer_tier", ctypes.c_int),
    ]


class VolatileParamsC(ctypes.Structure):
    _fields_ = [
        ("num_partitions", ctypes.c_size_t),
        ("allocation_rate", ctypes.c_size_t),
        ("initial_cache_rate", ctypes.c_double),
        ("num_threads", ctypes.c_size_t),
        ("pull_window_bytes", ctypes.c_size_t),
    ]


class SessionStatsC(ctypes.Structure):
    _fields_ = [
        ("lookups", ctypes.c_uint64),
        ("keys", ctypes.c_uint64),
        ("hits", ctypes.c_uint64),
        ("misses", ctypes.c_uint64),
        ("inserted", ctypes.c_uint64),
        ("default_filled", ctypes.c_uint64),
        ("h2d_bytes", ctypes.c_uint64),
        ("d2h_bytes", ctypes.c_uint64),
        ("kernel_launches", ctypes.c_uint64),
        ("probe_kernel_ms", ctypes.c_double),
        ("probe_kernel_launches", ctypes.c_uint64),
        ("probe_kernel_keys", ctypes.c_uint64),
        ("insert_kernel_ms", ctypes.c_double),
        ("host_gather_ms", ctypes.c_double),
        ("pull_kernel_ms", ctypes.c_double),
        ("tier_bytes", ctypes.c_uint64),
    ]


class PeerTierInfoC(ctypes.Structure):
    _fields_ = [
        ("rank", ctypes.c_uint32),
        ("world", ctypes.c_uint32),
        ("committed", ctypes.c_int),
        ("own_rows", ctypes.c_uint64),
        ("own_bytes", ctypes.c_uint64),
        ("index_entries_in_tier", ctypes.c_uint64),
    ]


class ShardStatsC(ctypes.Structure):
    _fields_ = [
        ("keys_sent_remote", ctypes.c_uint64),
        ("keys_received", ctypes.c_uint64),
        ("keys_received_remote", ctypes.c_uint64),
        ("misses", ctypes.c_uint64),
        ("status", ctypes.c_uint32),
        ("sent", ctypes.c_uint32 * 16),
        ("received", ctypes.c_uint32 * 16),
    ]


# every symbol include/hpsx.h declares: name -> (restype, argtypes)
_vp, _sz, _int, _cp = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p
_vpp = ctypes.POINTER(ctypes.c_void_p)
SYMBOLS = {
    "hpsx_abi_version": (_int, []),
    "hpsx_last_error": (_cp, []),
    "hpsx_device_count": (_int, []),
    "hpsx_ps_create_from_json": (_int, [_cp, _vpp]),
    "hpsx_ps_create": (_int, [ctypes.POINTER(VolatileParamsC), _vpp]),
    "hpsx_ps_destroy": (_int, [_vp]),
    "hpsx_ps_num_models": (_int, [_vp, c_size_p]),
    "hpsx_ps_model_name": (_int, [_vp, _sz, ctypes.POINTER(_cp)]),
    "hpsx_ps_has_model": (_int, [_vp, _cp]),
    "hpsx_ps_add_model": (_int, [_vp, ctypes.POINTER(ModelParamsC)]),
    "hpsx_ps_load_table": (_int, [_vp, _cp, _sz, _vp, _vp, _sz]),
    "hpsx_ps_load_table_procedural": (_int, [_vp, _cp, _sz, _sz, ctypes.c_uint64]),
    "hpsx_ps_load_table_procedural_shard": (_int, [_vp, _cp, _sz, _sz, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]),
    "hpsx_ps_table_rows": (_int, [_vp, _cp, _sz, c_size_p]),
    "hpsx_ps_get_model_params": (_int, [_vp, _cp, ctypes.POINTER(ModelParamsC)]),
    "hpsx_ps_sync_models_from_json": (_int, [_vp, _cp, c_size_p]),
    "hpsx_session_lookup_scatter": (_int, [_vp, _sz, _vp, _vp, _sz, _vp]),
    "hpsx_device_malloc": (_int, [_int, _sz, _vpp]),
    "hpsx_device_free": (_int, [_int, _vp]),
    "hpsx_ipc_export": (_int, [_int, _vp, _vp]),
    "hpsx_ipc_open": (_int, [_int, _vp, _vpp]),
    "hpsx_ipc_close": (_int, [_int, _vp]),
    "hpsx_shard_group_create": (_int, [_vp, _sz, ctypes.c_uint32, ctypes.c_uint32, _vpp, _vp]),
    "hpsx_shard_group_connect_ipc": (_int, [_vp, _vp]),
    "hpsx_shard_group_connect_local": (_int, [_vp, _vpp]),
    "hpsx_shard_group_lookup": (_int, [_vp, _vp, _sz, _vpp]),
    "hpsx_shard_group_get_stats": (_int, [_vp, ctypes.POINTER(ShardStatsC)]),
    "hpsx_shard_group_capacity": (_int, [_vp, c_size_p]),
    "hpsx_shard_group_set_timeout_ms": (_int, [_vp, ctypes.c_uint64]),
    "hpsx_shard_group_destroy": (_int, [_vp]),
    "hpsx_cache_peer_tier_build": (_int, [_vp, ctypes.c_uint32, ctypes.c_uint32]),
    "hpsx_cache_peer_tier_export": (_int, [_vp, _sz, _vp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]),
    "hpsx_cache_peer_tier_attach_ipc": (_int, [_vp, _sz, ctypes.c_uint32, _vp, ctypes.c_uint64, ctypes.c_uint64]),
    "hpsx_cache_peer_tier_attach_local": (_int, [_vp, ctypes.c_uint32, _vp]),
    "hpsx_cache_peer_tier_commit": (_int, [_vp]),
    "hpsx_cache_peer_tier_detach": (_int, [_vp]),
    "hpsx_cache_peer_tier_info": (_int, [_vp, ctypes.POINTER(PeerTierInfoC)]),
    "hpsx_ps_peer_tier_connect_local": (_int, [_vp, _cp]),
    "hpsx_copy_to_host": (_int, [_int, _vp, _vp, _sz]),
    "hpsx_session_lookup_ex": (_int, [_vp, _vpp, _int, _vpp, _int, c_size_p, _sz]),
    "hpsx_session_lookup_batch": (_int, [_vp, _sz, _vpp, _int, _vpp, _int, c_size_p]),
    "hpsx_ps_lookup": (_int, [_vp, _cp, _sz, _vp, _sz, _vp]),
    "hpsx_ps_create_embedding_cache_per_model": (_int, [_vp, _cp]),
    "hpsx_ps_get_embedding_cache": (_int, [_vp, _cp, _int, _vpp]),
    "hpsx_ps_destroy_embedding_cache_per_model": (_int, [_vp, _cp]),
    "hpsx_ps_update_database_per_model": (_int, [_vp, _cp]),
    "hpsx_ps_refresh_embedding_cache": (_int, [_vp, _cp, _int, c_size_p]),
    "hpsx_cache_num_tables": (_int, [_vp, c_size_p]),
    "hpsx_cache_device": (_int, [_vp, ctypes.POINTER(_int)]),
    "hpsx_cache_capacity": (_int, [_vp, _sz, c_size_p]),
    "hpsx_cache_resident": (_int, [_vp, _sz, c_size_p]),
    "hpsx_cache_dump_keys": (_int, [_vp, _sz, _vp, _sz, c_size_p]),
    "hpsx_session_create": (_int, [_vp, _cp, _int, _vpp]),
    "hpsx_session_destroy": (_int, [_vp]),
    "hpsx_session_device": (_int, [_vp, ctypes.POINTER(_int)]),
    "hpsx_session_stream": (_int, [_vp, _vpp]),
    "hpsx_session_lookup": (_int, [_vp, _vpp, _vpp, c_size_p, _sz]),
    "hpsx_session_lookup_device_keys": (_int, [_vp, _vpp, _vpp, c_size_p, _sz]),
    "hpsx_session_lookup_pooled": (_int, [_vp, _sz, _vp, _sz, _sz, _int, _vp]),
    "hpsx_session_lookup_pooled_device_keys": (_int, [_vp, _sz, _vp, _sz, _sz, _int, _vp]),
    "hpsx_session_lookup_pooled_ex": (_int, [_vp, _sz, _vp, _int, _sz, _sz, _int, _vp, _int]),
    "hpsx_session_get_stats": (_int, [_vp, ctypes.POINTER(SessionStatsC)]),
    "hpsx_session_reset_stats": (_int, [_vp]),
    "hpsx_session_set_insert_mode": (_int, [_vp, _int]),
    "hpsx_session_set_debug": (_int, [_vp, _int]),
    "hpsx_session_set_probe_variant": (_int, [_vp, _int]),
    "hpsx_cache_drain_async": (_int, [_vp]),
    "hpsx_mlp_create": (_int, [_int, _sz, c_size_p, _vpp, _vpp, ctypes.POINTER(_int), _vpp]),
    "hpsx_mlp_create_ex": (_int, [_int, _sz, c_size_p, _vpp, _vpp, ctypes.POINTER(_int), _int, _vpp]),
    "hpsx_mlp_forward": (_int, [_vp, _vp, _sz, _vp, _vp]),
    "hpsx_mlp_forward_bf16": (_int, [_vp, _vp, _sz, _vp, _vp]),
    "hpsx_mlp_destroy": (_int, [_vp]),
    "hpsx_session_lookup_bf16_mirror": (_int, [_vp, _sz, _vp, _int, _sz, _vp, _vp]),
    "hpsx_unique": (_int, [_int, _vp, _sz, _vp, _vp, c_size_p, _vp]),
    "hpsx_owner": (ctypes.c_uint32, [ctypes.c_int64, ctypes.c_uint32]),
    "hpsx_owner_batch": (_int, [_vp, _sz, ctypes.c_uint32, _vp]),
    "hpsx_route_keys": (_int, [_int, _vp, _sz, ctypes.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "hpsx_gather_rows": (_int, [_int, _vp, _vp, _sz, _sz, _vp, _vp]),
    "hpsx_scatter_rows": (_int, [_int, _vp, _vp, _sz, _sz, _vp, _vp]),
}

_lib = None


def lib_path() -> str:
    return _LIB


def lib() -> ctypes.CDLL:
    """The engine library.  Raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise ImportError(
                f"{_LIB} is missing: build it with `make` (or __graft_entry__.build()). "
                "There is no Python/CPU fallback for the HPS lookup path.")
        L = ctypes.CDLL(_LIB, mode=ctypes.RTLD_LOCAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        msg = lib().hpsx_last_error()
        raise HpsxError(rc, msg.decode() if msg else "")
