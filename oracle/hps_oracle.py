"""CPU restatement of the HPS lookup contract — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker or the reported CPU baseline.  The product
(``hugectr_backend_b200``) never imports it.

PARITY UNPINNED: the path's arithmetic lives in ``libhuge_ctr_hps.so`` (NVIDIA-Merlin/HugeCTR, branch
``main``, unpinned: ``/root/reference/test/CI.DockerFile:4,11``), which is neither vendored under
``/root/reference`` nor installed here, and the reference has no tests or golden vectors for the path
(SURVEY.md §4, §8c).  What is restated is the documented contract:

* request layout: ``KEYS`` table-major, ``NUMKEYS[t]`` keys for table ``t``
  (``docs/architecture.md:220-230``; builder ``hps_backend/samples/Hierarchical_Parameter_Server_Deployment.ipynb:738-742``)
* output: ``concat_t [row_t(k) for k in KEYS_t]``, ``sum_t NUMKEYS[t]*dim_t`` floats
  (``hps_backend/src/hps.cc:620-630``, ``hps_backend/src/model_instance_state.cpp:185-193``)
* absent key -> ``default_value_for_each_table[t]`` (``docs/hierarchical_parameter_server.md:244-246``)
* asynchronous insertion: a cache miss is answered with the default vector in this response
  (``docs/architecture.md:31-32,65-67``)
* sparse model files ``<dir>/key`` (int64) + ``<dir>/emb_vector`` (fp32 row-major)
  (``docs/architecture.md:185-218``; writer ``samples/hps-triton-ensemble/01_model_training.ipynb:498-505``)

Two independent implementations live here: pure numpy (sorted keys + ``searchsorted``) and a ctypes
binding of ``oracle/hps_oracle.c`` (hash-partitioned maps + pthreads, the timed CPU baseline).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Iterable, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhps_oracle.so")

_M64 = (1 << 64) - 1


# ---------------------------------------------------------------------------------------------
# synthetic rows
# ---------------------------------------------------------------------------------------------
def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def synth_rows(keys: np.ndarray, dim: int, seed: int) -> np.ndarray:
    """row(k)[j] = bitcast_f32(0x3F800000 | (splitmix64(k*131 + j + seed) >> 41)) - 1.5"""
    keys = np.asarray(keys, dtype=np.int64)
    with np.errstate(over="ignore"):
        base = keys.astype(np.uint64)[:, None] * np.uint64(131) + np.arange(dim, dtype=np.uint64)[None, :]
        base = base + np.uint64(seed & _M64)
        r = splitmix64(base)
    bits = (np.uint32(0x3F800000) | (r >> np.uint64(41)).astype(np.uint32)).astype(np.uint32)
    return bits.view(np.float32) - np.float32(1.5)


def fmix64(h: np.ndarray) -> np.ndarray:
    """MurmurHash3 64-bit finaliser."""
    h = np.asarray(h).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = h ^ (h >> np.uint64(33))
        h = h * np.uint64(0xFF51AFD7ED558CCD)
        h = h ^ (h >> np.uint64(33))
        h = h * np.uint64(0xC4CEB9FE1A85EC53)
        h = h ^ (h >> np.uint64(33))
    return h


def owner(keys: np.ndarray, num_shards: int) -> np.ndarray:
    """Shard owning each key in the model-parallel mode."""
    h = fmix64(np.asarray(keys, dtype=np.int64).astype(np.uint64))
    lo = h & np.uint64(0xFFFFFFFF)
    return ((lo * np.uint64(num_shards)) >> np.uint64(32)).astype(np.uint32)


# ---------------------------------------------------------------------------------------------
# numpy oracle
# ---------------------------------------------------------------------------------------------
class NumpyTable:
    """One embedding table: explicit (key,row) pairs and/or procedural rows for keys [0, rows)."""

    def __init__(self, dim: int, default_value: float = 0.0):
        self.dim = int(dim)
        self.default_value = np.float32(default_value)
        self._keys = np.empty(0, dtype=np.int64)
        self._rows = np.empty((0, self.dim), dtype=np.float32)
        self._proc_rows = 0
        self._proc_seed = 0

    def insert(self, keys: np.ndarray, vectors: np.ndarray) -> None:
        keys = np.asarray(keys, dtype=np.int64).ravel()
        vectors = np.asarray(vectors, dtype=np.float32).reshape(len(keys), self.dim)
        allk = np.concatenate([self._keys, keys])
        allv = np.concatenate([self._rows, vectors])
        # later duplicates overwrite: keep the LAST occurrence of every key
        _, first_of_reversed = np.unique(allk[::-1], return_index=True)
        keep = len(allk) - 1 - first_of_reversed
        order = np.argsort(allk[keep], kind="stable")
        self._keys = allk[keep][order]
        self._rows = allv[keep][order]

    def load_dir(self, path: str, key_dtype=np.int64) -> int:
        keys = np.fromfile(os.path.join(path, "key"), dtype=key_dtype).astype(np.int64)
        vecs = np.fromfile(os.path.join(path, "emb_vector"), dtype=np.float32).reshape(-1, self.dim)
        assert len(keys) == len(vecs)
        self.insert(keys, vecs)
        return len(keys)

    def fill_procedural(self, rows: int, seed: int) -> None:
        self._proc_rows = int(rows)
        self._proc_seed = int(seed)

    @property
    def rows(self) -> int:
        extra = int(np.count_nonzero((self._keys < 0) | (self._keys >= self._proc_rows)))
        return self._proc_rows + extra

    def lookup(self, keys: np.ndarray) -> np.ndarray:
        keys = np.asarray(keys, dtype=np.int64).ravel()
        out = np.full((len(keys), self.dim), self.default_value, dtype=np.float32)
        if self._proc_rows:
            m = (keys >= 0) & (keys < self._proc_rows)
            if m.any():
                out[m] = synth_rows(keys[m], self.dim, self._proc_seed)
        if len(self._keys):
            pos = np.searchsorted(self._keys, keys)
            pos_c = np.minimum(pos, len(self._keys) - 1)
            hit = self._keys[pos_c] == keys
            out[hit] = self._rows[pos_c[hit]]
        return out

    def contains(self, keys: np.ndarray) -> np.ndarray:
        keys = np.asarray(keys, dtype=np.int64).ravel()
        m = (keys >= 0) & (keys < self._proc_rows)
        if len(self._keys):
            pos = np.minimum(np.searchsorted(self._keys, keys), len(self._keys) - 1)
            m |= self._keys[pos] == keys
        return m


def request(tables: Sequence[NumpyTable], keys: np.ndarray, numkeys: Iterable[int]) -> np.ndarray:
    """OUTPUT0 of one Triton request (hps.cc:586-630): table-major concat of un-pooled rows."""
    keys = np.asarray(keys, dtype=np.int64).ravel()
    outs, off = [], 0
    for t, n in zip(tables, numkeys):
        n = int(n)
        outs.append(t.lookup(keys[off:off + n]).ravel())
        off += n
    assert off == len(keys), "NUMKEYS does not cover KEYS"
    return np.concatenate(outs) if outs else np.empty(0, dtype=np.float32)


def request_async_mode(tables: Sequence[NumpyTable], keys, numkeys, resident: Sequence[np.ndarray]) -> np.ndarray:
    """Asynchronous-insertion response: keys not resident in the GPU cache get the default vector."""
    keys = np.asarray(keys, dtype=np.int64).ravel()
    outs, off = [], 0
    for t, n, res in zip(tables, numkeys, resident):
        n = int(n)
        k = keys[off:off + n]
        rows = t.lookup(k)
        miss = ~np.isin(k, np.asarray(res, dtype=np.int64))
        rows[miss] = t.default_value
        outs.append(rows.ravel())
        off += n
    return np.concatenate(outs) if outs else np.empty(0, dtype=np.float32)


def pooled(table: NumpyTable, keys: np.ndarray, num_bags: int, hotness: int, combiner: str = "sum") -> np.ndarray:
    """dense[b] = sum_j row(keys[b,j]) accumulated in ascending j in fp32 (mean: / hotness)."""
    rows = table.lookup(np.asarray(keys, dtype=np.int64).ravel()).reshape(num_bags, hotness, table.dim)
    acc = np.zeros((num_bags, table.dim), dtype=np.float32)
    for j in range(hotness):
        acc = (acc + rows[:, j, :]).astype(np.float32)
    if combiner == "mean":
        acc = (acc / np.float32(hotness)).astype(np.float32)
    return acc


def unique_first_occurrence(keys: np.ndarray):
    """(unique, inverse) with unique in first-occurrence order."""
    keys = np.asarray(keys, dtype=np.int64).ravel()
    u, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return u[order], rank[inv].astype(np.uint32)


def write_sparse_dir(path: str, keys: np.ndarray, vectors: np.ndarray, key_dtype=np.int64) -> None:
    """Sparse model directory in the reference's format (01_model_training.ipynb:498-505)."""
    os.makedirs(path, exist_ok=True)
    np.asarray(keys).astype(key_dtype).tofile(os.path.join(path, "key"))
    np.asarray(vectors, dtype=np.float32).tofile(os.path.join(path, "emb_vector"))


# ---------------------------------------------------------------------------------------------
# C oracle (ctypes)
# ---------------------------------------------------------------------------------------------
def build_c_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "hps_oracle.c")
    hdr = os.path.join(_HERE, "hps_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-Wall", "-Wextra", "-fPIC", "-shared", "-pthread",
                               "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


_lib = None


def c_lib():
    global _lib
    if _lib is None:
        build_c_oracle()
        L = ctypes.CDLL(_LIB_PATH)
        vp, sz, i64p, f32p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p
        L.hps_oracle_table_create.restype = vp
        L.hps_oracle_table_create.argtypes = [sz, ctypes.c_float, sz]
        L.hps_oracle_table_destroy.argtypes = [vp]
        L.hps_oracle_table_rows.restype = sz
        L.hps_oracle_table_rows.argtypes = [vp]
        L.hps_oracle_table_insert.restype = ctypes.c_int
        L.hps_oracle_table_insert.argtypes = [vp, i64p, f32p, sz]
        L.hps_oracle_table_load_dir.restype = ctypes.c_longlong
        L.hps_oracle_table_load_dir.argtypes = [vp, ctypes.c_char_p]
        L.hps_oracle_table_fill_procedural.argtypes = [vp, sz, ctypes.c_uint64, sz]
        L.hps_oracle_synth_value.restype = ctypes.c_float
        L.hps_oracle_synth_value.argtypes = [ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64]
        L.hps_oracle_lookup.restype = sz
        L.hps_oracle_lookup.argtypes = [vp, i64p, sz, f32p, sz]
        L.hps_oracle_request.restype = sz
        L.hps_oracle_request.argtypes = [vp, sz, i64p, vp, f32p, sz]
        L.hps_oracle_pooled.argtypes = [vp, i64p, sz, sz, ctypes.c_int, f32p]
        L.hps_oracle_unique.restype = sz
        L.hps_oracle_unique.argtypes = [i64p, sz, vp, vp]
        L.hps_oracle_owner.restype = ctypes.c_uint32
        L.hps_oracle_owner.argtypes = [ctypes.c_int64, ctypes.c_uint32]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class CTable:
    """ctypes handle of one ``hps_oracle_table``."""

    def __init__(self, dim: int, default_value: float = 0.0, num_partitions: int = 8):
        self.L = c_lib()
        self.dim = int(dim)
        self.h = self.L.hps_oracle_table_create(dim, default_value, num_partitions)
        if not self.h:
            raise MemoryError("hps_oracle_table_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hps_oracle_table_destroy(self.h)
            self.h = None

    @property
    def rows(self) -> int:
        return int(self.L.hps_oracle_table_rows(self.h))

    def insert(self, keys, vectors) -> None:
        keys = np.ascontiguousarray(keys, dtype=np.int64).ravel()
        vectors = np.ascontiguousarray(vectors, dtype=np.float32).reshape(len(keys), self.dim)
        if self.L.hps_oracle_table_insert(self.h, _ptr(keys), _ptr(vectors), len(keys)) != 0:
            raise MemoryError("hps_oracle_table_insert failed")

    def load_dir(self, path: str) -> int:
        n = self.L.hps_oracle_table_load_dir(self.h, path.encode())
        if n < 0:
            raise IOError(f"cannot load sparse model directory {path}")
        return int(n)

    def fill_procedural(self, rows: int, seed: int, threads: int = 1) -> None:
        self.L.hps_oracle_table_fill_procedural(self.h, rows, seed & _M64, threads)

    def lookup(self, keys, threads: int = 1, out: np.ndarray | None = None) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.int64).ravel()
        if out is None:
            out = np.empty((len(keys), self.dim), dtype=np.float32)
        self.L.hps_oracle_lookup(self.h, _ptr(keys), len(keys), _ptr(out), threads)
        return out

    def pooled(self, keys, num_bags: int, hotness: int, combiner: str = "sum") -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.int64).ravel()
        out = np.empty((num_bags, self.dim), dtype=np.float32)
        self.L.hps_oracle_pooled(self.h, _ptr(keys), num_bags, hotness, 1 if combiner == "mean" else 0, _ptr(out))
        return out


def c_request(tables: Sequence[CTable], keys, numkeys, threads: int = 1) -> np.ndarray:
    L = c_lib()
    keys = np.ascontiguousarray(keys, dtype=np.int64).ravel()
    numkeys = np.ascontiguousarray(numkeys, dtype=np.int32).ravel()
    total = int(sum(int(n) * t.dim for n, t in zip(numkeys, tables)))
    out = np.empty(total, dtype=np.float32)
    arr = (ctypes.c_void_p * len(tables))(*[t.h for t in tables])
    wrote = L.hps_oracle_request(arr, len(tables), _ptr(keys), _ptr(numkeys), _ptr(out), threads)
    assert wrote == total
    return out


def c_unique(keys):
    L = c_lib()
    keys = np.ascontiguousarray(keys, dtype=np.int64).ravel()
    uniq = np.empty(len(keys), dtype=np.int64)
    inv = np.empty(len(keys), dtype=np.uint32)
    u = L.hps_oracle_unique(_ptr(keys), len(keys), _ptr(uniq), _ptr(inv))
    return uniq[:u].copy(), inv


def c_owner(key: int, shards: int) -> int:
    return int(c_lib().hps_oracle_owner(int(key), int(shards)))
