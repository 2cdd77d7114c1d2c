#!/usr/bin/env python
"""Generates tests/golden/hps_kat.npz + the sparse-model directory fixture — run once, outputs committed.

PARITY UNPINNED: the reference ships no golden vectors for the lookup path and its arithmetic lives in the
un-vendored libhuge_ctr_hps.so, so nothing here comes from a reference binary.  What IS taken
from the reference are the artefact formats, restated from its samples and rebuilt here WITHOUT the oracle:
  * the sparse-model writer of hps-triton-ensemble/01_model_training.ipynb:498-505 (struct.pack('q', key) +
    struct.pack('<D>f', *vec)), used verbatim below to produce tests/golden/sparse_d4/{key,emb_vector};
  * the W&D request builder of Hierarchical_Parameter_Server_Deployment.ipynb:738-753: KEYS = table-1 keys then
    table-2 keys, NUMKEYS = [[batch*2, batch*26]], OUTPUT0 shape [batch*2*d1 + batch*26*d2] = [4180];
  * the ensemble request of 02_model_inference_hps_tf_ensemble.ipynb:661-662 (keys 1..9, heavy duplication,
    default value 1.0 for absent keys :220).
Expected outputs are computed with plain Python dict lookups written out longhand in this file (not with
oracle/), so that the oracle, the C oracle, the engine and the CUDA path are all checked against an independent
statement of the contract.
"""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def synth_value_bits(key, j, seed):
    """bitcast_f32(0x3F800000 | (splitmix64(key*131 + j + seed) >> 41)) - 1.5, as fp32 bits."""
    r = splitmix64(((key & M64) * 131 + j + seed) & M64)
    f = np.uint32(0x3F800000 | (r >> 41)).view(np.float32) - np.float32(1.5)
    return np.float32(f).view(np.uint32)


def fmix64(h):
    h &= M64
    h ^= h >> 33
    h = (h * 0xFF51AFD7ED558CCD) & M64
    h ^= h >> 33
    h = (h * 0xC4CEB9FE1A85EC53) & M64
    h ^= h >> 33
    return h


def owner(key, shards):
    return ((fmix64(key) & 0xFFFFFFFF) * shards) >> 32


def dict_lookup(table, default, dim, keys):
    return np.array([table.get(int(k), [default] * dim) for k in keys], dtype=np.float32)


def convert_to_sparse_model(embeddings_weights, embedding_table_path, embedding_vec_size):
    # the reference sample's writer (01_model_training.ipynb:498-505), keys = row indices
    os.makedirs(embedding_table_path, exist_ok=True)
    with open("{}/key".format(embedding_table_path), "wb") as key_file, \
            open("{}/emb_vector".format(embedding_table_path), "wb") as vec_file:
        for key in range(embeddings_weights.shape[0]):
            vec = embeddings_weights[key]
            key_file.write(struct.pack("q", key))
            vec_file.write(struct.pack(str(embedding_vec_size) + "f", *vec))


def main():
    rng = np.random.default_rng(20261017)
    g = {}
    # --- procedural rows and routing hash, pure-Python integer arithmetic ---------------------------------
    synth_keys = np.array([0, 1, 2, 12345, 9_999_999, 99_999_999, 999_999_999, -1, -(1 << 63), (1 << 63) - 1], dtype=np.int64)
    seed = 0xB2000002
    g["synth_keys"], g["synth_seed"] = synth_keys, np.uint64(seed)
    g["synth_bits_dim8"] = np.array([[synth_value_bits(int(k), j, seed) for j in range(8)] for k in synth_keys], dtype=np.uint32)
    g["owner_keys"] = synth_keys
    g["owner_8"] = np.array([owner(int(k), 8) for k in synth_keys], dtype=np.uint32)
    g["owner_3"] = np.array([owner(int(k), 3) for k in synth_keys], dtype=np.uint32)

    # --- W&D sample request: 10 samples, tables d=[1,16], 2+26 keys/sample -> 4180 floats -------------------
    k0 = rng.choice(10_000, size=40, replace=False).astype(np.int64)
    v0 = rng.standard_normal((40, 1)).astype(np.float32)
    k1 = rng.choice(1_000_000, size=300, replace=False).astype(np.int64)
    v1 = rng.standard_normal((300, 16)).astype(np.float32)
    t0 = {int(k): v.tolist() for k, v in zip(k0, v0)}
    t1 = {int(k): v.tolist() for k, v in zip(k1, v1)}
    batch = 10
    q0 = rng.choice(k0, size=batch * 2)
    q1 = rng.choice(k1, size=batch * 26)
    q1[5], q1[77] = 7_000_001, -3  # two keys that are in no table -> default
    keys = np.array([list(q0) + list(q1)], dtype="int64")
    row_ptrs = np.array([[batch * 2, batch * 26]], dtype="int32")
    out = np.concatenate([dict_lookup(t0, 0.0, 1, q0).ravel(), dict_lookup(t1, 0.25, 16, q1).ravel()])
    assert out.shape == (4180,)
    g.update(wdl_k0=k0, wdl_v0=v0, wdl_k1=k1, wdl_v1=v1, wdl_KEYS=keys, wdl_NUMKEYS=row_ptrs, wdl_OUTPUT0=out,
             wdl_defaults=np.array([0.0, 0.25], dtype=np.float32))

    # --- ensemble sample: one table dim 16, 3 slots, keys 1..9 repeated, default 1.0 --------------------------
    ke = np.arange(0, 8, dtype=np.int64)  # keys 0..7 loaded; 8 and 9 absent
    ve = rng.standard_normal((8, 16)).astype(np.float32)
    te = {int(k): v.tolist() for k, v in zip(ke, ve)}
    req = rng.integers(1, 10, size=(64, 3)).astype(np.int64)
    g.update(ens_keys=ke, ens_vecs=ve, ens_KEYS=req.reshape(1, -1), ens_NUMKEYS=np.array([[req.size]], dtype=np.int32),
             ens_OUTPUT0=dict_lookup(te, 1.0, 16, req.ravel()).ravel())
    # slot-wise sum / mean over the 3 keys of a sample, fp32, ascending slot order (north-star stage a8)
    rows = dict_lookup(te, 1.0, 16, req.ravel()).reshape(64, 3, 16)
    acc = np.zeros((64, 16), dtype=np.float32)
    for j in range(3):
        acc = (acc + rows[:, j]).astype(np.float32)
    g["ens_pooled_sum"] = acc
    g["ens_pooled_mean"] = (acc / np.float32(3)).astype(np.float32)
    # dedup artefacts (first-occurrence order)
    seen, uniq, inv = {}, [], []
    for k in req.ravel():
        if int(k) not in seen:
            seen[int(k)] = len(uniq)
            uniq.append(int(k))
        inv.append(seen[int(k)])
    g["ens_unique"], g["ens_inverse"] = np.array(uniq, dtype=np.int64), np.array(inv, dtype=np.uint32)

    # --- sparse model directory written by the reference sample's own writer ----------------------------------
    w = rng.standard_normal((37, 4)).astype(np.float32)
    convert_to_sparse_model(w, os.path.join(HERE, "sparse_d4"), 4)
    g["sparse_d4_weights"] = w

    np.savez(os.path.join(HERE, "hps_kat.npz"), **g)
    print("wrote", os.path.join(HERE, "hps_kat.npz"), {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
