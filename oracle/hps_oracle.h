/*
 * hps_oracle.h — CPU restatement of the HPS lookup contract.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * this library, and only as the checker / the reported CPU baseline.  The product (libhpsx.so,
 * libtriton_hps.so) never links or calls it.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in libhuge_ctr_hps.so (NVIDIA-Merlin/HugeCTR,
 * branch `main`, no commit pin: /root/reference/test/CI.DockerFile:4,11; linked at
 * /root/reference/hps_backend/CMakeLists.txt:147).  It is not vendored under /root/reference, not
 * installed in this image, and the reference ships no golden vectors or unit tests for the path
 * (SURVEY.md §4, §8c).  This file therefore restates the *documented* contract; every function
 * cites the reference text / call site it follows (paths relative to /root/reference).
 */
#ifndef HPS_ORACLE_H_
#define HPS_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hps_oracle_table hps_oracle_table;

/* One embedding table of the volatile (host-memory) database, hash-partitioned into
 * `num_partitions` maps (docs/hierarchical_parameter_server.md:400-416; README.md:133 uses 8). */
hps_oracle_table* hps_oracle_table_create(size_t dim, float default_value, size_t num_partitions);
void hps_oracle_table_destroy(hps_oracle_table* t);
size_t hps_oracle_table_rows(const hps_oracle_table* t);
size_t hps_oracle_table_dim(const hps_oracle_table* t);

/* Load `n` rows: keys[i] -> vectors[i*dim .. (i+1)*dim).  Later duplicates overwrite.  This is the
 * content of the `key` and `emb_vector` files of a sparse model directory
 * (docs/architecture.md:185-218; writer samples/hps-triton-ensemble/01_model_training.ipynb:498-505). */
int hps_oracle_table_insert(hps_oracle_table* t, const int64_t* keys, const float* vectors, size_t n);
/* Load `<dir>/key` + `<dir>/emb_vector`.  Returns the number of rows or -1. */
long long hps_oracle_table_load_dir(hps_oracle_table* t, const char* dir);
/* keys [0,rows) with synthetic rows:
 *   row(k)[j] = bitcast_f32(0x3F800000 | (splitmix64(k*131 + j + seed) >> 41)) - 1.5            */
void hps_oracle_table_fill_procedural(hps_oracle_table* t, size_t rows, uint64_t seed,
                                      size_t num_threads);
float hps_oracle_synth_value(int64_t key, uint32_t j, uint64_t seed);

/* HierParameterServerBase::lookup(h_keys, n, h_vectors, model, table) — CPU path:
 * volatile-db fetch; a key found in no database gets default_value_for_each_table
 * (docs/hierarchical_parameter_server.md:67-78,244-246).  out is [n, dim] row-major.
 * `num_threads` key ranges are served concurrently (the reference fans out over a pool sized
 * HCTR_DEFAULT_CONCURRENCY / hardware_concurrency: hps_backend/src/thread_pool.cpp:25-41).
 * Returns the number of absent keys. */
size_t hps_oracle_lookup(const hps_oracle_table* t, const int64_t* keys, size_t n, float* out,
                         size_t num_threads);

/* One Triton request (hps_backend/src/hps.cc:586-630, src/model_instance_state.cpp:180-195):
 * KEYS is table-major (docs/architecture.md:220-230), NUMKEYS[t] keys belong to table t,
 * OUTPUT0 = concat_t [ row_t(k) for k in KEYS_t ], length sum_t NUMKEYS[t]*dim_t.
 * Returns the number of floats written. */
size_t hps_oracle_request(const hps_oracle_table* const* tables, size_t num_tables,
                          const int64_t* keys, const int32_t* numkeys, float* out,
                          size_t num_threads);

/* Slot-wise combiner (north-star stage a8): keys [num_bags, hotness];
 * out[b] = sum_{j<hotness} row(keys[b,j]) accumulated in ascending j in fp32; mean divides by
 * hotness (combiner: 0 = sum, 1 = mean). */
void hps_oracle_pooled(const hps_oracle_table* t, const int64_t* keys, size_t num_bags,
                       size_t hotness, int combiner, float* out);

/* Dedup in first-occurrence order: unique[inverse[i]] == keys[i].  Returns the number of unique
 * keys.  (What [UPSTREAM] unique_op computes, canonicalised: SURVEY.md §8c "bit-exact on indices".) */
size_t hps_oracle_unique(const int64_t* keys, size_t n, int64_t* unique, uint32_t* inverse);

/* Shard owning `key` in the model-parallel mode: the low 32 bits of the 64-bit
 * murmur3 finaliser, range-reduced by multiply-shift. */
uint32_t hps_oracle_owner(int64_t key, uint32_t num_shards);

#ifdef __cplusplus
}
#endif
#endif /* HPS_ORACLE_H_ */
