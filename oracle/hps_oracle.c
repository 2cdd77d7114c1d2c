/*
 * hps_oracle.c — CPU restatement of the HPS lookup contract.  TEST INFRASTRUCTURE ONLY; see the
 * header for who may use it and for the "PARITY UNPINNED" statement.  Plain C11 + pthreads.
 *
 * Structure follows the reference's CPU ParameterServer path (gpucache = false): a table is
 * hash-partitioned into `num_partitions` maps held in host memory
 * (/root/reference/docs/hierarchical_parameter_server.md:400-416); a lookup finds the key in its
 * partition and copies `dim` floats, or fills the table's default value (:244-246).
 */
#define _POSIX_C_SOURCE 200809L
#include "hps_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

typedef struct {
  int64_t key;
  uint8_t used;
  size_t row; /* index into `rows` */
} slot_t;

typedef struct {
  slot_t* slots;
  size_t cap; /* power of two */
  size_t count;
  float* rows; /* [rows_cap, dim] */
  size_t rows_cap;
} part_t;

struct hps_oracle_table {
  size_t dim;
  float default_value;
  size_t num_parts;
  part_t* parts;
  size_t rows;
};

/* Fibonacci-style multiplicative hash; deliberately NOT the engine's hash: returned values must
 * not depend on how either side hashes. */
static uint64_t oracle_hash(int64_t key) {
  uint64_t x = (uint64_t)key;
  x ^= x >> 29;
  x *= 0x9E3779B97F4A7C15ULL;
  x ^= x >> 32;
  x *= 0xD6E8FEB86659FD93ULL;
  x ^= x >> 29;
  return x;
}

hps_oracle_table* hps_oracle_table_create(size_t dim, float default_value, size_t num_partitions) {
  if (dim == 0) return NULL;
  if (num_partitions == 0) num_partitions = 8; /* README.md:133 */
  hps_oracle_table* t = (hps_oracle_table*)calloc(1, sizeof(*t));
  if (!t) return NULL;
  t->dim = dim;
  t->default_value = default_value;
  t->num_parts = num_partitions;
  t->parts = (part_t*)calloc(num_partitions, sizeof(part_t));
  if (!t->parts) {
    free(t);
    return NULL;
  }
  for (size_t p = 0; p < num_partitions; ++p) {
    t->parts[p].cap = 64;
    t->parts[p].slots = (slot_t*)calloc(64, sizeof(slot_t));
  }
  return t;
}

void hps_oracle_table_destroy(hps_oracle_table* t) {
  if (!t) return;
  for (size_t p = 0; p < t->num_parts; ++p) {
    free(t->parts[p].slots);
    free(t->parts[p].rows);
  }
  free(t->parts);
  free(t);
}

size_t hps_oracle_table_rows(const hps_oracle_table* t) { return t ? t->rows : 0; }
size_t hps_oracle_table_dim(const hps_oracle_table* t) { return t ? t->dim : 0; }

static int part_grow(part_t* p) {
  const size_t ncap = p->cap * 2;
  slot_t* ns = (slot_t*)calloc(ncap, sizeof(slot_t));
  if (!ns) return -1;
  for (size_t i = 0; i < p->cap; ++i) {
    if (!p->slots[i].used) continue;
    size_t j = (oracle_hash(p->slots[i].key) >> 20) & (ncap - 1);
    while (ns[j].used) j = (j + 1) & (ncap - 1);
    ns[j] = p->slots[i];
  }
  free(p->slots);
  p->slots = ns;
  p->cap = ncap;
  return 0;
}

static float* part_upsert(hps_oracle_table* t, part_t* p, int64_t key) {
  if ((p->count + 1) * 2 > p->cap && part_grow(p) != 0) return NULL;
  size_t j = (oracle_hash(key) >> 20) & (p->cap - 1);
  while (p->slots[j].used) {
    if (p->slots[j].key == key) return p->rows + p->slots[j].row * t->dim;
    j = (j + 1) & (p->cap - 1);
  }
  if (p->count == p->rows_cap) {
    const size_t nc = p->rows_cap ? p->rows_cap * 2 : 256;
    float* nr = (float*)realloc(p->rows, nc * t->dim * sizeof(float));
    if (!nr) return NULL;
    p->rows = nr;
    p->rows_cap = nc;
  }
  p->slots[j].used = 1;
  p->slots[j].key = key;
  p->slots[j].row = p->count++;
  return p->rows + p->slots[j].row * t->dim;
}

static const float* part_find(const hps_oracle_table* t, const part_t* p, int64_t key) {
  size_t j = (oracle_hash(key) >> 20) & (p->cap - 1);
  while (p->slots[j].used) {
    if (p->slots[j].key == key) return p->rows + p->slots[j].row * t->dim;
    j = (j + 1) & (p->cap - 1);
  }
  return NULL;
}

static size_t part_of(const hps_oracle_table* t, int64_t key) {
  return (size_t)(oracle_hash(key) % t->num_parts);
}

int hps_oracle_table_insert(hps_oracle_table* t, const int64_t* keys, const float* vectors, size_t n) {
  if (!t) return -1;
  for (size_t i = 0; i < n; ++i) {
    part_t* p = &t->parts[part_of(t, keys[i])];
    const size_t before = p->count;
    float* dst = part_upsert(t, p, keys[i]);
    if (!dst) return -1;
    if (p->count != before) ++t->rows;
    memcpy(dst, vectors + i * t->dim, t->dim * sizeof(float));
  }
  return 0;
}

long long hps_oracle_table_load_dir(hps_oracle_table* t, const char* dir) {
  char kp[4096], vp[4096];
  snprintf(kp, sizeof kp, "%s/key", dir);
  snprintf(vp, sizeof vp, "%s/emb_vector", dir);
  struct stat ks, vs;
  if (stat(kp, &ks) != 0 || stat(vp, &vs) != 0) return -1;
  const size_t row_bytes = t->dim * sizeof(float);
  if ((size_t)vs.st_size % row_bytes) return -1;
  const size_t rows = (size_t)vs.st_size / row_bytes;
  if (rows == 0) return 0;
  size_t kb;
  if ((size_t)ks.st_size == rows * 8)
    kb = 8; /* "long long" keys, struct.pack('q', key) in the sample writer */
  else if ((size_t)ks.st_size == rows * 4)
    kb = 4; /* supportlonglong = false: unsigned int keys */
  else
    return -1;
  FILE* kf = fopen(kp, "rb");
  FILE* vf = fopen(vp, "rb");
  if (!kf || !vf) {
    if (kf) fclose(kf);
    if (vf) fclose(vf);
    return -1;
  }
  float* vec = (float*)malloc(row_bytes);
  long long rc = (long long)rows;
  for (size_t i = 0; i < rows; ++i) {
    int64_t key;
    if (kb == 8) {
      if (fread(&key, 8, 1, kf) != 1) rc = -1;
    } else {
      uint32_t k32;
      if (fread(&k32, 4, 1, kf) != 1) rc = -1;
      key = (int64_t)k32;
    }
    if (rc < 0 || fread(vec, row_bytes, 1, vf) != 1 || hps_oracle_table_insert(t, &key, vec, 1) != 0) {
      rc = -1;
      break;
    }
  }
  free(vec);
  fclose(kf);
  fclose(vf);
  return rc;
}

/* splitmix64 (Steele, Lea, Flood 2014), public-domain constants. */
static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

float hps_oracle_synth_value(int64_t key, uint32_t j, uint64_t seed) {
  const uint64_t r = splitmix64((uint64_t)key * 131ULL + j + seed);
  const uint32_t bits = 0x3F800000u | (uint32_t)(r >> 41);
  float f;
  memcpy(&f, &bits, 4);
  return f - 1.5f;
}

typedef struct {
  hps_oracle_table* t;
  size_t part, rows;
  uint64_t seed;
} fill_arg;

static void* fill_worker(void* vp) {
  fill_arg* a = (fill_arg*)vp;
  hps_oracle_table* t = a->t;
  part_t* p = &t->parts[a->part];
  for (size_t k = 0; k < a->rows; ++k) {
    if (part_of(t, (int64_t)k) != a->part) continue;
    float* dst = part_upsert(t, p, (int64_t)k);
    if (!dst) return (void*)1;
    for (size_t j = 0; j < t->dim; ++j) dst[j] = hps_oracle_synth_value((int64_t)k, (uint32_t)j, a->seed);
  }
  return NULL;
}

void hps_oracle_table_fill_procedural(hps_oracle_table* t, size_t rows, uint64_t seed,
                                      size_t num_threads) {
  /* one worker per partition (partitions are independent maps), at most num_threads at a time */
  if (num_threads == 0) num_threads = 1;
  const size_t P = t->num_parts;
  pthread_t* th = (pthread_t*)malloc(P * sizeof(pthread_t));
  fill_arg* args = (fill_arg*)malloc(P * sizeof(fill_arg));
  for (size_t base = 0; base < P; base += num_threads) {
    const size_t m = (P - base < num_threads) ? P - base : num_threads;
    for (size_t i = 0; i < m; ++i) {
      args[base + i].t = t;
      args[base + i].part = base + i;
      args[base + i].rows = rows;
      args[base + i].seed = seed;
      pthread_create(&th[base + i], NULL, fill_worker, &args[base + i]);
    }
    for (size_t i = 0; i < m; ++i) pthread_join(th[base + i], NULL);
  }
  size_t total = 0;
  for (size_t p = 0; p < P; ++p) total += t->parts[p].count;
  t->rows = total;
  free(th);
  free(args);
}

static size_t lookup_range(const hps_oracle_table* t, const int64_t* keys, size_t b, size_t e,
                           float* out) {
  size_t absent = 0;
  for (size_t i = b; i < e; ++i) {
    const float* row = part_find(t, &t->parts[part_of(t, keys[i])], keys[i]);
    float* dst = out + i * t->dim;
    if (row) {
      memcpy(dst, row, t->dim * sizeof(float));
    } else {
      for (size_t j = 0; j < t->dim; ++j) dst[j] = t->default_value;
      ++absent;
    }
  }
  return absent;
}

typedef struct {
  const hps_oracle_table* t;
  const int64_t* keys;
  size_t b, e;
  float* out;
  size_t absent;
} lookup_arg;

static void* lookup_worker(void* vp) {
  lookup_arg* a = (lookup_arg*)vp;
  a->absent = lookup_range(a->t, a->keys, a->b, a->e, a->out);
  return NULL;
}

size_t hps_oracle_lookup(const hps_oracle_table* t, const int64_t* keys, size_t n, float* out,
                         size_t num_threads) {
  if (!t || n == 0) return 0;
  if (num_threads <= 1 || n < 4096) return lookup_range(t, keys, 0, n, out);
  if (num_threads > 1024) num_threads = 1024;
  pthread_t* th = (pthread_t*)malloc(num_threads * sizeof(pthread_t));
  lookup_arg* args = (lookup_arg*)malloc(num_threads * sizeof(lookup_arg));
  const size_t chunk = (n + num_threads - 1) / num_threads;
  size_t started = 0;
  for (size_t i = 0; i < num_threads; ++i) {
    const size_t b = i * chunk;
    if (b >= n) break;
    args[i].t = t;
    args[i].keys = keys;
    args[i].b = b;
    args[i].e = (b + chunk < n) ? b + chunk : n;
    args[i].out = out;
    args[i].absent = 0;
    pthread_create(&th[i], NULL, lookup_worker, &args[i]);
    ++started;
  }
  size_t absent = 0;
  for (size_t i = 0; i < started; ++i) {
    pthread_join(th[i], NULL);
    absent += args[i].absent;
  }
  free(th);
  free(args);
  return absent;
}

size_t hps_oracle_request(const hps_oracle_table* const* tables, size_t num_tables,
                          const int64_t* keys, const int32_t* numkeys, float* out,
                          size_t num_threads) {
  size_t koff = 0, ooff = 0;
  for (size_t t = 0; t < num_tables; ++t) {
    const size_t n = (size_t)numkeys[t];
    hps_oracle_lookup(tables[t], keys + koff, n, out + ooff, num_threads);
    koff += n;
    ooff += n * tables[t]->dim;
  }
  return ooff;
}

void hps_oracle_pooled(const hps_oracle_table* t, const int64_t* keys, size_t num_bags,
                       size_t hotness, int combiner, float* out) {
  float* row = (float*)malloc(t->dim * sizeof(float));
  for (size_t b = 0; b < num_bags; ++b) {
    float* acc = out + b * t->dim;
    for (size_t j = 0; j < t->dim; ++j) acc[j] = 0.0f;
    for (size_t h = 0; h < hotness; ++h) {
      lookup_range(t, keys + b * hotness + h, 0, 1, row);
      for (size_t j = 0; j < t->dim; ++j) acc[j] = acc[j] + row[j];
    }
    if (combiner == 1)
      for (size_t j = 0; j < t->dim; ++j) acc[j] = acc[j] / (float)hotness;
  }
  free(row);
}

size_t hps_oracle_unique(const int64_t* keys, size_t n, int64_t* unique, uint32_t* inverse) {
  size_t cap = 64;
  while (cap < 2 * n) cap <<= 1;
  int64_t* hk = (int64_t*)malloc(cap * sizeof(int64_t));
  uint32_t* hv = (uint32_t*)malloc(cap * sizeof(uint32_t));
  uint8_t* used = (uint8_t*)calloc(cap, 1);
  size_t u = 0;
  for (size_t i = 0; i < n; ++i) {
    size_t j = (oracle_hash(keys[i]) >> 16) & (cap - 1);
    while (used[j] && hk[j] != keys[i]) j = (j + 1) & (cap - 1);
    if (!used[j]) {
      used[j] = 1;
      hk[j] = keys[i];
      hv[j] = (uint32_t)u;
      unique[u++] = keys[i];
    }
    inverse[i] = hv[j];
  }
  free(hk);
  free(hv);
  free(used);
  return u;
}

/* MurmurHash3 fmix64 (Appleby, public domain). */
static uint64_t fmix64(uint64_t h) {
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdULL;
  h ^= h >> 33;
  h *= 0xc4ceb9fe1a85ec53ULL;
  h ^= h >> 33;
  return h;
}

uint32_t hps_oracle_owner(int64_t key, uint32_t num_shards) {
  const uint64_t h = fmix64((uint64_t)key);
  return (uint32_t)(((h & 0xffffffffULL) * (uint64_t)num_shards) >> 32);
}
