/*
 * cfk_oracle.c — CPU restatement (TEST INFRASTRUCTURE ONLY) of centroFlye's unique-k-mer
 * recruitment, for inputs too large for oracle/py_oracle.py and as the timed CPU baseline of
 * bench.py.  Nothing under centroflye_b200/ links or loads this file.
 *
 * Parity status: PINNED through tests/test_c_oracle.py, which checks every function here against
 * tests/golden/ (outputs of the unmodified reference, oracle/make_golden.py).
 *
 * Reference lines followed (relative to /root/reference/scripts):
 *   cfko_docfreq      distance_based_kmer_recruitment.py:39-63   (closed form: n_reads, n_multi per k-mer)
 *   cfko_clouds       read_kmer_cloud.py:18-31                   (set of indexed k-mers wholly inside a unit)
 *   cfko_dist_edges   distance_based_kmer_recruitment.py:85-149  (per source k-mer: sort-reduce of (b, d))
 * Inputs are the flat arrays centroflye_b200.ingest produces (one uint8 code 0..3 per base).
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC -o libcfk_oracle.so cfk_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#else
static int omp_get_thread_num(void) { return 0; }
static int omp_get_max_threads(void) { return 1; }
#endif

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

/* LSD radix sort of u64 keys on their low `bits` bits (tmp has n entries) */
static void radix_sort_u64(uint64_t* a, uint64_t* tmp, int64_t n, int bits) {
  for (int shift = 0; shift < bits; shift += 11) {
    int64_t hist[2048];
    memset(hist, 0, sizeof(hist));
    for (int64_t i = 0; i < n; ++i) hist[(a[i] >> shift) & 2047]++;
    int64_t sum = 0;
    for (int b = 0; b < 2048; ++b) { int64_t c = hist[b]; hist[b] = sum; sum += c; }
    for (int64_t i = 0; i < n; ++i) tmp[hist[(a[i] >> shift) & 2047]++] = a[i];
    uint64_t* t = a; a = tmp; tmp = t;
  }
  /* number of passes */
  int passes = (bits + 10) / 11;
  if (passes & 1) memcpy(tmp, a, (size_t)n * sizeof(uint64_t)); /* result currently in the caller's tmp: copy back */
}

void cfko_free(void* p) { free(p); }
int cfko_max_threads(void) { return omp_get_max_threads(); }

/* ---- stage A ------------------------------------------------------------------------------- */
typedef struct { uint64_t* v; int64_t n, cap; } vec64;
static void vec_push(vec64* v, uint64_t x) {
  if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->v = (uint64_t*)realloc(v->v, (size_t)v->cap * 8); }
  v->v[v->n++] = x;
}

int64_t cfko_docfreq(const uint8_t* codes, const int64_t* read_off, const int64_t* read_len, int64_t n_reads, int k,
                     int threads, uint64_t** out_keys, uint32_t** out_nreads, uint32_t** out_nmulti) {
  if (threads < 1) threads = 1;
  const int P = threads; /* hash partitions = merge parallelism */
  const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  vec64* buckets = (vec64*)calloc((size_t)threads * P, sizeof(vec64));
  /* phase 1: per read, distinct k-mers with a "seen more than once in this read" flag (dbkr.py:50-57) */
#pragma omp parallel num_threads(threads)
  {
    const int tid = omp_get_thread_num();
    uint64_t *buf = NULL, *tmp = NULL;
    int64_t cap = 0;
#pragma omp for schedule(dynamic, 4)
    for (int64_t r = 0; r < n_reads; ++r) {
      const int64_t n = read_len[r] - k + 1;
      if (n <= 0) continue;
      if (n > cap) { cap = n; free(buf); free(tmp); buf = (uint64_t*)malloc((size_t)cap * 8); tmp = (uint64_t*)malloc((size_t)cap * 8); }
      const uint8_t* s = codes + read_off[r];
      uint64_t km = 0;
      for (int i = 0; i < k - 1; ++i) km = (km << 2) | s[i];
      for (int64_t i = 0; i < n; ++i) { km = ((km << 2) | s[i + k - 1]) & mask; buf[i] = km; }
      radix_sort_u64(buf, tmp, n, 2 * k);
      for (int64_t i = 0; i < n;) {
        int64_t j = i + 1;
        while (j < n && buf[j] == buf[i]) ++j;
        vec_push(&buckets[(size_t)tid * P + (mix64(buf[i]) % (uint64_t)P)], (buf[i] << 1) | (uint64_t)(j - i > 1));
        i = j;
      }
    }
    free(buf); free(tmp);
  }
  /* phase 2: per partition, sum over reads (dbkr.py:55-59 in closed form) */
  int64_t* part_n = (int64_t*)calloc(P, sizeof(int64_t));
  uint64_t** pk = (uint64_t**)calloc(P, sizeof(void*));
  uint32_t** pr = (uint32_t**)calloc(P, sizeof(void*));
  uint32_t** pm = (uint32_t**)calloc(P, sizeof(void*));
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int p = 0; p < P; ++p) {
    int64_t total = 0;
    for (int t = 0; t < threads; ++t) total += buckets[(size_t)t * P + p].n;
    int64_t cap = 16;
    while (cap < total * 2) cap <<= 1;
    uint64_t* keys = (uint64_t*)malloc((size_t)cap * 8);
    uint32_t* nr = (uint32_t*)calloc(cap, 4);
    uint32_t* nm = (uint32_t*)calloc(cap, 4);
    memset(keys, 0xFF, (size_t)cap * 8);
    int64_t distinct = 0;
    for (int t = 0; t < threads; ++t) {
      vec64* b = &buckets[(size_t)t * P + p];
      for (int64_t i = 0; i < b->n; ++i) {
        const uint64_t key = b->v[i] >> 1;
        int64_t s = (int64_t)(mix64(key * 0x9E3779B97F4A7C15ull) & (uint64_t)(cap - 1));
        while (keys[s] != key && keys[s] != ~0ull) s = (s + 1) & (cap - 1);
        if (keys[s] == ~0ull) { keys[s] = key; ++distinct; }
        nr[s] += 1;
        nm[s] += (uint32_t)(b->v[i] & 1);
      }
      free(b->v);
    }
    /* compact */
    int64_t w = 0;
    for (int64_t s = 0; s < cap; ++s)
      if (keys[s] != ~0ull) { keys[w] = keys[s]; nr[w] = nr[s]; nm[w] = nm[s]; ++w; }
    part_n[p] = distinct; pk[p] = keys; pr[p] = nr; pm[p] = nm;
  }
  int64_t total = 0;
  for (int p = 0; p < P; ++p) total += part_n[p];
  *out_keys = (uint64_t*)malloc((size_t)(total ? total : 1) * 8);
  *out_nreads = (uint32_t*)malloc((size_t)(total ? total : 1) * 4);
  *out_nmulti = (uint32_t*)malloc((size_t)(total ? total : 1) * 4);
  int64_t off = 0;
  for (int p = 0; p < P; ++p) {
    memcpy(*out_keys + off, pk[p], (size_t)part_n[p] * 8);
    memcpy(*out_nreads + off, pr[p], (size_t)part_n[p] * 4);
    memcpy(*out_nmulti + off, pm[p], (size_t)part_n[p] * 4);
    off += part_n[p];
    free(pk[p]); free(pr[p]); free(pm[p]);
  }
  free(part_n); free(pk); free(pr); free(pm); free(buckets);
  return total;
}

/* ---- stage B ------------------------------------------------------------------------------- */
static int cmp_u32(const void* a, const void* b) {
  uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
  return (x > y) - (x < y);
}

/* unit_cnt[u] = |cloud(u)|; ids of unit u are written, sorted, to tmp_ids[unit_kbase[u] ...] */
void cfko_clouds(const uint8_t* codes, const int64_t* unit_off, const int32_t* unit_len, const int64_t* unit_kbase,
                 int64_t n_units, int k, const uint64_t* rare_sorted, int64_t n_rare, int threads, uint32_t* tmp_ids,
                 int32_t* unit_cnt) {
  if (threads < 1) threads = 1;
  const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  /* probe table over the rare set */
  int64_t cap = 16;
  while (cap < 2 * n_rare) cap <<= 1;
  uint64_t* tk = (uint64_t*)malloc((size_t)cap * 8);
  uint32_t* tv = (uint32_t*)malloc((size_t)cap * 4);
  memset(tk, 0xFF, (size_t)cap * 8);
  for (int64_t i = 0; i < n_rare; ++i) {
    int64_t s = (int64_t)(mix64(rare_sorted[i]) & (uint64_t)(cap - 1));
    while (tk[s] != ~0ull) s = (s + 1) & (cap - 1);
    tk[s] = rare_sorted[i]; tv[s] = (uint32_t)i;
  }
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16)
  for (int64_t u = 0; u < n_units; ++u) {
    const int64_t n = (int64_t)unit_len[u] - k + 1;
    int32_t cnt = 0;
    if (n > 0) {
      const uint8_t* s = codes + unit_off[u];
      uint32_t* out = tmp_ids + unit_kbase[u];
      uint64_t km = 0;
      for (int i = 0; i < k - 1; ++i) km = (km << 2) | s[i];
      for (int64_t i = 0; i < n; ++i) {
        km = ((km << 2) | s[i + k - 1]) & mask;
        int64_t h = (int64_t)(mix64(km) & (uint64_t)(cap - 1));
        while (tk[h] != ~0ull && tk[h] != km) h = (h + 1) & (cap - 1);
        if (tk[h] == km) out[cnt++] = tv[h];
      }
      qsort(out, (size_t)cnt, 4, cmp_u32);
      int32_t w = 0;
      for (int32_t i = 0; i < cnt; ++i)
        if (i == 0 || out[i] != out[i - 1]) out[w++] = out[i];
      cnt = w;
    }
    unit_cnt[u] = cnt;
  }
  free(tk); free(tv);
}

/* ---- stage C + D --------------------------------------------------------------------------- */
/* Sources are visited in the order a_i = (a_first + i * a_step) mod n_kmers for i < n_visit (a_step coprime
 * to n_kmers gives a pseudo-random sample when the time budget stops the loop early).
 * stats: [0] edges, [1] pair increments, [2] sources completed, [3] candidates (cnt >= min_cov). */
int64_t cfko_dist_edges(const int64_t* unit_ptr, const uint32_t* ids, const int32_t* unit_last, int64_t n_kmers,
                        int64_t unit_lo, int64_t unit_hi, int min_d, int max_d, uint32_t min_cov, double rel_threshold,
                        int64_t a_first, int64_t a_step, int64_t n_visit, double time_budget_s, int threads,
                        uint32_t** out_edges, uint8_t* selected, int64_t* stats) {
  if (threads < 1) threads = 1;
  const int dmin = min_d > 1 ? min_d : 1;
  /* occurrence lists: units holding each id, ascending */
  int64_t* occ_ptr = (int64_t*)calloc((size_t)n_kmers + 2, 8);
  for (int64_t e = unit_ptr[unit_lo]; e < unit_ptr[unit_hi]; ++e) occ_ptr[ids[e] + 2]++;
  for (int64_t a = 0; a < n_kmers; ++a) occ_ptr[a + 2] += occ_ptr[a + 1];
  uint32_t* occ = (uint32_t*)malloc((size_t)(occ_ptr[n_kmers + 1] ? occ_ptr[n_kmers + 1] : 1) * 4);
  for (int64_t u = unit_lo; u < unit_hi; ++u)
    for (int64_t e = unit_ptr[u]; e < unit_ptr[u + 1]; ++e) occ[occ_ptr[ids[e] + 1]++] = (uint32_t)u;
  /* occ_ptr[a] .. occ_ptr[a+1] now delimit id a */
  vec64* edge_bufs = (vec64*)calloc(threads, sizeof(vec64)); /* two u64 per edge: (a<<32|b), (d<<32|cnt) */
  int64_t n_incr = 0, n_done = 0, n_cand = 0;
  const double t_start = now_s();
  volatile int stop = 0;
#pragma omp parallel num_threads(threads) reduction(+ : n_incr, n_done, n_cand)
  {
    const int tid = omp_get_thread_num();
    uint64_t *buf = NULL, *tmp = NULL;
    int64_t cap = 0;
#pragma omp for schedule(dynamic, 8)
    for (int64_t i = 0; i < n_visit; ++i) {
      if (stop) continue;
      if (time_budget_s > 0 && (i & 63) == 0 && now_s() - t_start > time_budget_s) { stop = 1; continue; }
      const uint32_t a = (uint32_t)((a_first + (__int128)i * a_step) % n_kmers);
      /* gather (b, d) for every unit pair (g, g + d) with a in g, b in g + d, b != a (dbkr.py:121-127) */
      int64_t need = 0;
      for (int64_t t = occ_ptr[a]; t < occ_ptr[a + 1]; ++t) {
        const int64_t g = occ[t], last = unit_last[g];
        const int64_t hi = g + max_d < last ? g + max_d : last;
        if (g + dmin <= hi) need += unit_ptr[hi + 1] - unit_ptr[g + dmin];
      }
      if (need == 0) { ++n_done; continue; }
      if (need > cap) { cap = need; free(buf); free(tmp); buf = (uint64_t*)malloc((size_t)cap * 8); tmp = (uint64_t*)malloc((size_t)cap * 8); }
      int64_t n = 0;
      for (int64_t t = occ_ptr[a]; t < occ_ptr[a + 1]; ++t) {
        const int64_t g = occ[t], last = unit_last[g];
        const int64_t hi = g + max_d < last ? g + max_d : last;
        for (int64_t u = g + dmin; u <= hi; ++u)
          for (int64_t e = unit_ptr[u]; e < unit_ptr[u + 1]; ++e)
            if (ids[e] != a) buf[n++] = ((uint64_t)ids[e] << 16) | (uint64_t)(u - g);
      }
      n_incr += n;
      int bits = 16;
      while (bits < 48 && (n_kmers >> (bits - 16)) != 0) ++bits;
      radix_sort_u64(buf, tmp, n, bits);
      /* runs of equal (b, d) are the counters; runs of equal b give all_occ (dbkr.py:143) */
      for (int64_t s = 0; s < n;) {
        const uint64_t b = buf[s] >> 16;
        int64_t e_b = s;
        while (e_b < n && (buf[e_b] >> 16) == b) ++e_b;
        const double all_occ = (double)(e_b - s);
        for (int64_t q = s; q < e_b;) {
          int64_t q2 = q + 1;
          while (q2 < e_b && buf[q2] == buf[q]) ++q2;
          const uint64_t cnt = (uint64_t)(q2 - q);
          if (cnt >= min_cov) {
            ++n_cand;
            if ((double)cnt / all_occ >= rel_threshold) {
              vec_push(&edge_bufs[tid], ((uint64_t)a << 32) | b);
              vec_push(&edge_bufs[tid], ((buf[q] & 0xFFFF) << 32) | cnt);
              selected[a] = 1;
              selected[b] = 1;
            }
          }
          q = q2;
        }
        s = e_b;
      }
      ++n_done;
    }
    free(buf); free(tmp);
  }
  int64_t n_edges = 0;
  for (int t = 0; t < threads; ++t) n_edges += edge_bufs[t].n / 2;
  uint32_t* edges = (uint32_t*)malloc((size_t)(n_edges ? n_edges : 1) * 16);
  int64_t w = 0;
  for (int t = 0; t < threads; ++t) {
    for (int64_t i = 0; i + 1 < edge_bufs[t].n; i += 2, ++w) {
      edges[4 * w + 0] = (uint32_t)(edge_bufs[t].v[i] >> 32);
      edges[4 * w + 1] = (uint32_t)edge_bufs[t].v[i];
      edges[4 * w + 2] = (uint32_t)(edge_bufs[t].v[i + 1] >> 32);
      edges[4 * w + 3] = (uint32_t)edge_bufs[t].v[i + 1];
    }
    free(edge_bufs[t].v);
  }
  free(edge_bufs); free(occ); free(occ_ptr);
  *out_edges = edges;
  stats[0] = n_edges; stats[1] = n_incr; stats[2] = n_done; stats[3] = n_cand;
  return n_edges;
}
