// result_writer.cpp — native writer of the recruitment stage's edge file (host side).
//
// The reference writes `unique_edges_min_edge_cov_{mc}.txt` with one Python f-string per edge
// (scripts/distance_based_kmer_recruitment.py:165-171: "{dist} {kmer_i} {kmer_j} {cnt}\n").  At configs[1] scale that is
// 8.8e6 lines / 400 MB and, once the device path takes 26 ms, the slowest step of the CLI.  Here the lines are formatted by
// all host threads into per-thread buffers and written in order; the bytes are identical to the Python writer's
// (tests/test_result_writer.py).  Plain host C++, linked into libcfk.so.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cfk.h"

namespace {

thread_local char g_writer_err[256] = "";

inline char* put_u64(char* p, uint64_t v) {
  char tmp[20];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

inline char* put_i64(char* p, int64_t v) {
  if (v < 0) {
    *p++ = '-';
    return put_u64(p, (uint64_t)0 - (uint64_t)v);
  }
  return put_u64(p, (uint64_t)v);
}

inline char* put_kmer(char* p, uint64_t key, int k) {
  for (int b = k - 1; b >= 0; --b) *p++ = "ACGT"[(key >> (2 * b)) & 3u];
  return p;
}

// One edge as the writer needs it, whatever the caller's arrays look like.
struct Edge {
  int64_t dist, i, j, freq;
};

// Formats the lines of edges [0, n_edges) with all threads, a batch of T slabs at a time, and writes every batch in
// order; batch b + 1 is formatted while batch b is being written (two sets of buffers, one writer thread at a time).
template <class Get>
int write_edge_lines(const char* path, const uint64_t* keys_sorted_h, int64_t n_keys, int32_t k, int64_t n_edges,
                     int32_t n_threads, Get get) {
  FILE* f = fopen(path, "wb");
  if (!f) {
    snprintf(g_writer_err, sizeof(g_writer_err), "cfk_write_edges: cannot open %s", path);
    return CFK_ERR_INVALID;
  }
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency()),
                                                            (n_edges + 65535) / 65536));
  const size_t line_max = 20 + 1 + (size_t)k + 1 + (size_t)k + 1 + 20 + 1;  // two int64, two k-mers, separators
  const int64_t slab = 1 << 16;                                            // edges formatted per buffer
  std::vector<std::vector<char>> bufs[2] = {std::vector<std::vector<char>>((size_t)T), std::vector<std::vector<char>>((size_t)T)};
  std::vector<size_t> used[2] = {std::vector<size_t>((size_t)T, 0), std::vector<size_t>((size_t)T, 0)};
  std::vector<int> bad((size_t)T, 0);
  int rc = CFK_OK, write_rc = CFK_OK;
  std::thread writer;
  int64_t batch = 0;
  for (int64_t base = 0; base < n_edges && rc == CFK_OK; base += slab * T, ++batch) {
    const int cur = (int)(batch & 1);
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t) {
      const int64_t lo = base + (int64_t)t * slab, hi = std::min(n_edges, lo + slab);
      used[cur][(size_t)t] = 0;
      if (lo >= hi) continue;
      pool.emplace_back([&, t, lo, hi, cur] {
        std::vector<char>& b = bufs[cur][(size_t)t];
        b.resize((size_t)(hi - lo) * line_max);
        char* p = b.data();
        for (int64_t e = lo; e < hi; ++e) {
          const Edge x = get(e);
          if (x.i < 0 || x.i >= n_keys || x.j < 0 || x.j >= n_keys) {
            bad[(size_t)t] = 1;
            break;
          }
          p = put_i64(p, x.dist);
          *p++ = ' ';
          p = put_kmer(p, keys_sorted_h[x.i], k);
          *p++ = ' ';
          p = put_kmer(p, keys_sorted_h[x.j], k);
          *p++ = ' ';
          p = put_i64(p, x.freq);
          *p++ = '\n';
        }
        used[cur][(size_t)t] = (size_t)(p - b.data());
      });
    }
    for (auto& th : pool) th.join();
    if (writer.joinable()) writer.join();  // the previous batch is on its way to the file: its buffers are free again
    if (write_rc != CFK_OK) {
      snprintf(g_writer_err, sizeof(g_writer_err), "cfk_write_edges: short write to %s", path);
      rc = write_rc;
    }
    for (int t = 0; t < T && rc == CFK_OK; ++t)
      if (bad[(size_t)t]) {
        snprintf(g_writer_err, sizeof(g_writer_err), "cfk_write_edges: k-mer id outside [0, n_keys)");
        rc = CFK_ERR_INVALID;
      }
    if (rc != CFK_OK) break;
    writer = std::thread([&, cur] {
      for (int t = 0; t < T && write_rc == CFK_OK; ++t)
        if (used[cur][(size_t)t] && fwrite(bufs[cur][(size_t)t].data(), 1, used[cur][(size_t)t], f) != used[cur][(size_t)t])
          write_rc = CFK_ERR_INVALID;
    });
  }
  if (writer.joinable()) writer.join();
  if (write_rc != CFK_OK && rc == CFK_OK) {
    snprintf(g_writer_err, sizeof(g_writer_err), "cfk_write_edges: short write to %s", path);
    rc = write_rc;
  }
  if (fclose(f) != 0 && rc == CFK_OK) {
    snprintf(g_writer_err, sizeof(g_writer_err), "cfk_write_edges: close failed for %s", path);
    rc = CFK_ERR_INVALID;
  }
  return rc;
}

}  // namespace

extern "C" {

const char* cfk_writer_last_error(void) { return g_writer_err; }

int cfk_write_edges(const char* path, const uint64_t* keys_sorted_h, int64_t n_keys, int32_t k, const int64_t* dist_h,
                    const int64_t* i_h, const int64_t* j_h, const int64_t* freq_h, int64_t n_edges, int32_t n_threads) {
  if (!path || n_edges < 0 || n_keys < 0 || k < 1 || k > 31 || (n_edges > 0 && (!dist_h || !i_h || !j_h || !freq_h || !keys_sorted_h))) {
    snprintf(g_writer_err, sizeof(g_writer_err), "cfk_write_edges: bad arguments");
    return CFK_ERR_INVALID;
  }
  return write_edge_lines(path, keys_sorted_h, n_keys, k, n_edges, n_threads,
                          [=](int64_t e) { return Edge{dist_h[e], i_h[e], j_h[e], freq_h[e]}; });
}

int cfk_write_edges_rows(const char* path, const uint64_t* keys_sorted_h, int64_t n_keys, int32_t k, const uint32_t* rows_h,
                         int64_t n_edges, int32_t n_threads) {
  if (!path || n_edges < 0 || n_keys < 0 || k < 1 || k > 31 || (n_edges > 0 && (!rows_h || !keys_sorted_h))) {
    snprintf(g_writer_err, sizeof(g_writer_err), "cfk_write_edges_rows: bad arguments");
    return CFK_ERR_INVALID;
  }
  return write_edge_lines(path, keys_sorted_h, n_keys, k, n_edges, n_threads, [=](int64_t e) {
    const uint32_t* r = rows_h + 4 * e;  // (i, j, dist, freq): the rows cfk_pair_join writes
    return Edge{(int64_t)r[2], (int64_t)r[0], (int64_t)r[1], (int64_t)r[3]};
  });
}

}  // extern "C"
