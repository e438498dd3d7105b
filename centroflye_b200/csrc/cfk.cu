// cfk.cu — sm_100a kernels + C ABI of the unique-k-mer recruitment path (see include/cfk.h).
//
// Everything here is integer / byte work bounded by HBM, L2 and shared-memory
// throughput; there is no dense contraction, so no tensor-core (tcgen05) code.
// Kernel-by-kernel rooflines and the HBM data layout are in DESIGN.md.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo
//             -shared -Xcompiler -fPIC -o libcfk.so cfk.cu
#include <stdlib.h>
#include <string.h>

#include "cfk_common.cuh"

thread_local char cfk_g_err[512] = "";
long long cfk_g_launches = 0;  // kernels enqueued through this library (bench.py reports it)

namespace {

using cfk::EMPTY;
using cfk::FULL;
using cfk::blocks_for;
using cfk::fail;
using cfk::home_slot;
using cfk::mix64;

__device__ __forceinline__ uint32_t base_at(const uint32_t* __restrict__ packed, int64_t pos) {
  return (__ldg(packed + (pos >> 4)) >> ((pos & 15) << 1)) & 3u;
}

// find-or-insert in an open-addressing table of 64-bit keys; returns the slot or -1 if full
__device__ __forceinline__ int64_t table_upsert(uint64_t* keys, int64_t cap, uint64_t key) {
  int64_t slot = home_slot(mix64(key), cap);
  for (int64_t probes = 0; probes < cap; ++probes) {
    uint64_t cur = ((volatile uint64_t*)keys)[slot];
    if (cur == key) return slot;
    if (cur == EMPTY) {
      unsigned long long old = atomicCAS((unsigned long long*)(keys + slot), (unsigned long long)EMPTY,
                                         (unsigned long long)key);
      if (old == EMPTY || old == key) return slot;
    }
    if (++slot == cap) slot = 0;
  }
  return -1;
}

__device__ __forceinline__ int64_t table_find(const uint64_t* __restrict__ keys, int64_t cap, uint64_t key) {
  int64_t slot = home_slot(mix64(key), cap);
  for (int64_t probes = 0; probes < cap; ++probes) {
    uint64_t cur = __ldg(keys + slot);
    if (cur == key) return slot;
    if (cur == EMPTY) return -1;
    if (++slot == cap) slot = 0;
  }
  return -1;
}

// ============================================================================================
// Stage A: document frequency.
//
// Global table: open addressing over 16-byte slots { u64 key ; u32 n_reads ; u32 n_multi } (one
// 32-byte DRAM sector holds two slots, so a key probe and its counter update touch one sector).
// Per-read de-duplication happens in SHARED memory, so the global table sees one update per
// distinct (read, k-mer) instead of one per occurrence: one persistent block per read (longest
// reads first) keeps a set of the read's k-mers as 32-bit slots holding the POSITION of the
// k-mer's first occurrence (+1) and a "seen again" bit; equality of two k-mers is checked by
// re-extracting the stored position's k-mer from the packed read.  Reads whose k-mers do not fit
// the set in one go are processed in several passes over disjoint hash partitions of the k-mer
// space.  The first sighting of a k-mer in a read adds 1 to n_reads, the second adds 1 to n_multi
// (exactly once: the thread that flips the "seen again" bit).
// ============================================================================================
constexpr int DF_THREADS = 1024;
constexpr int DF_PER_THREAD = 8;
constexpr int DF_TILE = DF_THREADS * DF_PER_THREAD;  // k-mer starts staged per tile
constexpr int DF_TILE_WORDS = DF_TILE / 16 + 4;       // + up to 30 bases of overlap
constexpr int DF_SET_SLOTS = CFK_DOCFREQ_SET_SLOTS;   // u32 slots of the per-read set
constexpr int DF_SET_FILL = DF_SET_SLOTS * 53 / 100;  // k-mers planned per pass
constexpr uint32_t DF_MULTI = 0x80000000u;

// find-or-insert in the global table, probing linearly from `slot` (may equal cap: wraps); returns the slot or -1 if full
__device__ __forceinline__ int64_t slot_upsert(uint64_t* table, int64_t cap, uint64_t key, int64_t slot) {
  if (slot >= cap) slot = 0;
  for (int64_t probes = 0; probes < cap; ++probes) {
    const uint64_t cur = ((volatile uint64_t*)table)[2 * slot];
    if (cur == key) return slot;
    if (cur == EMPTY) {
      const unsigned long long old = atomicCAS((unsigned long long*)(table + 2 * slot), (unsigned long long)EMPTY,
                                               (unsigned long long)key);
      if (old == EMPTY || old == key) return slot;
    }
    if (++slot == cap) slot = 0;
  }
  return -1;
}

// k-mer starting at base q of a packed read (words = word holding the read's base 0)
__device__ __forceinline__ uint64_t kmer_at(const uint32_t* __restrict__ words, uint32_t q, int k) {
  const uint32_t w = q >> 4, sh = (q & 15u) << 1;
  uint64_t bits = ((uint64_t)__ldg(words + w) | ((uint64_t)__ldg(words + w + 1) << 32)) >> sh;
  if (sh) bits |= (uint64_t)__ldg(words + w + 2) << (64 - sh);
  // base q + j sits at bits 2j..2j+1; the k-mer wants base q in its top bits: reverse the 2-bit groups
  uint64_t r = __brevll(bits);
  r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);
  return r >> (64 - 2 * k);
}

__global__ void __launch_bounds__(DF_THREADS, 1)
docfreq_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ read_off,
               const int64_t* __restrict__ read_len, const int32_t* __restrict__ order, int64_t n_reads, int k,
               uint64_t* table, int64_t cap, int64_t* counters) {
  extern __shared__ __align__(16) uint32_t df_smem[];
  uint32_t* set = df_smem;                   // [DF_SET_SLOTS]
  uint32_t* s_words = df_smem + DF_SET_SLOTS;  // [DF_TILE_WORDS]
  __shared__ long long s_item;
  const uint64_t mask = (1ull << (2 * k)) - 1;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = (long long)atomicAdd((unsigned long long*)(counters + 2), 1ull);
    __syncthreads();
    const int64_t item = s_item;
    if (item >= n_reads) break;
    const int64_t r = order[item];
    const int64_t nk = read_len[r] - k + 1;
    if (nk <= 0) continue;
    const uint32_t* words = packed + (read_off[r] >> 4);  // every read starts on a 64-base boundary
    const uint32_t n_pass = (uint32_t)((nk + DF_SET_FILL - 1) / DF_SET_FILL);
    const uint32_t c_eff = (n_pass > 1) ? (uint32_t)DF_SET_SLOTS
                                        : (uint32_t)min((int64_t)DF_SET_SLOTS, max((int64_t)2048, 2 * nk));
    for (uint32_t pass = 0; pass < n_pass; ++pass) {
      __syncthreads();
      for (uint32_t i = threadIdx.x * 4; i < c_eff; i += DF_THREADS * 4)  // c_eff is a multiple of 4 or the full set
        *reinterpret_cast<uint4*>(set + i) = make_uint4(0, 0, 0, 0);
      for (int64_t tile0 = 0; tile0 < nk; tile0 += DF_TILE) {
        const int npos = (int)min((int64_t)DF_TILE, nk - tile0);
        const int nwords = (npos + k - 1 + 15) >> 4;
        __syncthreads();  // set cleared / previous tile consumed
        for (int i = threadIdx.x; i < nwords; i += DF_THREADS) s_words[i] = __ldg(words + (tile0 >> 4) + i);
        __syncthreads();
        const int p0 = threadIdx.x * DF_PER_THREAD;
        if (p0 >= npos) continue;
        uint64_t kmer = 0;
        for (int i = 0; i < k - 1; ++i) {
          const int p = p0 + i;
          kmer = (kmer << 2) | ((s_words[p >> 4] >> ((p & 15) << 1)) & 3u);
        }
        // phase 1: the read's own set (shared memory).  act holds 2 bits per k-mer of this thread:
        // 1 = first sighting in this read (n_reads += 1), 2 = second sighting (n_multi += 1).
        uint64_t km[DF_PER_THREAD];
        uint32_t act = 0;
#pragma unroll
        for (int j = 0; j < DF_PER_THREAD; ++j) {
          const int p = p0 + j;
          km[j] = 0;
          if (p >= npos) continue;
          const int q = p + k - 1;
          kmer = ((kmer << 2) | ((s_words[q >> 4] >> ((q & 15) << 1)) & 3u)) & mask;
          km[j] = kmer;
          if (n_pass > 1 && __umulhi((uint32_t)(kmer ^ (kmer >> 32)) * 0x9E3779B1u, n_pass) != pass) continue;
          const uint32_t pos = (uint32_t)(tile0 + p);
          uint32_t s = __umulhi((uint32_t)mix64(kmer), c_eff);
          uint32_t probes = 0;
          for (; probes < c_eff; ++probes) {
            uint32_t v = ((volatile uint32_t*)set)[s];
            if (v == 0) {
              v = atomicCAS(set + s, 0u, pos + 1);
              if (v == 0) {  // first sighting of this k-mer in this read
                act |= 1u << (2 * j);
                break;
              }
            }
            if (kmer_at(words, (v & ~DF_MULTI) - 1, k) == kmer) {
              if (!(v & DF_MULTI) && !(atomicOr(set + s, DF_MULTI) & DF_MULTI)) act |= 2u << (2 * j);  // exactly one thread flips the bit
              break;
            }
            if (++s == c_eff) s = 0;
          }
          if (probes == c_eff) counters[1] = 1;  // cannot happen: a pass is planned for <= 53 % load
        }
        if (act == 0) continue;
        // phase 2: the global table, four independent claims in flight per thread (the table is far larger than L2:
        // every probe is a DRAM round trip, so the round trips must overlap).  The claim itself is the probe:
        // CAS(EMPTY -> key) returns EMPTY (claimed) or the resident key.
#pragma unroll
        for (int half = 0; half < DF_PER_THREAD / 4; ++half) {
          if (((act >> (8 * half)) & 0xFFu) == 0) continue;
          int64_t slot[4];
          uint64_t old[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int jj = half * 4 + j;
            slot[j] = home_slot(mix64(km[jj]), cap);
            old[j] = km[jj];
            if ((act >> (2 * jj)) & 3u)
              old[j] = atomicCAS((unsigned long long*)(table + 2 * slot[j]), (unsigned long long)EMPTY,
                                 (unsigned long long)km[jj]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int jj = half * 4 + j;
            const uint32_t what = (act >> (2 * jj)) & 3u;
            if (!what) continue;
            int64_t sl = slot[j];
            if (old[j] != EMPTY && old[j] != km[jj]) sl = slot_upsert(table, cap, km[jj], sl + 1);
            if (sl < 0) counters[0] = 1;
            else atomicAdd(reinterpret_cast<uint32_t*>(table + 2 * sl + 1) + (what >> 1), 1u);
          }
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// Stage A, resident form (the default): the work item is one (read, pass) instead of one read,
// and the item's whole packed read is staged in shared memory next to the set.
//   * every k-mer is rolled out of shared memory and every "is the k-mer stored at this slot
//     mine?" check re-extracts the stored position's k-mer from shared memory, so phase 1 never
//     leaves the SM and -- with no tile to restage -- the position loop has no barrier at all;
//   * the passes of a long read (hash partitions of its k-mer space) are independent items, so
//     a 300 kb read no longer serialises a dozen passes on one SM at the end of the kernel;
//   * the shared memory is split per item: [ read words | set ], the set gets whatever the read
//     leaves.  docfreq_plan_kernel computes the number of passes of every read from the same
//     rule; the exclusive scan of those numbers (item_ptr) maps a ticket to (read, pass).
// Reads too long to leave DF3_MIN_SET slots keep their words in global memory (same code, other
// address space).
// --------------------------------------------------------------------------------------------
#ifndef CFK_DF3_THREADS
#define CFK_DF3_THREADS 1024  /* threads per block; 1024 / this many blocks share an SM and its shared memory */
#endif
constexpr int DF3_THREADS = CFK_DF3_THREADS;
constexpr int DF3_BLOCKS_PER_SM = 1024 / DF3_THREADS;
static_assert(DF3_THREADS == 1024 || DF3_THREADS == 512 || DF3_THREADS == 256, "1, 2 or 4 blocks per SM");
constexpr int DF3_SMEM_WORDS = DF3_BLOCKS_PER_SM == 1 ? 57344 : 57344 / DF3_BLOCKS_PER_SM - 512;  // dynamic shared memory of a block
constexpr int DF3_MIN_SET = 16384 / DF3_BLOCKS_PER_SM;
#ifndef CFK_DF3_FILL_PCT
#define CFK_DF3_FILL_PCT 53   /* planned load of the per-read set, percent */
#endif
#ifndef CFK_DF3_PER_THREAD
#define CFK_DF3_PER_THREAD 4  /* consecutive k-mer starts per lane and grab (4 or 8) */
#endif
constexpr int DF3_PER_THREAD = CFK_DF3_PER_THREAD;
static_assert(DF3_PER_THREAD == 4 || DF3_PER_THREAD == 8, "the window holds 8 + 7 + 30 bases; claims go in groups of 4");

struct Df3Geometry {
  uint32_t n_words;   // words staged (0: the read stays in global memory)
  uint32_t set_slots;
  uint32_t fill;      // k-mers planned per pass
};

__host__ __device__ __forceinline__ Df3Geometry df3_geometry(int64_t len) {
  Df3Geometry g;
  const int64_t nw = (((len + 15) >> 4) + 3 + 3) & ~(int64_t)3;  // + the 3-word extraction window, multiple of 4
  g.n_words = (nw <= DF3_SMEM_WORDS - DF3_MIN_SET) ? (uint32_t)nw : 0u;
  g.set_slots = (uint32_t)DF3_SMEM_WORDS - g.n_words;
  g.fill = (uint32_t)((uint64_t)g.set_slots * CFK_DF3_FILL_PCT / 100);
  return g;
}

__global__ void docfreq_plan_kernel(const int64_t* __restrict__ read_len, const int32_t* __restrict__ order, int64_t n_reads,
                                    int k, int32_t* __restrict__ n_pass) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_reads) return;
  const int64_t len = read_len[order[i]], nk = len - k + 1;
  int32_t np = 0;
  if (nk > 0) {
    const Df3Geometry g = df3_geometry(len);
    np = (int32_t)((nk + g.fill - 1) / g.fill);
  }
  n_pass[i] = np;
}

template <bool IN_SMEM>
__device__ __forceinline__ uint32_t df3_word(const uint32_t* words, uint32_t i) {
  if (IN_SMEM) return words[i];
  return __ldg(words + i);
}

// k-mer starting at base q; words[] must be readable up to word (q >> 4) + 2
template <bool IN_SMEM>
__device__ __forceinline__ uint64_t df3_kmer_at(const uint32_t* words, uint32_t q, int k) {
  const uint32_t w = q >> 4, sh = (q & 15u) << 1;
  uint64_t bits = ((uint64_t)df3_word<IN_SMEM>(words, w) | ((uint64_t)df3_word<IN_SMEM>(words, w + 1) << 32)) >> sh;
  if (sh) bits |= (uint64_t)df3_word<IN_SMEM>(words, w + 2) << (64 - sh);
  uint64_t r = __brevll(bits);
  r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);
  return r >> (64 - 2 * k);
}

// hash of a k-mer for the per-read set: the pass partition takes the top bits of f * golden, the slot and the
// fingerprint come from a second, independent mix of the same folded value
__device__ __forceinline__ uint32_t df3_fold(uint64_t kmer) { return (uint32_t)(kmer ^ (kmer >> 32)); }
__device__ __forceinline__ uint32_t df3_mix(uint32_t f) {
  uint32_t g = f * 0x85EBCA6Bu;
  g ^= g >> 13;
  g *= 0xC2B2AE35u;
  g ^= g >> 16;
  return g;
}

constexpr int DF3_CHUNK = 32 * DF3_PER_THREAD;  // k-mer starts one warp takes per grab

// Set slot (32 bit): bit 31 = "seen again", bits [pos_bits - 1 : 0] = position of the first occurrence + 1 (0 = empty),
// bits [30 : pos_bits] = fingerprint of the k-mer, so that most foreign slots of a probe chain are skipped without
// re-extracting their k-mer.
template <bool IN_SMEM>
__device__ __forceinline__ void df3_item(const uint32_t* words, uint32_t* set, uint32_t c_eff, int64_t nk, int k,
                                         uint32_t pass, uint32_t n_pass, uint64_t* table, int64_t cap, int64_t* counters,
                                         uint32_t* s_chunk) {
  const uint64_t mask = (1ull << (2 * k)) - 1;
  const int lane = threadIdx.x & 31;
  const int pos_bits = 64 - __clzll((unsigned long long)nk);  // pos + 1 <= nk fits
  const uint32_t pos_mask = pos_bits >= 31 ? 0x7FFFFFFFu : ((1u << pos_bits) - 1u);
  const uint32_t fp_mask = 0x7FFFFFFFu & ~pos_mask;
  for (;;) {
    uint32_t chunk = 0;
    if (lane == 0) chunk = atomicAdd(s_chunk, 1u);
    chunk = __shfl_sync(FULL, chunk, 0);
    const int64_t chunk0 = (int64_t)chunk * DF3_CHUNK;
    if (chunk0 >= nk) break;
    const int64_t base = chunk0 + lane * DF3_PER_THREAD;
    if (base >= nk) continue;
    const uint32_t p0 = (uint32_t)base;  // multiple of 4: sits at offset 0, 4, 8 or 12 of its word
    const int npos = (int)min((int64_t)DF3_PER_THREAD, nk - base);
    // 48-base window starting at the word of p0; offset + DF3_PER_THREAD - 1 + k - 1 <= 45 < 48 (12 + 3 + 30 or 8 + 7 + 30)
    const uint32_t w0 = p0 >> 4;
    uint64_t win_lo = (uint64_t)df3_word<IN_SMEM>(words, w0) | ((uint64_t)df3_word<IN_SMEM>(words, w0 + 1) << 32);
    uint32_t win_hi = df3_word<IN_SMEM>(words, w0 + 2);
    if (const uint32_t sh0 = (p0 & 15u) << 1) {
      win_lo = (win_lo >> sh0) | ((uint64_t)win_hi << (64 - sh0));
      win_hi >>= sh0;
    }
    // the first k - 1 bases in one go (reverse the 2-bit groups of the window), then roll
    uint64_t kmer = 0;
    const int s0 = 2 * (k - 1);
    if (s0) {
      uint64_t r = __brevll(win_lo);
      r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);
      kmer = r >> (64 - s0);
      win_lo = (win_lo >> s0) | ((uint64_t)win_hi << (64 - s0));
      win_hi = s0 < 32 ? (win_hi >> s0) : 0u;
    }
    uint64_t km[DF3_PER_THREAD];
    uint32_t act = 0;
#pragma unroll
    for (int j = 0; j < DF3_PER_THREAD; ++j) {
      kmer = ((kmer << 2) | (win_lo & 3u)) & mask;
      win_lo = (win_lo >> 2) | ((uint64_t)win_hi << 62);
      win_hi >>= 2;
      km[j] = kmer;
      if (j >= npos) continue;
      const uint32_t f = df3_fold(kmer);
      if (n_pass > 1 && __umulhi(f * 0x9E3779B1u, n_pass) != pass) continue;
      const uint32_t g = df3_mix(f);
      const uint32_t mine = (g << pos_bits) & fp_mask;  // pos_bits <= 31
      const uint32_t fresh = (p0 + (uint32_t)j + 1u) | mine;
      uint32_t s = __umulhi(g, c_eff);
      uint32_t probes = 0;
      for (; probes < c_eff; ++probes) {
        uint32_t v = ((volatile uint32_t*)set)[s];
        if (v == 0) {
          v = atomicCAS(set + s, 0u, fresh);
          if (v == 0) {  // first sighting of this k-mer in this read
            act |= 1u << (2 * j);
            break;
          }
        }
        if (((v ^ mine) & fp_mask) == 0 && df3_kmer_at<IN_SMEM>(words, (v & pos_mask) - 1u, k) == kmer) {
          if (!(v & DF_MULTI) && !(atomicOr(set + s, DF_MULTI) & DF_MULTI)) act |= 2u << (2 * j);  // exactly one thread flips the bit
          break;
        }
        if (++s == c_eff) s = 0;
      }
      if (probes == c_eff) counters[1] = 1;  // cannot happen: a pass is planned for <= 53 % load
    }
    if (act == 0) continue;
    // phase 2: the global table, four independent claims in flight per thread (see docfreq_kernel)
#pragma unroll
    for (int half = 0; half < DF3_PER_THREAD / 4; ++half) {
      if (((act >> (8 * half)) & 0xFFu) == 0) continue;
      int64_t slot[4];
      uint64_t old[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jj = half * 4 + j;
        slot[j] = home_slot(mix64(km[jj]), cap);
        old[j] = km[jj];
        if ((act >> (2 * jj)) & 3u)
          old[j] = atomicCAS((unsigned long long*)(table + 2 * slot[j]), (unsigned long long)EMPTY,
                             (unsigned long long)km[jj]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jj = half * 4 + j;
        const uint32_t what = (act >> (2 * jj)) & 3u;
        if (!what) continue;
        int64_t sl = slot[j];
        if (old[j] != EMPTY && old[j] != km[jj]) sl = slot_upsert(table, cap, km[jj], sl + 1);
        if (sl < 0) counters[0] = 1;
        else atomicAdd(reinterpret_cast<uint32_t*>(table + 2 * sl + 1) + (what >> 1), 1u);
      }
    }
  }
}

// ticket -> (index into order[], pass, passes of that read); index -1 when the items are used up
__device__ __forceinline__ void df3_fetch(const int64_t* __restrict__ item_ptr, int64_t n_reads, int64_t n_items,
                                          int64_t* counters, long long* s_read, uint32_t* s_pass, uint32_t* s_npass) {
  const int64_t t = (int64_t)atomicAdd((unsigned long long*)(counters + 2), 1ull);
  long long idx = -1;
  if (t < n_items) {
    int64_t lo = 0, hi = n_reads;  // item_ptr[lo] <= t < item_ptr[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(item_ptr + mid) > t) hi = mid; else lo = mid;
    }
    idx = lo;
    const int64_t first = __ldg(item_ptr + lo);
    *s_pass = (uint32_t)(t - first);
    *s_npass = (uint32_t)(__ldg(item_ptr + lo + 1) - first);
  }
  *s_read = idx;
}

__global__ void __launch_bounds__(DF3_THREADS, DF3_BLOCKS_PER_SM)
docfreq_resident_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ read_off,
                        const int64_t* __restrict__ read_len, const int32_t* __restrict__ order,
                        const int64_t* __restrict__ item_ptr, int64_t n_reads, int k, uint64_t* table, int64_t cap,
                        int64_t* counters) {
  extern __shared__ __align__(16) uint32_t df_smem[];
  __shared__ long long s_read[2];
  __shared__ uint32_t s_pass[2], s_npass[2], s_chunk;
  const int64_t n_items = __ldg(item_ptr + n_reads);
  if (threadIdx.x == 0) df3_fetch(item_ptr, n_reads, n_items, counters, &s_read[0], &s_pass[0], &s_npass[0]);
  for (int cur = 0;; cur ^= 1) {
    __syncthreads();  // item `cur` is published; everybody is done with the previous item's shared memory
    const int64_t idx = s_read[cur];
    if (idx < 0) break;
    const uint32_t pass = s_pass[cur], n_pass = s_npass[cur];
    const int64_t r = order[idx];
    const int64_t len = read_len[r], nk = len - k + 1;
    const uint32_t* gwords = packed + (read_off[r] >> 4);  // every read starts on a 64-base boundary
    const Df3Geometry g = df3_geometry(len);
    uint32_t* set = df_smem + g.n_words;
    const uint32_t c_eff = (n_pass > 1) ? g.set_slots : (uint32_t)min((int64_t)g.set_slots, max((int64_t)2048, (nk * 100 / CFK_DF3_FILL_PCT + 7) & ~(int64_t)3));
    for (uint32_t i = threadIdx.x * 4; i < c_eff; i += DF3_THREADS * 4)  // n_words and c_eff are multiples of 4
      *reinterpret_cast<uint4*>(set + i) = make_uint4(0, 0, 0, 0);
    if (g.n_words) {
      const uint32_t real_words = (uint32_t)((len + 15) >> 4);  // 16-byte loads stay inside the read's own 64-base blocks
      const uint32_t real_quads = (real_words + 3) >> 2;
      for (uint32_t i = threadIdx.x; i < g.n_words / 4; i += DF3_THREADS) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (i < real_quads) v = __ldg(reinterpret_cast<const uint4*>(gwords) + i);
        *reinterpret_cast<uint4*>(df_smem + 4 * i) = v;
      }
    }
    if (threadIdx.x == 0) s_chunk = 0;
    __syncthreads();
    // the next item's ticket and its search run behind the other warps' work on this one
    if (threadIdx.x == 0) df3_fetch(item_ptr, n_reads, n_items, counters, &s_read[cur ^ 1], &s_pass[cur ^ 1], &s_npass[cur ^ 1]);
    if (g.n_words) df3_item<true>(df_smem, set, c_eff, nk, k, pass, n_pass, table, cap, counters, &s_chunk);
    else df3_item<false>(gwords, set, c_eff, nk, k, pass, n_pass, table, cap, counters, &s_chunk);
  }
}

// --------------------------------------------------------------------------------------------
// Total-occurrence count: every k-mer occurrence of every read adds 1 -- no per-read
// de-duplication -- into the same 16-byte-slot table (the n_reads field holds the count, n_multi stays 0).
// Replaces get_kmer_counts_reads, scripts/better_consensus_unit_reconstruction.py:127-135.
// One block per tile of KC_TILE consecutive k-mer starts of one read (tile_read / tile_start from the host);
// a thread rolls 8 consecutive k-mers out of three packed words.
// --------------------------------------------------------------------------------------------
constexpr int KC_THREADS = 256;
constexpr int KC_PER_THREAD = 8;
constexpr int KC_TILE = KC_THREADS * KC_PER_THREAD;

// reverse complement of a k-mer key (first base most significant; A=0 C=1 G=2 T=3, so the complement is 3 - code)
__device__ __forceinline__ uint64_t kmer_revcomp(uint64_t kmer, int k) {
  uint64_t r = __brevll(~kmer);  // complement, then reverse the bit order ...
  r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);  // ... and restore the order inside each base
  return r >> (64 - 2 * k);
}

template <bool CANONICAL>
__global__ void __launch_bounds__(KC_THREADS)
kmer_count_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ read_off, const int64_t* __restrict__ read_len,
                  const int32_t* __restrict__ tile_read, const int64_t* __restrict__ tile_start, int k, uint64_t* table,
                  int64_t cap, int64_t* counters) {
  const int64_t r = tile_read[blockIdx.x];
  const int64_t nk = read_len[r] - k + 1;
  const int64_t base = tile_start[blockIdx.x] + (int64_t)threadIdx.x * KC_PER_THREAD;
  if (base >= nk) return;
  const uint32_t* words = packed + (read_off[r] >> 4);  // every read starts on a 64-base boundary
  const uint64_t mask = (1ull << (2 * k)) - 1;
  const int npos = (int)min((int64_t)KC_PER_THREAD, nk - base);
  // 48-base window from the word of `base` (a multiple of 8): offset + 7 + k - 1 <= 8 + 7 + 30 < 48
  const int64_t w0 = base >> 4;
  uint64_t win_lo = (uint64_t)__ldg(words + w0) | ((uint64_t)__ldg(words + w0 + 1) << 32);
  uint32_t win_hi = __ldg(words + w0 + 2);
  if (base & 8) {
    win_lo = (win_lo >> 16) | ((uint64_t)win_hi << 48);
    win_hi >>= 16;
  }
  uint64_t kmer = 0;
  for (int i = 0; i < k - 1; ++i) {
    kmer = (kmer << 2) | (win_lo & 3u);
    win_lo = (win_lo >> 2) | ((uint64_t)win_hi << 62);
    win_hi >>= 2;
  }
  for (int j = 0; j < npos; ++j) {
    kmer = ((kmer << 2) | (win_lo & 3u)) & mask;
    win_lo = (win_lo >> 2) | ((uint64_t)win_hi << 62);
    win_hi >>= 2;
    uint64_t key = kmer;
    if (CANONICAL) key = min(kmer, kmer_revcomp(kmer, k));  // the smaller of the two strands, like `jellyfish count -C`
    const int64_t slot = slot_upsert(table, cap, key, home_slot(mix64(key), cap));
    if (slot < 0) counters[0] = 1;
    else atomicAdd(reinterpret_cast<uint32_t*>(table + 2 * slot + 1), 1u);
  }
}

// ---- per-read repetitive k-mers: scripts/unit_extractor.py:23-40 -----------------------
// get_repetitive_kmers groups the positions of every k-mer of ONE sequence; get_convolution takes the gaps between
// neighbouring positions of a k-mer.  Here: one sort key (k-mer << pos_bits | position) per k-mer start, sorted with
// cfk_sort_u64; equal k-mers are then neighbours in position order and a gap is the difference of two neighbours.
__global__ void kmer_position_keys_kernel(const uint32_t* __restrict__ words, int64_t n_kmers, int k, int pos_bits,
                                          uint64_t* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_kmers) return;
  keys[i] = (kmer_at(words, (uint32_t)i, k) << pos_bits) | (uint64_t)i;
}

// gaps[i] = position(i) - position(i - 1) when keys i - 1 and i hold the same k-mer, else 0 (also for i = 0)
__global__ void adjacent_gaps_kernel(const uint64_t* __restrict__ keys, int64_t n, int pos_bits, uint32_t* __restrict__ gaps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t g = 0;
  if (i > 0) {
    const uint64_t a = keys[i - 1], b = keys[i];
    if ((a >> pos_bits) == (b >> pos_bits)) g = (uint32_t)((b - a) & ((1ull << pos_bits) - 1ull));
  }
  gaps[i] = g;
}

__global__ void table_init_kernel(uint64_t* table, int64_t cap) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) reinterpret_cast<ulonglong2*>(table)[i] = make_ulonglong2(EMPTY, 0ull);
}

__global__ void table_merge_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ nreads,
                                   const uint32_t* __restrict__ nmulti, int64_t n, uint64_t* table, int64_t cap,
                                   int64_t* counters) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = keys[i];
  int64_t slot = slot_upsert(table, cap, key, home_slot(mix64(key), cap));
  if (slot < 0) { counters[0] = 1; return; }
  uint32_t* cnt = reinterpret_cast<uint32_t*>(table + 2 * slot + 1);
  if (nreads[i]) atomicAdd(cnt, nreads[i]);
  if (nmulti[i]) atomicAdd(cnt + 1, nmulti[i]);
}

// counts of n given keys (0 / 0 for a key the table does not hold)
__global__ void table_lookup_kernel(const uint64_t* __restrict__ table, int64_t cap, const uint64_t* __restrict__ keys, int64_t n,
                                    uint32_t* __restrict__ out_nreads, uint32_t* __restrict__ out_nmulti) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = keys[i];
  int64_t slot = home_slot(mix64(key), cap);
  uint32_t nr = 0, nm = 0;
  for (int64_t probes = 0; probes < cap; ++probes) {
    const ulonglong2 sl = __ldg(reinterpret_cast<const ulonglong2*>(table) + slot);
    if (sl.x == key) {
      nr = (uint32_t)sl.y;
      nm = (uint32_t)(sl.y >> 32);
      break;
    }
    if (sl.x == EMPTY) break;
    if (++slot == cap) slot = 0;
  }
  out_nreads[i] = nr;
  out_nmulti[i] = nm;
}

// warp-aggregated append: returns the output position of this lane's item, or -1
__device__ __forceinline__ int64_t warp_append(bool take, int64_t* counter) {
  unsigned m = __ballot_sync(FULL, take);
  if (m == 0) return -1;
  int lane = threadIdx.x & 31;
  int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd((unsigned long long*)counter, (unsigned long long)__popc(m));
  base = __shfl_sync(FULL, base, leader);
  return take ? (int64_t)base + __popc(m & ((1u << lane) - 1)) : -1;
}

__device__ __forceinline__ int32_t key_owner(uint64_t key, int32_t n_parts) {
  return (int32_t)(mix64(key ^ 0x9E3779B97F4A7C15ull) % (uint64_t)n_parts);
}

// One block = SEL_THREADS x SEL_ITEMS consecutive slots and ONE atomicAdd on the output cursor (a warp-level append costs
// an atomic per warp with a match: 3 million serialised atomics on one address when 1 % of the slots match).
constexpr int SEL_THREADS = 256;
#ifndef CFK_SEL_ITEMS
#define CFK_SEL_ITEMS 2
#endif
constexpr int SEL_ITEMS = CFK_SEL_ITEMS;

__global__ void __launch_bounds__(SEL_THREADS)
table_select_kernel(const uint64_t* __restrict__ table, int64_t cap, uint32_t lo, uint32_t hi, uint32_t max_nonuniq,
                    int32_t n_parts, int32_t part, uint64_t* out_keys, uint32_t* out_nreads, uint32_t* out_nmulti,
                    int64_t max_out, int64_t* counters) {
  __shared__ int s_warp[SEL_THREADS / 32];
  __shared__ long long s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t first = (int64_t)blockIdx.x * (SEL_THREADS * SEL_ITEMS) + threadIdx.x;
  ulonglong2 slot[SEL_ITEMS];
  uint32_t takes = 0;
#pragma unroll
  for (int it = 0; it < SEL_ITEMS; ++it) {
    const int64_t i = first + (int64_t)it * SEL_THREADS;
    slot[it] = make_ulonglong2(EMPTY, 0ull);
    if (i < cap) slot[it] = __ldg(reinterpret_cast<const ulonglong2*>(table) + i);
    const uint64_t key = slot[it].x;
    if (key != EMPTY) {
      const uint32_t nr = (uint32_t)slot[it].y, nm = (uint32_t)(slot[it].y >> 32);
      bool take = nm <= max_nonuniq && nr >= lo && nr <= hi;
      if (take && n_parts > 0) take = key_owner(key, n_parts) == part;
      takes |= (uint32_t)take << it;
    }
  }
  // block-exclusive scan of the per-thread match counts
  const int mine = __popc(takes);
  int incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int w = 0; w < SEL_THREADS / 32; ++w) {
      const int c = s_warp[w];
      s_warp[w] = total;
      total += c;
    }
    s_base = total ? (long long)atomicAdd((unsigned long long*)counters, (unsigned long long)total) : 0ll;
  }
  __syncthreads();
  int64_t pos = (int64_t)s_base + s_warp[warp] + incl - mine;
#pragma unroll
  for (int it = 0; it < SEL_ITEMS; ++it) {
    if (!((takes >> it) & 1u)) continue;
    if (pos < max_out) {
      if (out_keys) out_keys[pos] = slot[it].x;
      if (out_nreads) out_nreads[pos] = (uint32_t)slot[it].y;
      if (out_nmulti) out_nmulti[pos] = (uint32_t)(slot[it].y >> 32);
    }
    ++pos;
  }
}

// ---- hash partition of a table for the multi-GPU exchange --------------------------------------
// owner(key) = mix64(key ^ golden) % n_parts -- the same rule table_select_kernel applies.
constexpr int TP_MAX_PARTS = 64;

__global__ void __launch_bounds__(256) table_part_count_kernel(const uint64_t* __restrict__ table, int64_t cap,
                                                              int32_t n_parts, int64_t* counts) {
  __shared__ int s_cnt[TP_MAX_PARTS];
  if (threadIdx.x < TP_MAX_PARTS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = table[2 * i];
    if (key != EMPTY) atomicAdd(&s_cnt[key_owner(key, n_parts)], 1);
  }
  __syncthreads();
  if (threadIdx.x < n_parts && s_cnt[threadIdx.x])
    atomicAdd((unsigned long long*)(counts + threadIdx.x), (unsigned long long)s_cnt[threadIdx.x]);
}

// cursors[p] starts at the first output index of partition p; records of one partition end up
// contiguous (in arbitrary order), which is the send layout of the all-to-all.
__global__ void __launch_bounds__(256) table_part_scatter_kernel(const uint64_t* __restrict__ table, int64_t cap,
                                                                int32_t n_parts, int64_t* cursors, uint64_t* out_keys,
                                                                uint32_t* out_nreads, uint32_t* out_nmulti) {
  const int lane = threadIdx.x & 31;
  const int64_t n_iter = (cap + (int64_t)gridDim.x * blockDim.x - 1) / ((int64_t)gridDim.x * blockDim.x);
  for (int64_t it = 0; it < n_iter; ++it) {  // every lane runs every iteration (warp collectives below)
    const int64_t i = (it * gridDim.x + blockIdx.x) * (int64_t)blockDim.x + threadIdx.x;
    ulonglong2 s = make_ulonglong2(EMPTY, 0ull);
    if (i < cap) s = reinterpret_cast<const ulonglong2*>(table)[i];
    const bool valid = s.x != EMPTY;
    const int32_t p = valid ? key_owner(s.x, n_parts) : -1;
    const unsigned peers = __match_any_sync(FULL, p);
    const int leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (valid && lane == leader) base = atomicAdd((unsigned long long*)(cursors + p), (unsigned long long)__popc(peers));
    base = __shfl_sync(FULL, base, leader);
    if (valid) {
      const int64_t pos = (int64_t)base + __popc(peers & ((1u << lane) - 1));
      out_keys[pos] = s.x;
      out_nreads[pos] = (uint32_t)s.y;
      out_nmulti[pos] = (uint32_t)(s.y >> 32);
    }
  }
}

// ============================================================================================
// Bitonic sort of u64 keys ("flip / disperse" form: every compare-exchange is ascending, so
// indices >= n behave as +inf padding without being stored).
// ============================================================================================
constexpr int SORT_TILE = 4096;
constexpr int SORT_THREADS = 512;

__device__ __forceinline__ void cmpx(uint64_t& a, uint64_t& b) {
  if (b < a) { uint64_t t = a; a = b; b = t; }
}

// full == 1: sort each tile completely; full == 0: only the disperse steps hh = TILE/2 .. 1
__global__ void __launch_bounds__(SORT_THREADS) sort_tile_kernel(uint64_t* d, int64_t n, int full) {
  __shared__ uint64_t s[SORT_TILE];
  const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
  for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) s[i] = (base + i < n) ? d[base + i] : EMPTY;
  __syncthreads();
  if (full) {
    for (int h = 1; h < SORT_TILE; h <<= 1) {
      for (int i = threadIdx.x; i < SORT_TILE / 2; i += SORT_THREADS) {
        int blk = i / h, off = i % h;
        int lo = blk * 2 * h + off, hi = blk * 2 * h + 2 * h - 1 - off;
        cmpx(s[lo], s[hi]);
      }
      __syncthreads();
      for (int hh = h >> 1; hh >= 1; hh >>= 1) {
        for (int i = threadIdx.x; i < SORT_TILE / 2; i += SORT_THREADS) {
          int blk = i / hh, off = i % hh;
          int lo = blk * 2 * hh + off;
          cmpx(s[lo], s[lo + hh]);
        }
        __syncthreads();
      }
    }
  } else {
    for (int hh = SORT_TILE / 2; hh >= 1; hh >>= 1) {
      for (int i = threadIdx.x; i < SORT_TILE / 2; i += SORT_THREADS) {
        int blk = i / hh, off = i % hh;
        int lo = blk * 2 * hh + off;
        cmpx(s[lo], s[lo + hh]);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS)
    if (base + i < n) d[base + i] = s[i];
}

__global__ void sort_flip_kernel(uint64_t* d, int64_t n, int64_t h, int64_t n_pairs) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  int64_t blk = i / h, off = i % h;
  int64_t lo = blk * 2 * h + off, hi = blk * 2 * h + 2 * h - 1 - off;
  if (hi < n) {
    uint64_t a = d[lo], b = d[hi];
    if (b < a) { d[lo] = b; d[hi] = a; }
  }
}

__global__ void sort_disperse_kernel(uint64_t* d, int64_t n, int64_t hh, int64_t n_pairs) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  int64_t blk = i / hh, off = i % hh;
  int64_t lo = blk * 2 * hh + off, hi = lo + hh;
  if (hi < n) {
    uint64_t a = d[lo], b = d[hi];
    if (b < a) { d[lo] = b; d[hi] = a; }
  }
}

// Several sorted runs of DISTINCT keys, back to back (run j = keys[run_ptr[j] .. run_ptr[j + 1])), merged by ranking:
// the place of a key in the merged order = its index in its own run + the number of smaller keys in every other run
// (one binary search each).  Every rank of the multi-GPU path sorts its own rare keys; the all-gathered runs are
// merged with this instead of one bitonic sort of the whole set (whose n log^2 n grows with the number of GPUs).
__global__ void merge_runs_kernel(const uint64_t* __restrict__ keys, const int64_t* __restrict__ run_ptr, int n_runs,
                                  uint64_t* __restrict__ out) {
  const int64_t n = run_ptr[n_runs];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t x = keys[i];
  int64_t rank = 0;
  for (int j = 0; j < n_runs; ++j) {
    int64_t lo = run_ptr[j], hi = run_ptr[j + 1];
    if (i >= lo && i < hi) {
      rank += i - lo;
      continue;
    }
    const int64_t base = lo;
    while (lo < hi) {  // first element of run j that is not smaller than x
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(keys + mid) < x) lo = mid + 1; else hi = mid;
    }
    rank += lo - base;
  }
  out[rank] = x;
}

__global__ void index_filter_kernel(const uint64_t* __restrict__ sorted_keys, int64_t n, int filter_bits, uint32_t* filter) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t b = mix64(sorted_keys[i]) >> (64 - filter_bits);
  atomicOr(filter + (b >> 5), 1u << (b & 31u));
}

__global__ void index_build_kernel(const uint64_t* __restrict__ sorted_keys, int64_t n, uint64_t* idx_keys,
                                   uint32_t* idx_vals, int64_t cap, int64_t* counters) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t slot = table_upsert(idx_keys, cap, sorted_keys[i]);
  if (slot < 0) { counters[0] = 1; return; }
  idx_vals[slot] = (uint32_t)i;
}

// ============================================================================================
// Block-level helpers
// ============================================================================================
// ascending bitonic sort of data[0..n) (shared or global memory), all threads of the block
__device__ void block_sort_u32(uint32_t* data, int n) {
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int h = 1; h < np2; h <<= 1) {
    for (int i = threadIdx.x; i < np2 / 2; i += blockDim.x) {
      int blk = i / h, off = i % h;
      int lo = blk * 2 * h + off, hi = blk * 2 * h + 2 * h - 1 - off;
      if (hi < n) {
        uint32_t a = data[lo], b = data[hi];
        if (b < a) { data[lo] = b; data[hi] = a; }
      }
    }
    __syncthreads();
    for (int hh = h >> 1; hh >= 1; hh >>= 1) {
      for (int i = threadIdx.x; i < np2 / 2; i += blockDim.x) {
        int blk = i / hh, off = i % hh;
        int lo = blk * 2 * hh + off, hi = lo + hh;
        if (hi < n) {
          uint32_t a = data[lo], b = data[hi];
          if (b < a) { data[lo] = b; data[hi] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// exclusive prefix sum of one int per thread across the block (blockDim.x <= 1024); also returns the total
__device__ int block_exclusive_scan(int v, int* total, int* s_warp /* [32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  int incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarps ? s_warp[lane] : 0;
    int wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(FULL, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;  // exclusive warp offsets; lane 31 keeps grand total below
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  int res = s_warp[warp] + incl - v;
  __syncthreads();
  return res;
}

// ============================================================================================
// Stage B: one block per unit.  Probe every k-mer of the unit in the rare-set index, collect
// hit ids, sort + unique them, write them to the unit's scratch run.
// ============================================================================================
constexpr int CL_THREADS = 256;
constexpr int CL_RUN = 8;        // consecutive k-mer starts rolled by one thread
constexpr int CL_SMEM_IDS = 8192;  // units with more k-mer starts than this sort in global scratch

__global__ void __launch_bounds__(CL_THREADS)
cloud_build_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ unit_off,
                   const int32_t* __restrict__ unit_len, const int64_t* __restrict__ unit_kbase, int k,
                   const uint64_t* __restrict__ idx_keys, const uint32_t* __restrict__ idx_vals, int64_t cap,
                   uint32_t* tmp_ids, int32_t* unit_cnt) {
  __shared__ uint32_t s_ids[CL_SMEM_IDS];
  __shared__ int s_n;
  __shared__ int s_total;
  __shared__ int s_warp[32];
  const int64_t u = blockIdx.x;
  const int nk = unit_len[u] - k + 1;
  if (nk <= 0) {
    if (threadIdx.x == 0) unit_cnt[u] = 0;
    return;
  }
  uint32_t* out = tmp_ids + unit_kbase[u];
  uint32_t* list = (nk <= CL_SMEM_IDS) ? s_ids : out;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const int64_t off = unit_off[u];
  const uint64_t mask = (1ull << (2 * k)) - 1;
  for (int p0 = threadIdx.x * CL_RUN; p0 < nk; p0 += CL_THREADS * CL_RUN) {
    uint64_t kmer = 0;
    for (int i = 0; i < k - 1; ++i) kmer = (kmer << 2) | base_at(packed, off + p0 + i);
    const int pend = min(p0 + CL_RUN, nk);
    for (int p = p0; p < pend; ++p) {
      kmer = ((kmer << 2) | base_at(packed, off + p + k - 1)) & mask;
      int64_t slot = table_find(idx_keys, cap, kmer);
      if (slot >= 0) list[atomicAdd(&s_n, 1)] = __ldg(idx_vals + slot);
    }
  }
  __syncthreads();
  const int n = s_n;
  if (n == 0) {
    if (threadIdx.x == 0) unit_cnt[u] = 0;
    return;
  }
  block_sort_u32(list, n);
  // unique: keep the first of every run; tiles of blockDim elements, read phase / write phase
  int written = 0;
  for (int t0 = 0; t0 < n; t0 += CL_THREADS) {
    int i = t0 + threadIdx.x;
    uint32_t v = 0;
    int keep = 0;
    if (i < n) {
      v = list[i];
      keep = (i == 0) || (list[i - 1] != v);
    }
    __syncthreads();
    int tile_total;
    int pos = block_exclusive_scan(keep, &s_total, s_warp);
    tile_total = s_total;
    if (keep) out[written + pos] = v;
    written += tile_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) unit_cnt[u] = written;
}

// --------------------------------------------------------------------------------------------
// Stage B, warp form (the default): one WARP per unit, no block barriers.
//   * lane l rolls the contiguous run [l * S, (l + 1) * S) of the unit's k-mer starts (S = ceil(nk / 32)): one packed
//     word per 16 bases instead of one load per base, and the first probes of CLW_BATCH consecutive k-mers are in
//     flight together (every probe is an L2 round trip into the rare index);
//   * hits are appended to a warp-private shared-memory list with ballot / popc (no atomics); a unit with more hits
//     than the list holds is collected again straight into its global scratch run;
//   * sort + unique are warp-synchronous (bitonic network over the list, __syncwarp between stages).
// --------------------------------------------------------------------------------------------
constexpr int CLW_WARPS = 8;      // warps (= units in flight) per block
constexpr int CLW_LIST = 512;     // ids per warp-private list
constexpr int CLW_BATCH = 4;      // k-mers whose first probe is issued together

// ascending bitonic sort of data[0..n) by one warp (shared or global memory)
__device__ void warp_sort_u32(uint32_t* data, int n) {
  const int lane = threadIdx.x & 31;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int h = 1; h < np2; h <<= 1) {
    for (int i = lane; i < np2 / 2; i += 32) {
      const int blk = i / h, off = i % h;
      const int lo = blk * 2 * h + off, hi = blk * 2 * h + 2 * h - 1 - off;
      if (hi < n) {
        const uint32_t a = data[lo], b = data[hi];
        if (b < a) { data[lo] = b; data[hi] = a; }
      }
    }
    __syncwarp();
    for (int hh = h >> 1; hh >= 1; hh >>= 1) {
      for (int i = lane; i < np2 / 2; i += 32) {
        const int blk = i / hh, off = i % hh;
        const int lo = blk * 2 * hh + off, hi = lo + hh;
        if (hi < n) {
          const uint32_t a = data[lo], b = data[hi];
          if (b < a) { data[lo] = b; data[hi] = a; }
        }
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(CLW_WARPS * 32)
cloud_build_warp_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ unit_off,
                        const int32_t* __restrict__ unit_len, const int64_t* __restrict__ unit_kbase, int64_t n_units, int k,
                        const uint64_t* __restrict__ idx_keys, const uint32_t* __restrict__ idx_vals, int64_t cap,
                        const uint32_t* __restrict__ filter, int filter_bits, uint32_t* tmp_ids, int32_t* unit_cnt) {
  __shared__ uint32_t s_list[CLW_WARPS][CLW_LIST];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const int64_t u = (int64_t)blockIdx.x * CLW_WARPS + warp;
  if (u >= n_units) return;
  const int nk = unit_len[u] - k + 1;
  if (nk <= 0) {
    if (lane == 0) unit_cnt[u] = 0;
    return;
  }
  uint32_t* out = tmp_ids + unit_kbase[u];  // nk slots
  const int64_t off = unit_off[u];
  const uint64_t mask = (1ull << (2 * k)) - 1;
  const int S = (nk + 31) >> 5;
  const int p_begin = min(lane * S, nk), p_end = min(p_begin + S, nk);
  uint32_t* list = s_list[warp];
  int list_cap = CLW_LIST;
  int n = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    n = 0;
    // the run's first k - 1 bases
    int64_t pos = off + p_begin;  // absolute base index of the next base to shift in
    uint32_t word = 0;
    uint64_t kmer = 0;
    if (p_begin < p_end) {
      word = __ldg(packed + (pos >> 4));
      for (int i = 0; i < k - 1; ++i, ++pos) {
        if ((pos & 15) == 0) word = __ldg(packed + (pos >> 4));
        kmer = (kmer << 2) | ((word >> ((pos & 15) << 1)) & 3u);
      }
    }
    for (int p0 = p_begin; __any_sync(FULL, p0 < p_end); p0 += CLW_BATCH) {
      uint64_t km[CLW_BATCH], key[CLW_BATCH];
      int64_t slot[CLW_BATCH];
#pragma unroll
      for (int j = 0; j < CLW_BATCH; ++j) {
        km[j] = EMPTY;  // never a k-mer (k <= 31)
        key[j] = EMPTY;
        slot[j] = 0;
        if (p0 + j < p_end) {
          if ((pos & 15) == 0) word = __ldg(packed + (pos >> 4));
          kmer = ((kmer << 2) | ((word >> ((pos & 15) << 1)) & 3u)) & mask;
          ++pos;
          const uint64_t h = mix64(kmer);
          // index larger than L2 (the rare set of several GPUs): one bit per hash prefix, L2 resident, answers "not
          // rare" for the ~97 % of k-mers that are not, before the probe that would go to DRAM
          bool maybe = true;
          if (filter != nullptr) {
            const uint64_t b = h >> (64 - filter_bits);
            maybe = (__ldg(filter + (b >> 5)) >> (b & 31u)) & 1u;
          }
          if (maybe) {
            km[j] = kmer;
            slot[j] = home_slot(h, cap);
            key[j] = __ldg(idx_keys + slot[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < CLW_BATCH; ++j) {
        uint32_t id = 0;
        bool hit = false;
        if (km[j] != EMPTY) {
          uint64_t cur = key[j];
          int64_t sl = slot[j];
          for (int64_t probes = 0; probes < cap; ++probes) {
            if (cur == km[j]) { hit = true; break; }
            if (cur == EMPTY) break;
            if (++sl == cap) sl = 0;
            cur = __ldg(idx_keys + sl);
          }
          if (hit) id = __ldg(idx_vals + sl);
        }
        const unsigned bal = __ballot_sync(FULL, hit);
        if (hit) {
          const int at = n + __popc(bal & lt);
          if (at < list_cap) list[at] = id;
        }
        n += __popc(bal);
      }
    }
    __syncwarp();
    if (n <= list_cap) break;
    list = out;  // more hits than the shared list holds: collect again into the unit's global run (nk slots)
    list_cap = nk;
  }
  if (n == 0) {
    if (lane == 0) unit_cnt[u] = 0;
    return;
  }
  warp_sort_u32(list, n);
  // unique: keep the first of every run; tiles of 32 (in place when list == out: a tile is read before it is written,
  // and writes never pass the read position)
  int written = 0;
  uint32_t prev = 0;
  for (int t0 = 0; t0 < n; t0 += 32) {
    const int i = t0 + lane;
    uint32_t v = 0;
    if (i < n) v = list[i];
    uint32_t before = __shfl_up_sync(FULL, v, 1);
    if (lane == 0) before = prev;
    const bool keep = i < n && (i == 0 || before != v);
    const unsigned bal = __ballot_sync(FULL, keep);
    prev = __shfl_sync(FULL, v, 31);
    __syncwarp();
    if (keep) out[written + __popc(bal & lt)] = v;
    written += __popc(bal);
    __syncwarp();
  }
  if (lane == 0) unit_cnt[u] = written;
}

// ---- exclusive scan int32 -> int64 (3 phases) ------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* partial) {
  __shared__ int64_t s[SCAN_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int64_t sum = 0;
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    int64_t i = base + j * SCAN_THREADS + threadIdx.x;
    if (i < n) sum += in[i];
  }
  for (int o = 16; o >= 1; o >>= 1) sum += __shfl_down_sync(FULL, sum, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += s[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void scan_partials_kernel(int64_t* partial, int64_t nb) {
  // single thread block, sequential over tiles of blockDim partials
  __shared__ int64_t s[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t t0 = 0; t0 < nb; t0 += blockDim.x) {
    int64_t i = t0 + threadIdx.x;
    int64_t v = i < nb ? partial[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {
      int64_t t = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) partial[i] = carry + s[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += s[threadIdx.x];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                  const int64_t* __restrict__ partial, int64_t* out) {
  // thread t owns SCAN_ITEMS consecutive inputs
  __shared__ int64_t s_warp[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int64_t sum = 0;
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    v[j] = (base + j < n) ? in[base + j] : 0;
    sum += v[j];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int64_t woff = 0;
  for (int w = 0; w < warp; ++w) woff += s_warp[w];
  int64_t run = partial[blockIdx.x] + woff + incl - sum;
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (base + j < n) out[base + j + 1] = run + v[j];
    run += v[j];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}

__global__ void cloud_compact_kernel(const uint32_t* __restrict__ tmp_ids, const int64_t* __restrict__ unit_kbase,
                                     const int64_t* __restrict__ unit_ptr, int64_t n_units, uint32_t* ids) {
  const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= n_units) return;
  const int lane = threadIdx.x & 31;
  const int64_t dst = unit_ptr[u], cnt = unit_ptr[u + 1] - dst, src = unit_kbase[u];
  for (int64_t i = lane; i < cnt; i += 32) ids[dst + i] = tmp_ids[src + i];
}

// ---- multiplicity histogram / filter / occurrence lists (warp per unit) ----------------------
__device__ __forceinline__ int64_t lower_bound_u32(const uint32_t* __restrict__ v, int64_t lo, int64_t hi, int64_t x) {
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)__ldg(v + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void id_histogram_kernel(const int64_t* __restrict__ unit_ptr, const uint32_t* __restrict__ ids,
                                    int64_t unit_lo, int64_t unit_hi, int32_t* mult) {
  const int64_t u = unit_lo + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (u >= unit_hi) return;
  const int lane = threadIdx.x & 31;
  const int64_t e1 = unit_ptr[u + 1];
  for (int64_t e = unit_ptr[u] + lane; e < e1; e += 32) atomicAdd(mult + ids[e], 1);
}

__global__ void cloud_filter_count_kernel(const int64_t* __restrict__ unit_ptr, const uint32_t* __restrict__ ids,
                                          int64_t n_units, const int32_t* __restrict__ mult, int64_t min_mult,
                                          int64_t max_mult, int32_t* new_cnt) {
  const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= n_units) return;
  const int lane = threadIdx.x & 31;
  const int64_t e0 = unit_ptr[u], e1 = unit_ptr[u + 1];
  int c = 0;
  for (int64_t e = e0 + lane; e < e1; e += 32) {
    int64_t m = mult[ids[e]];
    c += (m >= min_mult && m <= max_mult);
  }
  for (int o = 16; o >= 1; o >>= 1) c += __shfl_down_sync(FULL, c, o);
  if (lane == 0) new_cnt[u] = c;
}

__global__ void cloud_filter_write_kernel(const int64_t* __restrict__ unit_ptr, const uint32_t* __restrict__ ids,
                                          int64_t n_units, const int32_t* __restrict__ mult, int64_t min_mult,
                                          int64_t max_mult, const int64_t* __restrict__ new_ptr, uint32_t* new_ids) {
  const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= n_units) return;
  const int lane = threadIdx.x & 31;
  const int64_t e0 = unit_ptr[u], e1 = unit_ptr[u + 1];
  int64_t dst = new_ptr[u];
  for (int64_t eb = e0; eb < e1; eb += 32) {  // order-preserving (ids stay sorted inside the unit)
    int64_t e = eb + lane;
    uint32_t id = 0;
    bool keep = false;
    if (e < e1) {
      id = ids[e];
      int64_t m = mult[id];
      keep = m >= min_mult && m <= max_mult;
    }
    unsigned bal = __ballot_sync(FULL, keep);
    if (keep) new_ids[dst + __popc(bal & ((1u << lane) - 1))] = id;
    dst += __popc(bal);
  }
}

// cursor[a] starts at occ_ptr[a] (the list lengths sum to the number of cloud entries, < 2^32), so that filling an
// occurrence is one atomic and one store -- no second random load of occ_ptr[id]
__global__ void occ_cursor_kernel(const int64_t* __restrict__ occ_ptr, int64_t n_kmers, uint32_t* __restrict__ cursor) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n_kmers) cursor[a] = (uint32_t)occ_ptr[a];
}

__global__ void occ_fill_kernel(const int64_t* __restrict__ unit_ptr, const uint32_t* __restrict__ ids,
                                int64_t unit_lo, int64_t unit_hi, uint32_t* cursor, uint32_t* occ) {
  const int64_t u = unit_lo + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (u >= unit_hi) return;
  const int lane = threadIdx.x & 31;
  const int64_t e1 = unit_ptr[u + 1];
  for (int64_t e = unit_ptr[u] + lane; e < e1; e += 32) occ[atomicAdd(cursor + __ldg(ids + e), 1u)] = (uint32_t)u;
}

// The same two passes for the ids of [id_lo, id_hi) only (the lists inside a unit are sorted: two binary searches bound
// the slice): with G GPUs every rank inverts 1/G of the id space of the all-gathered clouds and the lists are
// all-gathered, instead of every rank inverting everything.  mult / cursor / occ_ptr are indexed by id - id_lo.
template <bool FILL>
__global__ void occ_slice_kernel(const int64_t* __restrict__ unit_ptr, const uint32_t* __restrict__ ids, int64_t n_units,
                                 int64_t id_lo, int64_t id_hi, int32_t* mult, uint32_t* cursor, uint32_t* occ) {
  const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;  // half a warp per unit
  if (u >= n_units) return;
  const int lane = threadIdx.x & 15;
  const int64_t b = __ldg(unit_ptr + u), e = __ldg(unit_ptr + u + 1);
  const int64_t e0 = lower_bound_u32(ids, b, e, id_lo), e1 = lower_bound_u32(ids, e0, e, id_hi);
  for (int64_t i = e0 + lane; i < e1; i += 16) {
    const int64_t a = (int64_t)__ldg(ids + i) - id_lo;
    if (FILL) occ[atomicAdd(cursor + a, 1u)] = (uint32_t)u;
    else atomicAdd(mult + a, 1);
  }
}

__global__ void occ_sort_kernel(const int64_t* __restrict__ occ_ptr, uint32_t* occ, int64_t n_kmers) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_kmers) return;
  const int64_t b = occ_ptr[a], m = occ_ptr[a + 1] - b;
  for (int64_t i = 1; i < m; ++i) {  // insertion sort; lists are a few dozen entries
    uint32_t v = occ[b + i];
    int64_t j = i - 1;
    while (j >= 0 && occ[b + j] > v) { occ[b + j + 1] = occ[b + j]; --j; }
    occ[b + j + 1] = v;
  }
}

// usplit[7 u + j - 1] = position of the first id >= (n_kmers * j) >> 3 in unit u's sorted list, j = 1..7:
// lets stage C cut any unit list at octant boundaries of the id space without searching.
__global__ void occ_last_kernel(const uint32_t* __restrict__ occ, int64_t n, const uint32_t* __restrict__ unit_last,
                                uint32_t* __restrict__ occ_last) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) occ_last[i] = __ldg(unit_last + occ[i]);
}

__global__ void unit_split_kernel(const int64_t* __restrict__ unit_ptr, const uint32_t* __restrict__ ids, int64_t n_units,
                                  int64_t n_kmers, uint32_t* usplit) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_units * 7) return;
  const int64_t u = i / 7;
  const int j = (int)(i % 7) + 1;
  usplit[i] = (uint32_t)lower_bound_u32(ids, unit_ptr[u], unit_ptr[u + 1], (n_kmers * j) >> 3);
}

// ============================================================================================
// Stage C: pair candidates.  One warp per source id a.  The distances [dmin, dlim] are cut into
// chunks [d0, d1]; for one chunk the warp streams, per occurrence g of a, the CONTIGUOUS id run
// ids[unit_ptr[g+d0] .. unit_ptr[min(g+d1,last)+1]) into a warp-private shared-memory table
//     slot = (b + 1) << cb | count            count = sum over d in the chunk of cnt[d][a][b]
// and emits (a, b, d0, d1) for every b whose chunk total reaches min_cov -- a necessary
// condition for any single cnt[d][a][b] >= min_cov.  Stage D resolves the exact per-distance
// counts of those few pairs by joining the two occurrence lists.
//
// Insert protocol (no atomics, no per-probe synchronisation): the lanes of one step hold
// DISTINCT keys (one sorted-unique unit list, or a 32-run spanning unit boundaries de-duplicated
// with match.any), so an occupied slot is only ever updated by the single lane that holds its
// key.  A lane that finds an empty slot CLAIMS it speculatively, writing key and increment in
// one store, and moves on; at the start of the next step (after one __syncwarp) it re-reads the
// slot: if another lane's claim landed there instead, the loser re-inserts its key in a rare
// replay round.  The table is kept at most half full so that probe chains stay short.
// Pruning (exact): cnt[d][a][b] >= min_cov needs >= min_cov occurrences g with g + d inside the
// read, so distances beyond the min_cov-th largest remainder are never streamed.
// ============================================================================================
constexpr int PC_WARPS = CFK_PAIR_WARPS;
constexpr int PC_TBL_BYTES = CFK_PAIR_TABLE_BYTES;
constexpr uint32_t PC_NONE = 0xFFFFFFFFu;

constexpr bool pc_is_prime(uint32_t n) {
  if (n < 2) return false;
  for (uint32_t d = 2; d * d <= n; ++d)
    if (n % d == 0) return false;
  return true;
}
constexpr uint32_t pc_prime_le(uint32_t n) { return pc_is_prime(n) ? n : pc_prime_le(n - 1); }
// Slots used of a warp's table: the largest prime that fits, so that double hashing
// (slot += step, any step in [1, NS - 1]) visits every slot.  Linear probing's long clusters
// are poison here: a step of the warp costs the LONGEST probe chain among its 64 keys.
template <typename S>
struct PairTable {
  static constexpr uint32_t NS = pc_prime_le(PC_TBL_BYTES / (uint32_t)sizeof(S));
};
__device__ __forceinline__ uint32_t pc_slot(uint32_t b, uint32_t ns) { return __umulhi(b * 2654435761u, ns); }
__device__ __forceinline__ uint32_t pc_step(uint32_t b, uint32_t ns) { return 1u + __umulhi(b * 0x85EBCA77u + 0x9E3779B9u, ns - 1u); }

struct PairArgs {
  const int64_t* __restrict__ unit_ptr;
  const uint32_t* __restrict__ ids;
  const uint32_t* __restrict__ unit_last;
  const uint32_t* __restrict__ occ_a;
  const uint32_t* __restrict__ usplit;  // [7 per unit] octant split positions, may be null
  int64_t m;
  int64_t n_kmers;
  uint32_t a;
  uint32_t min_cov;  // >= 1
  uint4* cand;
  int64_t max_cand;
  int64_t* counters;
};

// shared-memory accesses by 32-bit shared-space address (cheaper addressing than generic pointers,
// and the "memory" clobber keeps the compiler from caching table words across the warp barrier)
template <typename S> __device__ __forceinline__ S lds(uint32_t a);
template <> __device__ __forceinline__ uint32_t lds<uint32_t>(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
template <> __device__ __forceinline__ uint64_t lds<uint64_t>(uint32_t a) {
  uint64_t v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts(uint32_t a, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename S>
struct PairLane {      // per-lane insert state carried from one step to the next
  uint32_t vaddr[2];   // shared addresses of the not yet verified claims, PC_NONE = none
  S vword[2];          // what the claims wrote
  int claims;          // slots this lane has claimed in the current pass
#ifdef CFK_PAIR_STATS
  long long st_iters = 0, st_keys = 0, st_windows = 0;
#endif
};

// One insert step: every lane counts up to two ids (all ids of a step are distinct).  Probes until
// each key is counted or an empty slot is claimed; per-lane loop, no votes inside, the two
// probe chains of a lane are independent (ILP).
template <typename S>
__device__ __forceinline__ void pair_probe2(uint32_t tbase, PairLane<S>& L, S key0, uint32_t h0, uint32_t s0, bool v0,
                                            S key1, uint32_t h1, uint32_t s1, bool v1, int cb) {
  constexpr uint32_t NS = PairTable<S>::NS;
  while (v0 || v1) {
#ifdef CFK_PAIR_STATS
    ++L.st_iters;
#endif
    const uint32_t a0 = tbase + h0 * (uint32_t)sizeof(S), a1 = tbase + h1 * (uint32_t)sizeof(S);
    S w0 = 0, w1 = 0;
    if (v0) w0 = lds<S>(a0);
    if (v1) w1 = lds<S>(a1);
    if (v0) {
      if ((w0 >> cb) == key0) {
        sts(a0, (S)(w0 + 1));
        v0 = false;
      } else if (w0 == 0) {
        const S nw = (key0 << cb) | (S)1;
        sts(a0, nw);  // speculative claim, verified at the start of the next step
        L.vaddr[0] = a0;
        L.vword[0] = nw;
        ++L.claims;
        v0 = false;
      } else {
        h0 += s0;
        if (h0 >= NS) h0 -= NS;
      }
    }
    if (v1) {
      if ((w1 >> cb) == key1) {
        sts(a1, (S)(w1 + 1));
        v1 = false;
      } else if (w1 == 0) {
        const S nw = (key1 << cb) | (S)1;
        sts(a1, nw);
        L.vaddr[1] = a1;
        L.vword[1] = nw;
        ++L.claims;
        v1 = false;
      } else {
        h1 += s1;
        if (h1 >= NS) h1 -= NS;
      }
    }
  }
}

// Verify the claims of the previous step; a lane whose claim was overwritten by another claim of
// the same slot re-inserts its key (rare).
template <typename S>
__device__ __forceinline__ void pair_verify(uint32_t tbase, PairLane<S>& L, int cb) {
  constexpr uint32_t NS = PairTable<S>::NS;
  __syncwarp();
  bool lost0 = false, lost1 = false;
  if (L.vaddr[0] != PC_NONE) lost0 = lds<S>(L.vaddr[0]) != L.vword[0];
  if (L.vaddr[1] != PC_NONE) lost1 = lds<S>(L.vaddr[1]) != L.vword[1];
  while (__any_sync(FULL, lost0 || lost1)) {  // replay round: the losers hold distinct keys, nobody else inserts
    const S k0 = L.vword[0] >> cb, k1 = L.vword[1] >> cb;
    const uint32_t s0 = pc_step((uint32_t)k0 - 1u, NS), s1 = pc_step((uint32_t)k1 - 1u, NS);
    uint32_t h0 = (L.vaddr[0] - tbase) / (uint32_t)sizeof(S) + s0, h1 = (L.vaddr[1] - tbase) / (uint32_t)sizeof(S) + s1;
    if (h0 >= NS) h0 -= NS;
    if (h1 >= NS) h1 -= NS;
    if (!lost0) h0 = 0;
    if (!lost1) h1 = 0;
    L.vaddr[0] = PC_NONE;
    L.vaddr[1] = PC_NONE;
    L.claims -= (int)lost0 + (int)lost1;
    pair_probe2<S>(tbase, L, k0, h0, s0, lost0, k1, h1, s1, lost1, cb);
    __syncwarp();
    lost0 = lost0 && L.vaddr[0] != PC_NONE && lds<S>(L.vaddr[0]) != L.vword[0];
    lost1 = lost1 && L.vaddr[1] != PC_NONE && lds<S>(L.vaddr[1]) != L.vword[1];
  }
  L.vaddr[0] = PC_NONE;
  L.vaddr[1] = PC_NONE;
}

// position of the first id >= octant boundary j (0..8) in unit u's sorted list
__device__ __forceinline__ uint32_t split_pos(const PairArgs& A, int64_t u, int j) {
  if (j <= 0) return (uint32_t)__ldg(A.unit_ptr + u);
  if (j >= 8) return (uint32_t)__ldg(A.unit_ptr + u + 1);
  return __ldg(A.usplit + u * 7 + (j - 1));
}

// One table pass over distances [d0, d1] restricted to ids [lo_id, hi_id) (j_lo / j_hi: the same
// range as octant indices of the precomputed per-unit split table, or -1).  Returns the number of
// distinct keys, or -1 if the table passed its maximum load (nothing emitted).
template <typename S>
__device__ int pair_chunk_pass(uint32_t tbase, const PairArgs& A, int d0, int d1, int64_t lo_id, int64_t hi_id, int j_lo,
                               int j_hi, int cb) {
  constexpr uint32_t NS = PairTable<S>::NS;
  constexpr int MAXLOAD = (int)(NS / 2);
  const int lane = threadIdx.x & 31;
  const bool whole = (lo_id == 0 && hi_id >= A.n_kmers);
#pragma unroll 4
  for (int i = lane; i < PC_TBL_BYTES / 16; i += 32) sts128(tbase + (uint32_t)i * 16u, make_uint4(0, 0, 0, 0));
  PairLane<S> L;
  L.vaddr[0] = L.vaddr[1] = PC_NONE;
  L.vword[0] = L.vword[1] = 0;
  L.claims = 0;
  for (int64_t t0 = 0; t0 < A.m; t0 += 32) {
    const int64_t t = t0 + lane;
    int64_t ua = 1, ub = 0;  // unit run [ua, ub] of this lane's occurrence
    uint32_t cut_b = 0, cut_e = 0;
    if (t < A.m) {
      const int64_t g = (int64_t)__ldg(A.occ_a + t);
      const int64_t last = (int64_t)__ldg(A.unit_last + g);
      if (g + d0 <= last) {
        ua = g + d0;
        ub = min(g + (int64_t)d1, last);
        if (!whole) {  // single unit (d0 == d1): cut its sorted list to the id range
          if (j_lo >= 0 && A.usplit != nullptr) {
            cut_b = split_pos(A, ua, j_lo);
            cut_e = split_pos(A, ua, j_hi);
          } else {
            const int64_t b64 = __ldg(A.unit_ptr + ua), e64 = __ldg(A.unit_ptr + ua + 1);
            const int64_t nb = lower_bound_u32(A.ids, b64, e64, lo_id);
            cut_b = (uint32_t)nb;
            cut_e = (uint32_t)((hi_id >= A.n_kmers) ? e64 : lower_bound_u32(A.ids, nb, e64, hi_id));
          }
          if (cut_e <= cut_b) { ua = 1; ub = 0; }
        }
      }
    }
    unsigned runs = __ballot_sync(FULL, ua <= ub);
    while (runs) {
      const int src = __ffs(runs) - 1;
      runs &= runs - 1;
      const int64_t us = __shfl_sync(FULL, ua, src);
      const int nu = (int)(__shfl_sync(FULL, ub, src) - us) + 1;  // 1..31 units
      // unit boundaries of the run, one per lane: unit j spans [bnd_j, bnd_{j+1})
      uint32_t bnd = 0;
      if (whole) {
        if (lane <= nu) bnd = (uint32_t)__ldg(A.unit_ptr + us + lane);  // < 2^32 cloud entries (checked by the host)
      } else {
        const uint32_t cb0 = __shfl_sync(FULL, cut_b, src), ce0 = __shfl_sync(FULL, cut_e, src);
        bnd = lane == 0 ? cb0 : ce0;
      }
      int j = 0;
      uint32_t p = __shfl_sync(FULL, bnd, 0), e = __shfl_sync(FULL, bnd, 1);
      while (p >= e && j + 1 < nu) { ++j; p = e; e = __shfl_sync(FULL, bnd, j + 1); }
      if (p >= e) continue;
      uint32_t b0 = (p + lane < e) ? __ldg(A.ids + p + lane) : A.a;
      uint32_t b1 = (p + 32u + lane < e) ? __ldg(A.ids + p + 32u + lane) : A.a;
      for (;;) {
        // next window: the rest of this unit, else the next non-empty unit of the run
        uint32_t np = p + 64u, ne = e;
        int nj = j;
        while (np >= ne && nj + 1 < nu) { ++nj; np = ne; ne = __shfl_sync(FULL, bnd, nj + 1); }
        const bool more = np < ne;
        uint32_t n0 = A.a, n1 = A.a;
        if (more) {  // its ids fly during the insert of the current window
          if (np + lane < ne) n0 = __ldg(A.ids + np + lane);
          if (np + 32u + lane < ne) n1 = __ldg(A.ids + np + 32u + lane);
        }
        pair_verify<S>(tbase, L, cb);
#ifdef CFK_PAIR_STATS
        L.st_keys += (b0 != A.a) + (b1 != A.a);
        ++L.st_windows;
#endif
        if (__reduce_add_sync(FULL, L.claims) > MAXLOAD) return -1;
        // ids of one sorted-unique unit list are distinct; id a itself (also the filler of idle lanes) is skipped
        pair_probe2<S>(tbase, L, (S)b0 + 1, pc_slot(b0, NS), pc_step(b0, NS), b0 != A.a, (S)b1 + 1, pc_slot(b1, NS),
                       pc_step(b1, NS), b1 != A.a, cb);
        if (!more) break;
        p = np; e = ne; j = nj; b0 = n0; b1 = n1;
      }
    }
  }
  pair_verify<S>(tbase, L, cb);
  const int distinct = __reduce_add_sync(FULL, L.claims);
#ifdef CFK_PAIR_STATS
  {  // [4] passes, [5] windows, [6] lane-iterations of the probe loop, [7] keys inserted
    atomicAdd((unsigned long long*)A.counters + 6, (unsigned long long)L.st_iters);
    atomicAdd((unsigned long long*)A.counters + 7, (unsigned long long)L.st_keys);
    if (lane == 0) {
      atomicAdd((unsigned long long*)A.counters + 4, 1ull);
      atomicAdd((unsigned long long*)A.counters + 5, (unsigned long long)L.st_windows);
    }
  }
#endif
  // emit (a, b, d0, d1) for every key whose chunk total reached min_cov
  const S cmask = ((S)1 << cb) - 1;
  constexpr int PER16 = 16 / (int)sizeof(S);
  for (int i = lane; i < PC_TBL_BYTES / 16; i += 32) {
    const uint4 q = lds128(tbase + (uint32_t)i * 16u);
    S w[PER16];
    if constexpr (sizeof(S) == 4) { w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; }
    else { w[0] = ((uint64_t)q.y << 32) | q.x; w[1] = ((uint64_t)q.w << 32) | q.z; }
    uint32_t best = 0;
#pragma unroll
    for (int jj = 0; jj < PER16; ++jj) best = max(best, (uint32_t)(w[jj] & cmask));
    if (__any_sync(FULL, best >= A.min_cov)) {  // rare
#pragma unroll
      for (int jj = 0; jj < PER16; ++jj) {
        const bool take = (uint32_t)(w[jj] & cmask) >= A.min_cov;  // min_cov >= 1, so empty slots never pass
        const int64_t pos = warp_append(take, A.counters);
        if (take && pos < A.max_cand)
          A.cand[pos] = make_uint4(A.a, (uint32_t)(w[jj] >> cb) - 1u, (uint32_t)d0, (uint32_t)d1);
      }
    }
  }
  __syncwarp();
  return distinct;
}

__device__ __forceinline__ int64_t warp_sum_i64(int64_t v) {
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

template <typename S>
__device__ void pair_source(uint32_t tbase, const PairArgs& A, int dmin, int dlim, int cb, float& ratio, int64_t& splits) {
  constexpr int NS = (int)PairTable<S>::NS;
  constexpr int TARGET = NS * 3 / 10;  // planned number of distinct keys per pass (hard limit NS / 2)
  const int lane = threadIdx.x & 31;
  const int64_t cnt_limit = (cb >= 32) ? (int64_t)0x7FFFFFFF : (((int64_t)1 << cb) - 1);
  int d0 = dmin;
  int nd_force = 31;
  while (d0 <= dlim) {
    // distinct/entries of the previous pass predicts this one; farther distances repeat less, hence the margin
    const int64_t cap = max((int64_t)64, (int64_t)((float)TARGET / fminf(1.0f, 1.25f * ratio)));
    // plan: extend the chunk one distance at a time (4 looked up per round) while its id runs fit
    const int nd_max = min(nd_force, dlim - d0 + 1);
    int nd = 0;
    int64_t tot = 0, first = 0;
    bool stop = false;
    for (int jb = 0; jb < nd_max && !stop; jb += 4) {
      int64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      for (int64_t t0 = 0; t0 < A.m; t0 += 32) {
        const int64_t t = t0 + lane;
        if (t < A.m) {
          const int64_t g = (int64_t)__ldg(A.occ_a + t);
          const int64_t lim = (int64_t)__ldg(A.unit_last + g) + 1;
          const int64_t u = g + d0 + jb;
          const int64_t p0 = __ldg(A.unit_ptr + min(u, lim)), p1 = __ldg(A.unit_ptr + min(u + 1, lim)),
                        p2 = __ldg(A.unit_ptr + min(u + 2, lim)), p3 = __ldg(A.unit_ptr + min(u + 3, lim)),
                        p4 = __ldg(A.unit_ptr + min(u + 4, lim));
          c0 += p1 - p0; c1 += p2 - p1; c2 += p3 - p2; c3 += p4 - p3;
        }
      }
      const int64_t c[4] = {warp_sum_i64(c0), warp_sum_i64(c1), warp_sum_i64(c2), warp_sum_i64(c3)};
      if (jb == 0) first = c[0];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (stop || jb + k >= nd_max) { stop = true; continue; }
        const int64_t nt = tot + c[k];
        if (nt <= cap && min(nt, A.m * (int64_t)(jb + k + 1)) <= cnt_limit) { tot = nt; ++nd; } else stop = true;
      }
    }
    if (nd < 1) { nd = 1; tot = first; }
    if (tot == 0) { d0 += nd; nd_force = 31; continue; }
    const int d1 = d0 + nd - 1;
    // A single distance too large for one table is cut by id range.  Octants of the id space have
    // precomputed split positions in every unit list (usplit); finer cuts use binary search.
    int64_t parts = 1;
    if (nd == 1 && tot > cap) parts = (tot + cap - 1) / cap;
    int pw = (parts <= 1) ? 8 : (parts <= 2) ? 4 : (parts <= 4) ? 2 : (parts <= 8) ? 1 : 0;  // octants per pass, 0 = free widths
    int64_t width = (pw > 0) ? 0 : max((int64_t)1, A.n_kmers / parts);
    bool redo = false;
    int64_t lo_id = 0;
    int jl = 0;
    while (lo_id < A.n_kmers) {
      int64_t hi_id;
      int j_lo = -1, j_hi = -1;
      if (pw > 0) {
        j_lo = jl;
        j_hi = min(8, jl + pw);
        hi_id = (j_hi == 8) ? A.n_kmers : (A.n_kmers * j_hi) >> 3;
      } else {
        hi_id = min(A.n_kmers, lo_id + width);
      }
      const int distinct = (hi_id > lo_id) ? pair_chunk_pass<S>(tbase, A, d0, d1, lo_id, hi_id, j_lo, j_hi, cb) : 0;
      if (distinct < 0) {  // too many distinct ids for one table: smaller chunk, then smaller id ranges
        ++splits;
        ratio = 1.0f;
        if (nd > 1) { nd_force = nd >> 1; redo = true; break; }
        if (pw > 1) { pw >>= 1; continue; }
        if (pw == 1) { pw = 0; width = max((int64_t)1, (hi_id - lo_id) >> 1); continue; }
        width = max((int64_t)1, width >> 1);
        continue;
      }
      if (pw == 8) ratio = fminf(1.0f, fmaxf(0.05f, (float)distinct / (float)tot));
      lo_id = hi_id;
      jl = j_hi;
    }
    if (redo) continue;
    d0 += nd;
    nd_force = 31;
  }
}

// index of the last unit of the read holding unit g = occurrence t of the source: from the per-occurrence copy
// (occ_last, contiguous like occ itself) when the caller built one, else through unit_last[g] (a dependent random load)
__device__ __forceinline__ int64_t last_unit_of(const uint32_t* __restrict__ unit_last, const uint32_t* __restrict__ last_a,
                                                int64_t t, int64_t g) {
  return (int64_t)(last_a ? __ldg(last_a + t) : __ldg(unit_last + g));
}

// Per-source prologue shared by both stage-C kernels.  Adds the reference's number of `+= 1` executions for
// this source (dbkr.py:126) to incr_total, in closed form: every id of every unit g + d, d in [dmin, max_d],
// minus the occurrences of a itself in them.  Returns the largest distance worth streaming (exact pruning:
// cnt[d][a][b] >= min_cov needs >= min_cov occurrences whose read still has a unit g + d), dmin - 1 if none.
__device__ int source_scope(const int64_t* __restrict__ unit_ptr, const uint32_t* __restrict__ unit_last,
                            const uint32_t* __restrict__ occ_a, int64_t m, int dmin, int max_d, uint32_t min_cov,
                            int64_t& incr_total, const uint32_t* __restrict__ last_a = nullptr) {
  const int lane = threadIdx.x & 31;
  int64_t entries = 0, self = 0, maxrem = 0;
  for (int64_t t0 = 0; t0 < m; t0 += 32) {
    const int64_t t = t0 + lane;
    if (t < m) {
      const int64_t g = (int64_t)__ldg(occ_a + t);
      const int64_t last = last_unit_of(unit_last, last_a, t, g);
      const int64_t lo = g + dmin, hi = min(g + (int64_t)max_d, last);
      if (lo <= hi) {
        entries += __ldg(unit_ptr + hi + 1) - __ldg(unit_ptr + lo);
        self += lower_bound_u32(occ_a, t + 1, m, hi + 1) - lower_bound_u32(occ_a, t + 1, m, lo);
      }
      maxrem = max(maxrem, last - g);
    }
  }
  for (int o = 16; o >= 1; o >>= 1) {
    entries += __shfl_xor_sync(FULL, entries, o);
    self += __shfl_xor_sync(FULL, self, o);
    maxrem = max(maxrem, __shfl_xor_sync(FULL, maxrem, o));
  }
  incr_total += entries - self;
  int dlim = (int)min((int64_t)max_d, maxrem);
  if (min_cov > 1) {
    if ((uint64_t)m < (uint64_t)min_cov) return dmin - 1;
    // largest d with at least min_cov occurrences whose read still has a unit g + d
    int lo_d = dmin - 1, hi_d = dlim;  // invariant: count(lo_d) >= min_cov or lo_d == dmin - 1
    while (lo_d < hi_d) {
      const int mid = (lo_d + hi_d + 1) >> 1;
      int64_t c = 0;
      for (int64_t t0 = 0; t0 < m; t0 += 32) {
        const int64_t t = t0 + lane;
        bool ge = false;
        if (t < m) {
          const int64_t g = (int64_t)__ldg(occ_a + t);
          ge = last_unit_of(unit_last, last_a, t, g) - g >= mid;
        }
        c += __popc(__ballot_sync(FULL, ge));
      }
      if (c >= (int64_t)min_cov) lo_d = mid; else hi_d = mid - 1;
    }
    dlim = lo_d;
  }
  return dlim;
}

__global__ void __launch_bounds__(PC_WARPS * 32, 1)
pair_candidates_kernel(const int64_t* __restrict__ unit_ptr, const uint32_t* __restrict__ ids,
                       const uint32_t* __restrict__ unit_last, const int64_t* __restrict__ occ_ptr,
                       const uint32_t* __restrict__ occ, const uint32_t* __restrict__ usplit, int64_t n_kmers,
                       int64_t a_begin, int64_t a_end, int32_t a_stride, int32_t min_d, int32_t max_d,
                       uint32_t min_cov, uint4* cand, int64_t max_cand, int64_t* counters) {
  extern __shared__ __align__(16) unsigned char pc_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(pc_smem) + (uint32_t)warp * PC_TBL_BYTES;
  const int dmin = max(min_d, 1);
  min_cov = max(min_cov, 1u);  // the reference only ever looks at counters that exist, i.e. are >= 1 (dbkr.py:133-136)
  const int kb = 64 - __clzll((unsigned long long)n_kmers);  // bits of the largest key b + 1 = n_kmers
  int64_t incr_total = 0, splits = 0;
  float ratio = 1.0f;
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd((unsigned long long*)(counters + 1), 1ull);
    item = __shfl_sync(FULL, item, 0);
    const int64_t a64 = a_begin + (int64_t)item * a_stride;
    if (a64 >= a_end) break;
    const uint32_t a = (uint32_t)a64;
    const int64_t o0 = occ_ptr[a], m = occ_ptr[a + 1] - o0;
    if (m == 0) continue;
    const uint32_t* occ_a = occ + o0;
    const int dlim = source_scope(unit_ptr, unit_last, occ_a, m, dmin, max_d, min_cov, incr_total);
    if (dlim < dmin) continue;
    PairArgs A{unit_ptr, ids, unit_last, occ_a, usplit, m, n_kmers, a, min_cov, cand, max_cand, counters};
    const int mb = 64 - __clzll((unsigned long long)m);  // a single-distance count never exceeds m
    if (kb + mb <= 32 && (32 - kb) >= 8)
      pair_source<uint32_t>(tbase, A, dmin, dlim, 32 - kb, ratio, splits);
    else
      pair_source<uint64_t>(tbase, A, dmin, dlim, 32, ratio, splits);
  }
  if (lane == 0) {
    if (incr_total) atomicAdd((unsigned long long*)(counters + 2), (unsigned long long)incr_total);
    if (splits) atomicAdd((unsigned long long*)(counters + 3), (unsigned long long)splits);
  }
}

// ---- stage D core (pair_join_kernel, one thread per candidate).  All 32 lanes of a warp call it; `active` lanes hold a
// candidate (a, b, d0, d1).  (Running it inside the sketch kernel -- one lane per parked candidate of the warp's source,
// no candidate array -- was measured and dropped: 22.0 ms against 16.7 + 2.8 ms, the sketch kernel paid 16 more
// registers per thread.)
// Joins the sorted occurrence lists of a and b: all_occ = #{(g, g') : g' - g in [dmin, dmax], same read}
// (dbkr.py:143) and the exact cnt[d][a][b] for the distances of the chunk; keeps (a, b, d, cnt) iff cnt >= min_cov and
// (double)cnt / (double)all_occ >= rel_threshold (dbkr.py:136,145).  Returns the lane's number of (d) with cnt >= min_cov.
struct JoinArgs {
  const int64_t* __restrict__ occ_ptr;
  const uint32_t* __restrict__ occ;
  const uint32_t* __restrict__ occ_last;   // may be nullptr
  const uint32_t* __restrict__ unit_last;
  uint32_t dmin, dmax, min_cov;
  double rel_threshold;
  uint4* edges;
  int64_t max_edges;
  uint8_t* selected;
  int64_t* edge_counter;
};

__device__ __forceinline__ uint32_t join_candidate(bool active, uint4 c, const JoinArgs& J) {
  // 32-bit cursors into occ[] (its length is the number of cloud entries, < 2^32) and unit distances relative to g,
  // the current unit of b kept in a register (0xFFFFFFFF behind the end of the list: no unit has that index)
  constexpr uint32_t END = 0xFFFFFFFFu;
  uint32_t dmask = 0;  // distances d0 + j of the chunk at which the pair co-occurs
  uint32_t cnt0 = 0;   // cnt[d0][a][b], gathered by the first join (most chunks are a single distance)
  uint64_t all_occ = 0;
  uint32_t ia = 0, ia_end = 0, ib0 = 0, ib_end = 0;
  if (active) {
    ia = (uint32_t)J.occ_ptr[c.x]; ia_end = (uint32_t)J.occ_ptr[c.x + 1];
    ib0 = (uint32_t)J.occ_ptr[c.y]; ib_end = (uint32_t)J.occ_ptr[c.y + 1];
    const uint32_t width = c.w - c.z;
    uint32_t ib = ib0;
    uint32_t bj = ib < ib_end ? __ldg(J.occ + ib) : END;
    for (uint32_t t = ia; t < ia_end; ++t) {
      const uint32_t g = __ldg(J.occ + t);
      const uint32_t rem = (J.occ_last ? __ldg(J.occ_last + t) : __ldg(J.unit_last + g)) - g;  // units behind g in its read
      if (J.dmin > rem) continue;
      const uint32_t lo = g + J.dmin, hi = g + min(J.dmax, rem);  // <= last unit of the read: no wrap
      while (bj < lo) {
        ++ib;
        bj = ib < ib_end ? __ldg(J.occ + ib) : END;
      }
      uint32_t jj = ib, v = bj;
      while (v <= hi) {  // END > hi always
        ++all_occ;
        const uint32_t dj = (v - g) - c.z;  // wraps to a huge value below d0
        if (dj <= width) dmask |= 1u << dj;
        cnt0 += dj == 0u;
        ++jj;
        v = jj < ib_end ? __ldg(J.occ + jj) : END;
      }
    }
  }
  uint32_t n_cand_d = 0;
  while (__any_sync(FULL, dmask != 0)) {
    bool keep = false;
    uint32_t d = 0, cnt = 0;
    if (dmask) {
      const uint32_t dj = (uint32_t)(__ffs(dmask) - 1);
      d = c.z + dj;
      dmask &= dmask - 1;
      if (dj == 0u) {
        cnt = cnt0;
      } else {
        uint32_t ib = ib0;
        uint32_t bj = ib < ib_end ? __ldg(J.occ + ib) : END;
        for (uint32_t t = ia; t < ia_end; ++t) {  // cnt[d][a][b]: occurrences g of a with g + d holding b inside the read
          const uint32_t g = __ldg(J.occ + t);
          const uint32_t rem = (J.occ_last ? __ldg(J.occ_last + t) : __ldg(J.unit_last + g)) - g;
          if (d > rem) continue;
          const uint32_t want = g + d;
          while (bj < want) {
            ++ib;
            bj = ib < ib_end ? __ldg(J.occ + ib) : END;
          }
          cnt += bj == want;
        }
      }
      if (cnt >= J.min_cov) {
        ++n_cand_d;
        keep = ((double)cnt / (double)all_occ) >= J.rel_threshold;
      }
    }
    const int64_t pos = warp_append(keep, J.edge_counter);
    if (keep) {
      if (pos < J.max_edges) J.edges[pos] = make_uint4(c.x, c.y, d, cnt);
      J.selected[c.x] = 1;
      J.selected[c.y] = 1;
    }
  }
  return n_cand_d;
}

// ============================================================================================
// Stage C, sketch form (the default for 3 <= min_cov <= 255).
//
// The exact tables above spend ~6 instructions and a probe chain per pair increment, although
// > 99 % of the increments go to counters that never reach min_cov.  Here one pass over the
// distances [d0, d0 + nd) of a source a only answers "can b reach min_cov?":
//
//   level 1  a warp-private array of 2^SK_BITS saturating BYTE counters indexed by a hash of b.
//            counter[h] counts the units of the pass holding some id with hash h, which is
//            >= the number of units holding b itself: an upper bound of sum_d cnt[d][a][b].
//            The hash and a "same hash as an earlier id of this unit" flag are precomputed per
//            cloud entry (codes[], 2 bytes per entry, cfk_sketch_codes): flagged entries do not
//            store, so the lanes of one step always write DISTINCT bytes -- plain ld/st, no
//            atomics, no keys, no probe chains.
//   level 2  an entry whose counter is already >= min_cov - 1 is "hot": its position is queued
//            and, 32 at a time, the real id is fetched and put into a small exact hash SET.
//            Every b with sum_d cnt[d][a][b] >= min_cov is hot at its min_cov-th unit at the
//            latest, so the set is a superset of the true candidates (hash collisions only add
//            members); the set is emitted as (a, b, d0, d1) once the pass is complete.
//
// cfk_pair_join then computes the exact per-distance counts of those pairs, exactly as for the
// exact tables.  A pass whose set overflows is abandoned before anything is emitted and redone
// over fewer distances / a narrower id range.
// ============================================================================================
constexpr int SK_BITS = CFK_SKETCH_BITS;
constexpr int SK_TBL_BYTES = 1 << SK_BITS;
constexpr uint32_t SK_OFF_MASK = 0x3FFFu;          // code bits 0..13: byte offset in the warp's table (hash, or the null byte)
constexpr uint32_t SK_DUP = 0x8000u;               // code bit 15: do not store (duplicate hash inside the unit, or null)
constexpr uint32_t SK_NULL = SK_DUP | (uint32_t)SK_TBL_BYTES;  // padding entry: reads the always-zero byte behind the table
constexpr int SK_PAD_BYTES = 128;                  // always-zero words behind the table, one per bank: null entries of lane l read word l
constexpr int SK_L2_SLOTS = 256;   // level-2 set, u32 slots holding b + 1
constexpr int SK_L2_MAX = 192;
constexpr int SK_Q_SLOTS = 256;    // queue of hot positions
constexpr int SK_WINDOW = 128;     // cloud entries per step (4 per lane) = one block of codes[]
constexpr int SK_WARP_BYTES = SK_TBL_BYTES + SK_PAD_BYTES + SK_L2_SLOTS * 4 + SK_Q_SLOTS * 4;
constexpr int SK_WARPS_FIT = (227 * 1024 - 256) / SK_WARP_BYTES;
constexpr int SK_WARPS = SK_WARPS_FIT > 32 ? 32 : SK_WARPS_FIT;
constexpr uint32_t SK_CAP = (uint32_t)SK_TBL_BYTES / CFK_SKETCH_LOAD_DIV;  // planned cloud entries per pass
constexpr int SK_ND_MAX = 16;
static_assert(SK_BITS >= 10 && SK_BITS <= 13, "codes are 16 bit: offset in bits 0..13 (the null byte sits at 2^SK_BITS), bit 15 = no store");
static_assert(SK_WARPS >= 1, "sketch does not fit shared memory");

// codes[] layout (private to the two sketch kernels): blocks of SK_WINDOW codes; unit u owns the blocks
// sk_block(unit_ptr[u], u) .. + ceil(|C_u| / SK_WINDOW) - 1 (closed form, no scan: consecutive units are at least
// that far apart).  Inside a block the entry 32 * j + l of the window sits at 4 * l + j, so lane l fetches its four
// entries -- one per row j -- with ONE 8-byte load; slots behind the end of the unit hold SK_NULL.
__host__ __device__ __forceinline__ int64_t sk_block(int64_t first_entry, int64_t unit) { return (first_entry >> 7) + unit; }
static_assert(SK_WINDOW == 128, "sk_block shifts by 7");

__device__ __forceinline__ uint32_t sk_hash(uint32_t id) { return (id * 0x9E3779B1u) >> (32 - SK_BITS); }

__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t atom_cas_shared(uint32_t a, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(a), "r"(cmp), "r"(val) : "memory");
  return old;
}

// Per cloud entry: hash of the id | SK_DUP if another id of the same unit with the same hash was served first, written
// in the blocked layout above.  Warp per unit.
__global__ void __launch_bounds__(256) sketch_codes_kernel(const int64_t* __restrict__ unit_ptr, const uint32_t* __restrict__ ids,
                                                           int64_t n_units, uint2* __restrict__ codes) {
  __shared__ uint32_t bm[8][SK_TBL_BYTES / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t u = (int64_t)blockIdx.x * 8 + warp;
  if (u >= n_units) return;
  const int64_t b = unit_ptr[u], e = unit_ptr[u + 1];
  if (b == e) return;
  for (int i = lane; i < SK_TBL_BYTES / 32; i += 32) bm[warp][i] = 0;
  __syncwarp();
  uint2* out = codes + sk_block(b, u) * 32 + lane;
  for (int64_t p0 = b; p0 < e; p0 += SK_WINDOW, out += 32) {
    uint32_t c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t p = p0 + 32 * j + lane;
      c[j] = SK_NULL;
      if (p < e) {
        const uint32_t h = sk_hash(ids[p]);
        const uint32_t old = atomicOr(&bm[warp][h >> 5], 1u << (h & 31u));
        c[j] = h | (((old >> (h & 31u)) & 1u) ? SK_DUP : 0u);
      }
    }
    *out = make_uint2(c[0] | (c[1] << 16), c[2] | (c[3] << 16));
  }
}

// Bank-aware form of sketch_codes_kernel.  The byte counters of a pass are hit at random, so a row of 32 entries
// costs ~2 shared-memory wavefronts for the load and again for the store (bank conflicts).  Here the entries of a
// 128-entry window are PLACED: the r-th entry of the window whose counter lies in bank B goes to row r, lane B, so the
// rows hold at most one entry per bank -- conflict-free -- as long as a bank has no more entries than the window has
// rows; the overflow (about one entry in seven) fills the free slots, highest rows first, so that the conflicts
// concentrate in the last row.  Free slots of lane l hold a null code that reads the always-zero pad word of bank l.
// Because the order inside a window changes, the ids travel with the codes: perm_ids[block * 128 + 4 * lane + row]
// is the id behind the code at that slot (the hot-id lookup of pair_sketch_kernel reads it instead of ids[]).
__global__ void __launch_bounds__(256) sketch_codes_banked_kernel(const int64_t* __restrict__ unit_ptr,
                                                                  const uint32_t* __restrict__ ids, int64_t n_units,
                                                                  uint2* __restrict__ codes, uint4* __restrict__ perm_ids) {
  __shared__ uint32_t bm[8][SK_TBL_BYTES / 32];
  __shared__ int s_cnt[8][32];
  __shared__ int s_nover[8];
  __shared__ uint16_t s_code[8][SK_WINDOW], s_ocode[8][SK_WINDOW];
  __shared__ uint32_t s_id[8][SK_WINDOW], s_oid[8][SK_WINDOW];
  __shared__ uint8_t s_free[8][SK_WINDOW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const int64_t u = (int64_t)blockIdx.x * 8 + warp;
  if (u >= n_units) return;
  const int64_t b = unit_ptr[u], e = unit_ptr[u + 1];
  if (b == e) return;
  for (int i = lane; i < SK_TBL_BYTES / 32; i += 32) bm[warp][i] = 0;
  __syncwarp();
  int64_t blk = sk_block(b, u);
  for (int64_t p0 = b; p0 < e; p0 += SK_WINDOW, ++blk) {
    const int n = (int)min((int64_t)SK_WINDOW, e - p0), rows = (n + 31) >> 5;
    s_cnt[warp][lane] = 0;
    if (lane == 0) s_nover[warp] = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // slot 4 * l + j of lane l: null until an entry is placed there
      s_code[warp][4 * lane + j] = (uint16_t)(SK_DUP | (uint32_t)(SK_TBL_BYTES + 4 * lane));
      s_id[warp][4 * lane + j] = 0u;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = 32 * j + lane;
      if (q < n) {
        const uint32_t id = ids[p0 + q];
        const uint32_t h = sk_hash(id);
        const uint32_t old = atomicOr(&bm[warp][h >> 5], 1u << (h & 31u));
        const uint16_t code = (uint16_t)(h | (((old >> (h & 31u)) & 1u) ? SK_DUP : 0u));
        const int bank = (int)((h >> 2) & 31u);
        const int r = atomicAdd(&s_cnt[warp][bank], 1);
        if (r < rows) {
          s_code[warp][4 * bank + r] = code;
          s_id[warp][4 * bank + r] = id;
        } else {
          const int o = atomicAdd(&s_nover[warp], 1);
          s_ocode[warp][o] = code;
          s_oid[warp][o] = id;
        }
      }
    }
    __syncwarp();
    // free slots of the rows in use, highest row first
    const int mine = min(s_cnt[warp][lane], rows);
    int base = 0;
    for (int r = rows - 1; r >= 0; --r) {
      const bool fr = mine <= r;
      const unsigned bal = __ballot_sync(FULL, fr);
      if (fr) s_free[warp][base + __popc(bal & lt)] = (uint8_t)(4 * lane + r);
      base += __popc(bal);
    }
    __syncwarp();
    const int nover = s_nover[warp];  // <= base: the rows in use hold 32 * rows >= n slots
    for (int o = lane; o < nover; o += 32) {
      const int slot = s_free[warp][o];
      s_code[warp][slot] = s_ocode[warp][o];
      s_id[warp][slot] = s_oid[warp][o];
    }
    __syncwarp();
    const uint32_t c0 = s_code[warp][4 * lane], c1 = s_code[warp][4 * lane + 1], c2 = s_code[warp][4 * lane + 2],
                   c3 = s_code[warp][4 * lane + 3];
    codes[blk * 32 + lane] = make_uint2(c0 | (c1 << 16), c2 | (c3 << 16));
    perm_ids[blk * 32 + lane] = make_uint4(s_id[warp][4 * lane], s_id[warp][4 * lane + 1], s_id[warp][4 * lane + 2],
                                           s_id[warp][4 * lane + 3]);
    __syncwarp();
  }
}

struct SketchArgs {
  const int64_t* __restrict__ unit_ptr;
  const uint32_t* __restrict__ ids;      // hot-id lookup: ids[] positions, or (banked codes) perm_ids[] slots
  bool banked;
  const uint2* __restrict__ codes;
  const uint32_t* __restrict__ unit_last;
  const uint32_t* __restrict__ last_a;  // unit_last[occ_a[t]] per occurrence, or nullptr
  const uint32_t* __restrict__ occ_a;
  int64_t m;
  int64_t n_kmers;
  uint32_t a;
  uint32_t thr;  // min_cov - 1, 2 <= thr <= 254
  uint4* cand;
  int64_t max_cand;
  int64_t* counters;
  struct SketchChunk* chunk;  // the warp's current chunk of cand[]
  bool prefetch;              // ask the code blocks of a pass's units into L2 up front (graphs larger than L2)
};

constexpr int SK_CAND_CHUNK = 256;
constexpr uint32_t SK_NO_CAND = 0xFFFFFFFFu;  // a = 0xFFFFFFFF: hole at the end of a chunk (cfk_pair_join skips it)
struct SketchChunk {
  int64_t pos, end, emitted;
};

// the unused tail of the warp's chunk becomes holes
__device__ __forceinline__ void sk_close_chunk(const SketchArgs& A) {
  const int lane = threadIdx.x & 31;
  for (int64_t q = A.chunk->pos + lane; q < A.chunk->end; q += 32)
    if (q < A.max_cand) A.cand[q] = make_uint4(SK_NO_CAND, 0u, 0u, 0u);
  A.chunk->pos = A.chunk->end;
}

// queued hot positions -> level-2 set.  Returns false if the set overflowed.
__device__ bool sk_drain(uint32_t l2base, uint32_t qbase, const SketchArgs& A, int qn, int64_t lo_id, int64_t hi_id,
                         int& claims) {
  const int lane = threadIdx.x & 31;
  bool ovf = false;
  __syncwarp();
  for (int base = 0; base < qn; base += 32) {
    const int idx = base + lane;
    if (idx < qn) {
      const uint32_t pos = lds<uint32_t>(qbase + (uint32_t)idx * 4u);
      const uint32_t b = __ldg(A.ids + pos);
      if (b != A.a && (int64_t)b >= lo_id && (int64_t)b < hi_id) {
        const uint32_t key = b + 1u;
        uint32_t slot = (b * 0x85EBCA77u) >> 24;
        static_assert(SK_L2_SLOTS == 256, "slot hash takes the top 8 bits");
        int probes = 0;
        for (; probes < SK_L2_SLOTS; ++probes) {
          const uint32_t old = atom_cas_shared(l2base + slot * 4u, 0u, key);
          if (old == 0u) { ++claims; break; }
          if (old == key) break;
          slot = (slot + 1u) & (uint32_t)(SK_L2_SLOTS - 1);
        }
        if (probes == SK_L2_SLOTS) ovf = true;
      }
    }
  }
  __syncwarp();
  const int total = __reduce_add_sync(FULL, claims);
  return !(__any_sync(FULL, ovf) || total > SK_L2_MAX);
}

// One pass: distances d0 .. d0 + nd - 1, hot ids restricted to [lo_id, hi_id).  Returns the number
// of candidates emitted, or -1 if the level-2 set overflowed (nothing emitted).
// pre != nullptr (sources with at most 32 occurrences, one distance per pass): lane t already holds its unit g_t + d0 as
// (first entry, end entry, unit index) -- sketch_source keeps a sliding window of unit_ptr values in registers, so the
// pass starts without a single dependent load.
struct SketchPre {
  uint32_t up, ue;
  int64_t unit;
};

// The code blocks of the units a pass is about to walk (one unit per lane) are asked into L2 up front: on one GPU the
// whole cloud CSR sits in L2 anyway, on N GPUs the all-gathered CSR is N times larger and every window would otherwise
// wait for DRAM behind a one-window look-ahead.  A block of 128 codes = 256 B = two lines.
__device__ __forceinline__ void sk_prefetch_unit(const uint2* codes, uint32_t up, uint32_t ue, uint32_t ub) {
  if (ue <= up) return;
  const uint32_t n_blocks = (ue - up + (uint32_t)SK_WINDOW - 1u) / (uint32_t)SK_WINDOW;
  const char* base = reinterpret_cast<const char*>(codes + (size_t)ub * 32u);
  for (uint32_t b = 0; b < n_blocks && b < 4u; ++b) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + b * 256u));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + b * 256u + 128u));
  }
}

__device__ int sketch_pass(uint32_t tbase, const SketchArgs& A, int d0, int nd, int64_t lo_id, int64_t hi_id,
                           const SketchPre* pre = nullptr) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t l2base = tbase + SK_TBL_BYTES + SK_PAD_BYTES, qbase = l2base + SK_L2_SLOTS * 4;
  const uint32_t thr = A.thr;
  uint32_t up = 0, ue = 0, ub = 0;  // this lane's unit: entries [up, ue) of ids[], first block of codes[]
  unsigned rest = 0;
  int src = 0;
  uint32_t p = 0, e = 0, cb = 0;
  uint2 w = make_uint2(0, 0);
  if (pre != nullptr) {  // the units are known: the first code block is requested before the table is cleared
    up = pre->up;
    ue = pre->ue;
    ub = (uint32_t)sk_block((int64_t)up, pre->unit);
    if (A.prefetch) sk_prefetch_unit(A.codes, up, ue, ub);
    rest = __ballot_sync(FULL, ue > up);
    if (rest) {
      src = __ffs(rest) - 1;
      rest &= rest - 1;
      p = __shfl_sync(FULL, up, src); e = __shfl_sync(FULL, ue, src); cb = __shfl_sync(FULL, ub, src);
      w = __ldg(A.codes + (size_t)cb * 32u + lane);
    }
  }
#pragma unroll 4
  for (int i = lane; i < (SK_TBL_BYTES + SK_PAD_BYTES + SK_L2_SLOTS * 4) / 16; i += 32)
    sts128(tbase + (uint32_t)i * 16u, make_uint4(0, 0, 0, 0));
  __syncwarp();
  int qn = 0, claims = 0;
  const int64_t n_items = A.m * (int64_t)nd;
  for (int64_t i0 = 0; i0 < n_items; i0 += 32) {
    if (pre != nullptr) {
      if (e <= p) continue;  // no occurrence has a unit at this distance (one group only: m <= 32)
    } else {
      const int64_t it = i0 + lane;
      up = 0; ue = 0; ub = 0;
      if (it < n_items) {
        const int64_t t = (nd == 1) ? it : it / nd;
        const int64_t g = (int64_t)__ldg(A.occ_a + t);
        const int64_t u = g + d0 + (it - t * nd);
        if (u <= last_unit_of(A.unit_last, A.last_a, t, g)) {
          const int64_t p0 = __ldg(A.unit_ptr + u);
          up = (uint32_t)p0;
          ue = (uint32_t)__ldg(A.unit_ptr + u + 1);
          ub = (uint32_t)sk_block(p0, u);
        }
      }
      if (A.prefetch) sk_prefetch_unit(A.codes, up, ue, ub);
      rest = __ballot_sync(FULL, ue > up);
      if (!rest) continue;
      src = __ffs(rest) - 1;
      rest &= rest - 1;
      p = __shfl_sync(FULL, up, src); e = __shfl_sync(FULL, ue, src); cb = __shfl_sync(FULL, ub, src);
      w = __ldg(A.codes + (size_t)cb * 32u + lane);
    }
    for (;;) {
      // the next window (rest of this unit, else the next unit of the group): its codes fly during this step
      uint32_t np = p + SK_WINDOW, ne = e, ncb = cb + 1u;
      // branch-free: the next unit's descriptor is fetched whether or not this unit is finished (it is, four steps out
      // of five), so that the bookkeeping shares one basic block with the rows below
      const bool sw = np >= ne;  // warp-uniform
      const int nsrc = (__ffs(rest) - 1) & 31;
      const uint32_t sp = __shfl_sync(FULL, up, nsrc), se = __shfl_sync(FULL, ue, nsrc), sb = __shfl_sync(FULL, ub, nsrc);
      const bool more = !sw || rest != 0u;
      np = sw ? sp : np;
      ne = sw ? se : ne;
      ncb = sw ? sb : ncb;
      rest = sw ? (rest & (rest - 1u)) : rest;
      uint2 nw = make_uint2(0, 0);
      if (more) nw = __ldg(A.codes + (size_t)ncb * 32u + lane);
      // level 1: four rows of 32 entries.  Non-flagged entries of one unit have distinct hashes, so all stores of a step
      // hit distinct bytes: plain ld/st.  All four rows run unconditionally -- the slots behind the end of a unit hold
      // null codes (they read an always-zero pad word and do not store) -- because without per-row branches the four
      // load / add / store chains are one basic block that the scheduler interleaves: with 5.5 warps per scheduler the
      // fixed ALU latencies were the top stall (15.96 -> 15.08 ms; skipping the unused rows saved wavefronts, not time)
      const uint32_t a0 = tbase + (w.x & SK_OFF_MASK), a1 = tbase + ((w.x >> 16) & SK_OFF_MASK);
      const uint32_t a2 = tbase + (w.y & SK_OFF_MASK), a3 = tbase + ((w.y >> 16) & SK_OFF_MASK);
      const uint32_t o0 = lds_u8(a0), o1 = lds_u8(a1), o2 = lds_u8(a2), o3 = lds_u8(a3);
      if (!(w.x & SK_DUP)) sts_u8(a0, min(o0 + 1u, 255u));
      if ((int32_t)w.x >= 0) sts_u8(a1, min(o1 + 1u, 255u));
      if (!(w.y & SK_DUP)) sts_u8(a2, min(o2 + 1u, 255u));
      if ((int32_t)w.y >= 0) sts_u8(a3, min(o3 + 1u, 255u));
      // one test for the common case "nothing hot in this step"
      if (__any_sync(FULL, max(max(o0, o1), max(o2, o3)) >= thr)) {
        const bool hs[4] = {o0 >= thr, o1 >= thr, o2 >= thr, o3 >= thr};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const unsigned bal = __ballot_sync(FULL, hs[i]);
          if (hs[i]) sts(qbase + (uint32_t)(qn + __popc(bal & lt)) * 4u, A.banked ? (cb << 7) + 4u * (uint32_t)lane + (uint32_t)i : p + (uint32_t)lane + 32u * i);
          qn += __popc(bal);
        }
        if (qn > SK_Q_SLOTS - SK_WINDOW) {
          if (!sk_drain(l2base, qbase, A, qn, lo_id, hi_id, claims)) return -1;
          qn = 0;
        }
      }
      __syncwarp();
      if (!more) break;
      p = np; e = ne; cb = ncb; w = nw;
    }
  }
  if (!sk_drain(l2base, qbase, A, qn, lo_id, hi_id, claims)) return -1;
  // emit the set
  int emitted = 0;
  for (int i = lane; i < SK_L2_SLOTS / 4; i += 32) {
    const uint4 q = lds128(l2base + (uint32_t)i * 16u);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    if (__any_sync(FULL, (q.x | q.y | q.z | q.w) != 0u)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool take = w[j] != 0u;
        const unsigned bal = __ballot_sync(FULL, take);
        if (bal) {
          // the warp owns a chunk of the candidate array and fills it without atomics; one atomicAdd on the (single,
          // contended) cursor per SK_CAND_CHUNK candidates instead of one round trip per pass
          const int n = __popc(bal);
          if (A.chunk->pos + n > A.chunk->end) {
            sk_close_chunk(A);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd((unsigned long long*)A.counters, (unsigned long long)SK_CAND_CHUNK);
            base = __shfl_sync(FULL, base, 0);
            A.chunk->pos = (int64_t)base;
            A.chunk->end = (int64_t)base + SK_CAND_CHUNK;
          }
          const int64_t pos = A.chunk->pos + __popc(bal & lt);
          if (take && pos < A.max_cand) A.cand[pos] = make_uint4(A.a, w[j] - 1u, (uint32_t)d0, (uint32_t)(d0 + nd - 1));
          A.chunk->pos += n;
          A.chunk->emitted += n;
        }
        emitted += take;
      }
    }
  }
  __syncwarp();
  return emitted;
}

__device__ void sketch_source(uint32_t tbase, const SketchArgs& A, int dmin, int dlim, int64_t& splits) {
  const int lane = threadIdx.x & 31;
  int d0 = dmin;
  int nd_force = SK_ND_MAX;
  // Sources with at most 32 occurrences (nearly all: the rare band caps the reads per k-mer): lane t keeps occurrence t
  // -- its unit g_t and the end of its read -- and a sliding window W[i] = unit_ptr[min(g_t + d0 + i, last_t + 1)],
  // i = 0..4, in registers.  Planning a pass (how many distances fit the sketch) and starting it then need no load at
  // all; the one value that enters the window per distance is requested four distances before it is used.
  const bool cached = A.m <= 32;
  int64_t cg = 0, clim = 0;
  uint32_t W0 = 0, W1 = 0, W2 = 0, W3 = 0, W4 = 0;
  auto window_at = [&](int d) -> uint32_t { return lane < A.m ? (uint32_t)__ldg(A.unit_ptr + min(cg + d, clim)) : 0u; };
  if (cached) {
    if (lane < A.m) {
      cg = (int64_t)__ldg(A.occ_a + lane);
      clim = last_unit_of(A.unit_last, A.last_a, lane, cg) + 1;
    }
    W0 = window_at(d0); W1 = window_at(d0 + 1); W2 = window_at(d0 + 2); W3 = window_at(d0 + 3); W4 = window_at(d0 + 4);
  }
  auto advance_d = [&](int by) {
    for (int i = 0; i < by; ++i) {
      ++d0;
      if (cached) {
        W0 = W1; W1 = W2; W2 = W3; W3 = W4;
        W4 = window_at(d0 + 4);
      }
    }
  };
  while (d0 <= dlim) {
    // plan: extend the pass one distance at a time (4 looked up per round) while the cloud entries fit SK_CAP
    const int nd_max = min(nd_force, dlim - d0 + 1);
    int nd = 0;
    uint32_t tot = 0, first = 0;
    bool stop = false;
    for (int jb = 0; jb < nd_max && !stop; jb += 4) {
      uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      if (cached && jb == 0) {
        c0 = W1 - W0; c1 = W2 - W1; c2 = W3 - W2; c3 = W4 - W3;
      } else
      for (int64_t t0 = 0; t0 < A.m; t0 += 32) {
        const int64_t t = t0 + lane;
        if (t < A.m) {
          const int64_t g = (int64_t)__ldg(A.occ_a + t);
          const int64_t lim = last_unit_of(A.unit_last, A.last_a, t, g) + 1;
          const int64_t u = g + d0 + jb;
          const uint32_t p0 = (uint32_t)__ldg(A.unit_ptr + min(u, lim)), p1 = (uint32_t)__ldg(A.unit_ptr + min(u + 1, lim)),
                         p2 = (uint32_t)__ldg(A.unit_ptr + min(u + 2, lim)), p3 = (uint32_t)__ldg(A.unit_ptr + min(u + 3, lim)),
                         p4 = (uint32_t)__ldg(A.unit_ptr + min(u + 4, lim));
          c0 += p1 - p0; c1 += p2 - p1; c2 += p3 - p2; c3 += p4 - p3;
        }
      }
      const uint32_t c[4] = {__reduce_add_sync(FULL, c0), __reduce_add_sync(FULL, c1), __reduce_add_sync(FULL, c2),
                             __reduce_add_sync(FULL, c3)};
      if (jb == 0) first = c[0];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (stop || jb + k >= nd_max) { stop = true; continue; }
        if (c[k] <= SK_CAP && tot <= SK_CAP - c[k]) { tot += c[k]; ++nd; } else stop = true;
      }
    }
    if (nd < 1) { nd = 1; tot = first; }
    if (tot == 0) { advance_d(nd); nd_force = SK_ND_MAX; continue; }
    bool redo = false;
    int64_t lo_id = 0, width = A.n_kmers;
    const SketchPre pre{W0, W1, cg + d0};  // W0 == W1 for the lanes without a unit at this distance
    const SketchPre* prep = (cached && nd == 1) ? &pre : nullptr;
    while (lo_id < A.n_kmers) {
      const int64_t hi_id = min(A.n_kmers, lo_id + width);
      const int emitted = sketch_pass(tbase, A, d0, nd, lo_id, hi_id, prep);
      if (emitted < 0) {  // level-2 set overflow: fewer distances, then narrower id ranges
        ++splits;
        if (nd > 1) { nd_force = nd >> 1; redo = true; break; }
        width = max((int64_t)1, (hi_id - lo_id) >> 1);
        continue;
      }
      lo_id = hi_id;
      if (emitted < SK_L2_MAX / 2 && width < A.n_kmers) width <<= 1;  // sparse stretch of the id space: widen again
    }
    if (redo) continue;
    advance_d(nd);
    nd_force = SK_ND_MAX;
  }
}

__global__ void __launch_bounds__(SK_WARPS * 32, 1)
pair_sketch_kernel(const int64_t* __restrict__ unit_ptr, const uint32_t* __restrict__ ids, const uint2* __restrict__ codes,
                   const uint32_t* __restrict__ perm_ids, const uint32_t* __restrict__ unit_last, const int64_t* __restrict__ occ_ptr,
                   const uint32_t* __restrict__ occ, const uint32_t* __restrict__ occ_last, int64_t n_kmers, int64_t a_begin,
                   int64_t a_end, int32_t a_stride, int32_t min_d, int32_t max_d, uint32_t min_cov, uint4* cand,
                   int64_t max_cand, int64_t* counters, int prefetch) {
  extern __shared__ __align__(16) unsigned char pc_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(pc_smem) + (uint32_t)warp * SK_WARP_BYTES;
  const int dmin = max(min_d, 1);
  int64_t incr_total = 0, splits = 0;
  SketchChunk chunk{0, 0, 0};
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd((unsigned long long*)(counters + 1), 1ull);
    item = __shfl_sync(FULL, item, 0);
    const int64_t a64 = a_begin + (int64_t)item * a_stride;
    if (a64 >= a_end) break;
    const uint32_t a = (uint32_t)a64;
    const int64_t o0 = occ_ptr[a], m = occ_ptr[a + 1] - o0;
    if (m == 0) continue;
    const uint32_t* occ_a = occ + o0;
    const uint32_t* last_a = occ_last ? occ_last + o0 : nullptr;
    const int dlim = source_scope(unit_ptr, unit_last, occ_a, m, dmin, max_d, min_cov, incr_total, last_a);
    if (dlim < dmin) continue;
    SketchArgs A{unit_ptr, perm_ids ? perm_ids : ids, perm_ids != nullptr, codes, unit_last, last_a, occ_a, m, n_kmers, a, min_cov - 1u, cand,
                 max_cand, counters, &chunk, prefetch != 0};
    sketch_source(tbase, A, dmin, dlim, splits);
  }
  {
    SketchArgs A{};
    A.cand = cand;
    A.max_cand = max_cand;
    A.chunk = &chunk;
    sk_close_chunk(A);
  }
  if (lane == 0) {
    if (chunk.emitted) atomicAdd((unsigned long long*)(counters + 4), (unsigned long long)chunk.emitted);
    if (incr_total) atomicAdd((unsigned long long*)(counters + 2), (unsigned long long)incr_total);
    if (splits) atomicAdd((unsigned long long*)(counters + 3), (unsigned long long)splits);
  }
}

// ============================================================================================
// Stage D: one thread per pair candidate (a, b, d0, d1).  Joins the sorted occurrence lists of a
// and b: all_occ = #{(g, g') : g' - g in [dmin, dmax], same read} (dbkr.py:143) and the exact
// cnt[d][a][b] for the distances of the chunk; keeps (a, b, d, cnt) iff cnt >= min_cov and
// (double)cnt / (double)all_occ >= rel_threshold (dbkr.py:136,145).
// ============================================================================================
__global__ void pair_join_kernel(const uint4* __restrict__ cand, int64_t n_cand, const int64_t* __restrict__ occ_ptr,
                                 const uint32_t* __restrict__ occ, const uint32_t* __restrict__ occ_last,
                                 const uint32_t* __restrict__ unit_last, int32_t min_d, int32_t max_d, uint32_t min_cov,
                                 double rel_threshold, uint4* edges, int64_t max_edges, uint8_t* selected, int64_t* counters) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const JoinArgs J{occ_ptr, occ, occ_last, unit_last, (uint32_t)max(min_d, 1), (uint32_t)max(max_d, 0), min_cov, rel_threshold,
                   edges, max_edges, selected, counters};
  uint4 c = make_uint4(0xFFFFFFFFu, 0, 0, 0);
  if (i < n_cand) c = cand[i];
  uint32_t n_cand_d = join_candidate(c.x != 0xFFFFFFFFu, c, J);  // a = 0xFFFFFFFF: hole left by cfk_pair_sketch's chunked output
  for (int o = 16; o >= 1; o >>= 1) n_cand_d += __shfl_xor_sync(FULL, n_cand_d, o);
  if ((threadIdx.x & 31) == 0 && n_cand_d) atomicAdd((unsigned long long*)(counters + 2), (unsigned long long)n_cand_d);
}

__global__ void flag_indices_kernel(const uint8_t* __restrict__ flags, int64_t n, uint32_t* out, int64_t* counters) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool take = i < n && flags[i] != 0;
  int64_t pos = warp_append(take, counters);
  if (take) out[pos] = (uint32_t)i;
}


}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int cfk_abi_version(void) { return 1; }
const char* cfk_last_error(void) { return cfk_g_err; }
int cfk_pair_table_bytes_per_warp(void) { return PC_TBL_BYTES; }
int cfk_pair_warps_per_block(void) { return PC_WARPS; }
int64_t cfk_launch_count(void) { return (int64_t)__atomic_load_n(&cfk_g_launches, __ATOMIC_RELAXED); }

int cfk_table_init(uint64_t* table, int64_t cap, cfk_stream_t stream) {
  if (cap < 1) return fail(CFK_ERR_INVALID, "cfk_table_init: cap < 1");
  table_init_kernel<<<(unsigned)blocks_for(cap, 256), 256, 0, (cudaStream_t)stream>>>(table, cap);
  CFK_CHECK_LAUNCH("table_init_kernel", 1);
  return CFK_OK;
}

int cfk_docfreq_count(const uint32_t* packed, const int64_t* read_off, const int64_t* read_len, const int32_t* order,
                      int64_t n_reads, int k, uint64_t* table, int64_t cap, int64_t* counters, int32_t n_blocks,
                      cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_docfreq_count: k must be in [1, 31]");
  if (cap < 1 || n_reads < 0 || n_blocks < 1) return fail(CFK_ERR_INVALID, "cfk_docfreq_count: bad sizes");
  if (n_reads == 0) return CFK_OK;
  static unsigned long long attr_done = 0;  // per device (bit = device ordinal)
  const int smem = (DF_SET_SLOTS + DF_TILE_WORDS) * 4;
  {
    cudaError_t e = cfk::ensure_dynamic_smem(docfreq_kernel, smem, &attr_done);
    if (e != cudaSuccess) return fail(CFK_ERR_CUDA, "cfk_docfreq_count: cudaFuncSetAttribute", e);
  }
  docfreq_kernel<<<(unsigned)n_blocks, DF_THREADS, smem, (cudaStream_t)stream>>>(packed, read_off, read_len, order, n_reads,
                                                                               k, table, cap, counters);
  CFK_CHECK_LAUNCH("docfreq_kernel", 1);
  return CFK_OK;
}

int cfk_docfreq_plan(const int64_t* read_len, const int32_t* order, int64_t n_reads, int k, int32_t* n_pass,
                     cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_docfreq_plan: k must be in [1, 31]");
  if (n_reads < 0) return fail(CFK_ERR_INVALID, "cfk_docfreq_plan: bad sizes");
  if (n_reads == 0) return CFK_OK;
  docfreq_plan_kernel<<<(unsigned)blocks_for(n_reads, 256), 256, 0, (cudaStream_t)stream>>>(read_len, order, n_reads, k, n_pass);
  CFK_CHECK_LAUNCH("docfreq_plan_kernel", 1);
  return CFK_OK;
}

int cfk_docfreq_count_resident(const uint32_t* packed, const int64_t* read_off, const int64_t* read_len,
                               const int32_t* order, const int64_t* item_ptr, int64_t n_reads, int k, uint64_t* table,
                               int64_t cap, int64_t* counters, int32_t n_blocks, cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_docfreq_count_resident: k must be in [1, 31]");
  if (cap < 1 || n_reads < 0 || n_blocks < 1) return fail(CFK_ERR_INVALID, "cfk_docfreq_count_resident: bad sizes");
  if (n_reads == 0) return CFK_OK;
  static unsigned long long attr_done = 0;  // per device (bit = device ordinal)
  const int smem = DF3_SMEM_WORDS * 4;
  {
    cudaError_t e = cfk::ensure_dynamic_smem(docfreq_resident_kernel, smem, &attr_done);
    if (e != cudaSuccess) return fail(CFK_ERR_CUDA, "cfk_docfreq_count_resident: cudaFuncSetAttribute", e);
  }
  docfreq_resident_kernel<<<(unsigned)n_blocks * DF3_BLOCKS_PER_SM, DF3_THREADS, smem, (cudaStream_t)stream>>>(
      packed, read_off, read_len, order, item_ptr, n_reads, k, table, cap, counters);
  CFK_CHECK_LAUNCH("docfreq_resident_kernel", 1);
  return CFK_OK;
}

int cfk_kmer_count_total(const uint32_t* packed, const int64_t* read_off, const int64_t* read_len, const int32_t* tile_read,
                         const int64_t* tile_start, int64_t n_tiles, int k, uint64_t* table, int64_t cap, int64_t* counters,
                         cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_kmer_count_total: k must be in [1, 31]");
  if (cap < 1 || n_tiles < 0 || n_tiles >= (1ll << 31)) return fail(CFK_ERR_INVALID, "cfk_kmer_count_total: bad sizes");
  if (n_tiles == 0) return CFK_OK;
  kmer_count_kernel<false><<<(unsigned)n_tiles, KC_THREADS, 0, (cudaStream_t)stream>>>(packed, read_off, read_len, tile_read,
                                                                                     tile_start, k, table, cap, counters);
  CFK_CHECK_LAUNCH("kmer_count_kernel", 1);
  return CFK_OK;
}

int cfk_kmer_count_canonical(const uint32_t* packed, const int64_t* read_off, const int64_t* read_len, const int32_t* tile_read,
                             const int64_t* tile_start, int64_t n_tiles, int k, uint64_t* table, int64_t cap,
                             int64_t* counters, cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_kmer_count_canonical: k must be in [1, 31]");
  if (cap < 1 || n_tiles < 0 || n_tiles >= (1ll << 31)) return fail(CFK_ERR_INVALID, "cfk_kmer_count_canonical: bad sizes");
  if (n_tiles == 0) return CFK_OK;
  kmer_count_kernel<true><<<(unsigned)n_tiles, KC_THREADS, 0, (cudaStream_t)stream>>>(packed, read_off, read_len, tile_read,
                                                                                    tile_start, k, table, cap, counters);
  CFK_CHECK_LAUNCH("kmer_count_kernel", 1);
  return CFK_OK;
}

int cfk_kmer_count_tile(void) { return KC_TILE; }

int cfk_kmer_position_keys(const uint32_t* packed, int64_t n_bases, int k, int pos_bits, uint64_t* keys, cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_kmer_position_keys: k must be in [1, 31]");
  if (pos_bits < 1 || 2 * k + pos_bits > 63 || n_bases < 0 || n_bases >= (1ll << pos_bits) || n_bases >= (1ll << 32))
    return fail(CFK_ERR_INVALID, "cfk_kmer_position_keys: need n_bases < 2^pos_bits and 2k + pos_bits <= 63");
  const int64_t n = n_bases - k + 1;
  if (n <= 0) return CFK_OK;
  kmer_position_keys_kernel<<<(unsigned)blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(packed, n, k, pos_bits, keys);
  CFK_CHECK_LAUNCH("kmer_position_keys_kernel", 1);
  return CFK_OK;
}

int cfk_adjacent_gaps(const uint64_t* keys, int64_t n, int pos_bits, uint32_t* gaps, cfk_stream_t stream) {
  if (n < 0 || pos_bits < 1 || pos_bits > 62) return fail(CFK_ERR_INVALID, "cfk_adjacent_gaps: bad sizes");
  if (n == 0) return CFK_OK;
  adjacent_gaps_kernel<<<(unsigned)blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(keys, n, pos_bits, gaps);
  CFK_CHECK_LAUNCH("adjacent_gaps_kernel", 1);
  return CFK_OK;
}

int cfk_table_merge(const uint64_t* keys, const uint32_t* nreads, const uint32_t* nmulti, int64_t n, uint64_t* table,
                    int64_t cap, int64_t* counters, cfk_stream_t stream) {
  if (n < 0 || cap < 1) return fail(CFK_ERR_INVALID, "cfk_table_merge: bad sizes");
  if (n == 0) return CFK_OK;
  table_merge_kernel<<<(unsigned)blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(keys, nreads, nmulti, n, table, cap,
                                                                                      counters);
  CFK_CHECK_LAUNCH("table_merge_kernel", 1);
  return CFK_OK;
}

int cfk_table_lookup(const uint64_t* table, int64_t cap, const uint64_t* keys, int64_t n, uint32_t* out_nreads,
                     uint32_t* out_nmulti, cfk_stream_t stream) {
  if (n < 0 || cap < 1) return fail(CFK_ERR_INVALID, "cfk_table_lookup: bad sizes");
  if (n == 0) return CFK_OK;
  table_lookup_kernel<<<(unsigned)blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(table, cap, keys, n, out_nreads, out_nmulti);
  CFK_CHECK_LAUNCH("table_lookup_kernel", 1);
  return CFK_OK;
}

int cfk_table_select(const uint64_t* table, int64_t cap, uint32_t lo, uint32_t hi, uint32_t max_nonuniq, int32_t n_parts,
                     int32_t part, uint64_t* out_keys, uint32_t* out_nreads, uint32_t* out_nmulti, int64_t max_out,
                     int64_t* counters, cfk_stream_t stream) {
  if (cap < 1 || max_out < 0) return fail(CFK_ERR_INVALID, "cfk_table_select: bad sizes");
  if (n_parts > 0 && (part < 0 || part >= n_parts)) return fail(CFK_ERR_INVALID, "cfk_table_select: bad partition");
  table_select_kernel<<<(unsigned)blocks_for(cap, SEL_THREADS * SEL_ITEMS), SEL_THREADS, 0, (cudaStream_t)stream>>>(
      table, cap, lo, hi, max_nonuniq, n_parts, part, out_keys, out_nreads, out_nmulti, max_out, counters);
  CFK_CHECK_LAUNCH("table_select_kernel", 1);
  return CFK_OK;
}

int cfk_table_part_count(const uint64_t* table, int64_t cap, int32_t n_parts, int64_t* counts, cfk_stream_t stream) {
  if (cap < 1 || n_parts < 1 || n_parts > TP_MAX_PARTS)
    return fail(CFK_ERR_INVALID, "cfk_table_part_count: need cap >= 1 and 1 <= n_parts <= 64");
  const int64_t nb = blocks_for(cap, 256 * 8);
  table_part_count_kernel<<<(unsigned)(nb < 1 ? 1 : nb), 256, 0, (cudaStream_t)stream>>>(table, cap, n_parts, counts);
  CFK_CHECK_LAUNCH("table_part_count_kernel", 1);
  return CFK_OK;
}

int cfk_table_part_scatter(const uint64_t* table, int64_t cap, int32_t n_parts, int64_t* cursors, uint64_t* out_keys,
                           uint32_t* out_nreads, uint32_t* out_nmulti, cfk_stream_t stream) {
  if (cap < 1 || n_parts < 1 || n_parts > TP_MAX_PARTS)
    return fail(CFK_ERR_INVALID, "cfk_table_part_scatter: need cap >= 1 and 1 <= n_parts <= 64");
  const int64_t nb = blocks_for(cap, 256 * 8);
  table_part_scatter_kernel<<<(unsigned)(nb < 1 ? 1 : nb), 256, 0, (cudaStream_t)stream>>>(table, cap, n_parts, cursors,
                                                                                           out_keys, out_nreads, out_nmulti);
  CFK_CHECK_LAUNCH("table_part_scatter_kernel", 1);
  return CFK_OK;
}

int cfk_sort_u64(uint64_t* keys, int64_t n, cfk_stream_t stream) {
  if (n < 0) return fail(CFK_ERR_INVALID, "cfk_sort_u64: n < 0");
  if (n <= 1) return CFK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t np2 = SORT_TILE;
  while (np2 < n) np2 <<= 1;
  const unsigned tiles = (unsigned)blocks_for(n, SORT_TILE);
  sort_tile_kernel<<<tiles, SORT_THREADS, 0, st>>>(keys, n, 1);
  CFK_CHECK_LAUNCH("sort_tile_kernel", 1);
  const int64_t n_pairs = np2 / 2;
  int launched = 0;
  for (int64_t h = SORT_TILE; h < np2; h <<= 1) {
    sort_flip_kernel<<<(unsigned)blocks_for(n_pairs, 256), 256, 0, st>>>(keys, n, h, n_pairs);
    ++launched;
    for (int64_t hh = h >> 1; hh >= SORT_TILE; hh >>= 1, ++launched)
      sort_disperse_kernel<<<(unsigned)blocks_for(n_pairs, 256), 256, 0, st>>>(keys, n, hh, n_pairs);
    sort_tile_kernel<<<tiles, SORT_THREADS, 0, st>>>(keys, n, 0);
    ++launched;
  }
  CFK_CHECK_LAUNCH("sort_u64", launched);
  return CFK_OK;
}

int cfk_merge_sorted_runs(const uint64_t* keys, const int64_t* run_ptr, int32_t n_runs, int64_t n, uint64_t* out,
                          cfk_stream_t stream) {
  if (n_runs < 1 || n < 0) return fail(CFK_ERR_INVALID, "cfk_merge_sorted_runs: bad sizes");
  if (n == 0) return CFK_OK;
  merge_runs_kernel<<<(unsigned)blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(keys, run_ptr, n_runs, out);
  CFK_CHECK_LAUNCH("merge_runs_kernel", 1);
  return CFK_OK;
}

int cfk_index_filter_build(const uint64_t* sorted_keys, int64_t n, int32_t filter_bits, uint32_t* filter, cfk_stream_t stream) {
  if (n < 0 || filter_bits < 5 || filter_bits > 36) return fail(CFK_ERR_INVALID, "cfk_index_filter_build: bad sizes");
  if (n == 0) return CFK_OK;
  index_filter_kernel<<<(unsigned)blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(sorted_keys, n, filter_bits, filter);
  CFK_CHECK_LAUNCH("index_filter_kernel", 1);
  return CFK_OK;
}

int cfk_index_build(const uint64_t* sorted_keys, int64_t n, uint64_t* idx_keys, uint32_t* idx_vals, int64_t cap,
                    int64_t* counters, cfk_stream_t stream) {
  if (n < 0 || cap < 1 || n >= (1ll << 32) - 1) return fail(CFK_ERR_INVALID, "cfk_index_build: bad sizes");
  if (n == 0) return CFK_OK;
  index_build_kernel<<<(unsigned)blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(sorted_keys, n, idx_keys, idx_vals,
                                                                                      cap, counters);
  CFK_CHECK_LAUNCH("index_build_kernel", 1);
  return CFK_OK;
}

int cfk_cloud_build(const uint32_t* packed, const int64_t* unit_off, const int32_t* unit_len,
                    const int64_t* unit_kbase, int64_t n_units, int k, const uint64_t* idx_keys,
                    const uint32_t* idx_vals, int64_t cap, const uint32_t* filter, int32_t filter_bits, uint32_t* tmp_ids,
                    int32_t* unit_cnt, cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_cloud_build: k must be in [1, 31]");
  if (n_units < 0 || n_units >= (1ll << 31) || cap < 1) return fail(CFK_ERR_INVALID, "cfk_cloud_build: bad sizes");
  if (filter != nullptr && (filter_bits < 5 || filter_bits > 36)) return fail(CFK_ERR_INVALID, "cfk_cloud_build: filter_bits");
  if (n_units == 0) return CFK_OK;
  static const bool block_form = [] { const char* e = getenv("CFK_CLOUD_MODE"); return e && !strcmp(e, "block"); }();
  if (block_form) {  // the first version of the kernel, one block per unit (A/B and cross-checks)
    cloud_build_kernel<<<(unsigned)n_units, CL_THREADS, 0, (cudaStream_t)stream>>>(packed, unit_off, unit_len, unit_kbase, k,
                                                                                   idx_keys, idx_vals, cap, tmp_ids, unit_cnt);
    CFK_CHECK_LAUNCH("cloud_build_kernel", 1);
    return CFK_OK;
  }
  cloud_build_warp_kernel<<<(unsigned)blocks_for(n_units, CLW_WARPS), CLW_WARPS * 32, 0, (cudaStream_t)stream>>>(
      packed, unit_off, unit_len, unit_kbase, n_units, k, idx_keys, idx_vals, cap, filter, filter_bits, tmp_ids, unit_cnt);
  CFK_CHECK_LAUNCH("cloud_build_warp_kernel", 1);
  return CFK_OK;
}

int64_t cfk_scan_scratch_elems(int64_t n) { return blocks_for(n > 0 ? n : 1, SCAN_TILE) + 1; }

int cfk_exclusive_scan(const int32_t* in, int64_t* out, int64_t n, int64_t* scratch, cfk_stream_t stream) {
  if (n < 0) return fail(CFK_ERR_INVALID, "cfk_exclusive_scan: n < 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(int64_t), st);
    if (e != cudaSuccess) return fail(CFK_ERR_CUDA, "cfk_exclusive_scan memset", e);
    return CFK_OK;
  }
  const int64_t nb = blocks_for(n, SCAN_TILE);
  scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, scratch);
  scan_partials_kernel<<<1, 1024, 0, st>>>(scratch, nb);
  scan_apply_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, scratch, out);
  CFK_CHECK_LAUNCH("exclusive_scan", 3);
  return CFK_OK;
}

int cfk_cloud_compact(const uint32_t* tmp_ids, const int64_t* unit_kbase, const int64_t* unit_ptr, int64_t n_units,
                      uint32_t* ids, cfk_stream_t stream) {
  if (n_units < 0) return fail(CFK_ERR_INVALID, "cfk_cloud_compact: n_units < 0");
  if (n_units == 0) return CFK_OK;
  cloud_compact_kernel<<<(unsigned)blocks_for(n_units * 32, 256), 256, 0, (cudaStream_t)stream>>>(tmp_ids, unit_kbase,
                                                                                                  unit_ptr, n_units, ids);
  CFK_CHECK_LAUNCH("cloud_compact_kernel", 1);
  return CFK_OK;
}

int cfk_id_histogram(const int64_t* unit_ptr, const uint32_t* ids, int64_t unit_lo, int64_t unit_hi, int32_t* mult,
                     cfk_stream_t stream) {
  if (unit_lo < 0 || unit_hi < unit_lo) return fail(CFK_ERR_INVALID, "cfk_id_histogram: bad unit range");
  if (unit_hi == unit_lo) return CFK_OK;
  id_histogram_kernel<<<(unsigned)blocks_for((unit_hi - unit_lo) * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      unit_ptr, ids, unit_lo, unit_hi, mult);
  CFK_CHECK_LAUNCH("id_histogram_kernel", 1);
  return CFK_OK;
}

int cfk_cloud_filter_count(const int64_t* unit_ptr, const uint32_t* ids, int64_t n_units, const int32_t* mult,
                           int64_t min_mult, int64_t max_mult, int32_t* new_cnt, cfk_stream_t stream) {
  if (n_units < 0) return fail(CFK_ERR_INVALID, "cfk_cloud_filter_count: n_units < 0");
  if (n_units == 0) return CFK_OK;
  cloud_filter_count_kernel<<<(unsigned)blocks_for(n_units * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      unit_ptr, ids, n_units, mult, min_mult, max_mult, new_cnt);
  CFK_CHECK_LAUNCH("cloud_filter_count_kernel", 1);
  return CFK_OK;
}

int cfk_cloud_filter_write(const int64_t* unit_ptr, const uint32_t* ids, int64_t n_units, const int32_t* mult,
                           int64_t min_mult, int64_t max_mult, const int64_t* new_ptr, uint32_t* new_ids,
                           cfk_stream_t stream) {
  if (n_units < 0) return fail(CFK_ERR_INVALID, "cfk_cloud_filter_write: n_units < 0");
  if (n_units == 0) return CFK_OK;
  cloud_filter_write_kernel<<<(unsigned)blocks_for(n_units * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      unit_ptr, ids, n_units, mult, min_mult, max_mult, new_ptr, new_ids);
  CFK_CHECK_LAUNCH("cloud_filter_write_kernel", 1);
  return CFK_OK;
}

int cfk_occ_fill(const int64_t* unit_ptr, const uint32_t* ids, int64_t unit_lo, int64_t unit_hi,
                 const int64_t* occ_ptr, int64_t n_kmers, uint32_t* cursor, uint32_t* occ, cfk_stream_t stream) {
  if (unit_lo < 0 || unit_hi < unit_lo || n_kmers < 0) return fail(CFK_ERR_INVALID, "cfk_occ_fill: bad unit range");
  if (unit_hi == unit_lo || n_kmers == 0) return CFK_OK;
  occ_cursor_kernel<<<(unsigned)blocks_for(n_kmers, 256), 256, 0, (cudaStream_t)stream>>>(occ_ptr, n_kmers, cursor);
  CFK_CHECK_LAUNCH("occ_cursor_kernel", 1);
  occ_fill_kernel<<<(unsigned)blocks_for((unit_hi - unit_lo) * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      unit_ptr, ids, unit_lo, unit_hi, cursor, occ);
  CFK_CHECK_LAUNCH("occ_fill_kernel", 1);
  return CFK_OK;
}

int cfk_occ_slice_histogram(const int64_t* unit_ptr, const uint32_t* ids, int64_t n_units, int64_t id_lo, int64_t id_hi,
                            int32_t* mult, cfk_stream_t stream) {
  if (n_units < 0 || id_lo < 0 || id_hi < id_lo || id_hi > (1ll << 32)) return fail(CFK_ERR_INVALID, "cfk_occ_slice_histogram: bad sizes");
  if (n_units == 0 || id_hi == id_lo) return CFK_OK;
  occ_slice_kernel<false><<<(unsigned)blocks_for(n_units * 16, 256), 256, 0, (cudaStream_t)stream>>>(
      unit_ptr, ids, n_units, id_lo, id_hi, mult, nullptr, nullptr);
  CFK_CHECK_LAUNCH("occ_slice_kernel<histogram>", 1);
  return CFK_OK;
}

int cfk_occ_slice_fill(const int64_t* unit_ptr, const uint32_t* ids, int64_t n_units, int64_t id_lo, int64_t id_hi,
                       const int64_t* occ_ptr, uint32_t* cursor, uint32_t* occ, cfk_stream_t stream) {
  if (n_units < 0 || id_lo < 0 || id_hi < id_lo || id_hi > (1ll << 32)) return fail(CFK_ERR_INVALID, "cfk_occ_slice_fill: bad sizes");
  if (n_units == 0 || id_hi == id_lo) return CFK_OK;
  occ_cursor_kernel<<<(unsigned)blocks_for(id_hi - id_lo, 256), 256, 0, (cudaStream_t)stream>>>(occ_ptr, id_hi - id_lo, cursor);
  CFK_CHECK_LAUNCH("occ_cursor_kernel", 1);
  occ_slice_kernel<true><<<(unsigned)blocks_for(n_units * 16, 256), 256, 0, (cudaStream_t)stream>>>(
      unit_ptr, ids, n_units, id_lo, id_hi, nullptr, cursor, occ);
  CFK_CHECK_LAUNCH("occ_slice_kernel<fill>", 1);
  return CFK_OK;
}

int cfk_occ_sort(const int64_t* occ_ptr, uint32_t* occ, int64_t n_kmers, cfk_stream_t stream) {
  if (n_kmers < 0) return fail(CFK_ERR_INVALID, "cfk_occ_sort: n_kmers < 0");
  if (n_kmers == 0) return CFK_OK;
  occ_sort_kernel<<<(unsigned)blocks_for(n_kmers, 128), 128, 0, (cudaStream_t)stream>>>(occ_ptr, occ, n_kmers);
  CFK_CHECK_LAUNCH("occ_sort_kernel", 1);
  return CFK_OK;
}

int cfk_occ_last(const uint32_t* occ, int64_t n, const uint32_t* unit_last, uint32_t* occ_last, cfk_stream_t stream) {
  if (n < 0) return fail(CFK_ERR_INVALID, "cfk_occ_last: n < 0");
  if (n == 0) return CFK_OK;
  occ_last_kernel<<<(unsigned)blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(occ, n, unit_last, occ_last);
  CFK_CHECK_LAUNCH("occ_last_kernel", 1);
  return CFK_OK;
}

int cfk_unit_splits(const int64_t* unit_ptr, const uint32_t* ids, int64_t n_units, int64_t n_entries, int64_t n_kmers,
                    uint32_t* usplit, cfk_stream_t stream) {
  if (n_units < 0 || n_kmers < 0 || n_entries < 0 || n_entries >= (1ll << 32))
    return fail(CFK_ERR_INVALID, "cfk_unit_splits: need 0 <= n_entries < 2^32");
  if (n_units == 0) return CFK_OK;
  unit_split_kernel<<<(unsigned)blocks_for(n_units * 7, 256), 256, 0, (cudaStream_t)stream>>>(unit_ptr, ids, n_units,
                                                                                           n_kmers, usplit);
  CFK_CHECK_LAUNCH("unit_split_kernel", 1);
  return CFK_OK;
}

int cfk_pair_candidates(const int64_t* unit_ptr, const uint32_t* ids, const uint32_t* unit_last,
                        const int64_t* occ_ptr, const uint32_t* occ, const uint32_t* usplit, int64_t n_entries,
                        int64_t n_kmers, int64_t a_begin, int64_t a_end, int32_t a_stride, int32_t min_d,
                        int32_t max_d, uint32_t min_cov, uint32_t* cand, int64_t max_cand, int64_t* counters,
                        int32_t n_blocks, cfk_stream_t stream) {
  if (n_entries < 0 || n_entries >= (1ll << 32))
    return fail(CFK_ERR_INVALID, "cfk_pair_candidates: need 0 <= n_entries < 2^32 (32-bit positions in the id array)");
  if (n_kmers < 0 || n_kmers >= (1ll << 32) - 1) return fail(CFK_ERR_INVALID, "cfk_pair_candidates: bad n_kmers");
  if (min_d < 0) return fail(CFK_ERR_INVALID, "cfk_pair_candidates: min_d < 0 is not defined by the reference loop");
  if (a_begin < 0 || a_end > n_kmers || a_stride < 1) return fail(CFK_ERR_INVALID, "cfk_pair_candidates: bad id range");
  if (max_cand < 0 || n_blocks < 1) return fail(CFK_ERR_INVALID, "cfk_pair_candidates: bad sizes");
  if (a_begin >= a_end || max_d < (min_d > 1 ? min_d : 1)) return CFK_OK;
  static unsigned long long attr_done = 0;  // per device (bit = device ordinal)
  const int smem = PC_WARPS * PC_TBL_BYTES;
  {
    cudaError_t e = cfk::ensure_dynamic_smem(pair_candidates_kernel, smem, &attr_done);
    if (e != cudaSuccess) return fail(CFK_ERR_CUDA, "cfk_pair_candidates: cudaFuncSetAttribute", e);
  }
  pair_candidates_kernel<<<(unsigned)n_blocks, PC_WARPS * 32, smem, (cudaStream_t)stream>>>(
      unit_ptr, ids, unit_last, occ_ptr, occ, usplit, n_kmers, a_begin, a_end, a_stride, min_d, max_d, min_cov,
      (uint4*)cand, max_cand, counters);
  CFK_CHECK_LAUNCH("pair_candidates_kernel", 1);
  return CFK_OK;
}

int cfk_sketch_bits(void) { return SK_BITS; }
int cfk_sketch_warps_per_block(void) { return SK_WARPS; }

int64_t cfk_sketch_codes_elems(int64_t n_entries, int64_t n_units) {
  if (n_entries < 0 || n_units < 0) return 0;
  return (sk_block(n_entries, n_units) + 1) * SK_WINDOW;
}

int cfk_sketch_codes(const int64_t* unit_ptr, const uint32_t* ids, int64_t n_units, uint16_t* codes, uint32_t* perm_ids,
                     cfk_stream_t stream) {
  if (n_units < 0) return fail(CFK_ERR_INVALID, "cfk_sketch_codes: n_units < 0");
  if (n_units == 0) return CFK_OK;
  if (((uintptr_t)codes & 7u) != 0) return fail(CFK_ERR_INVALID, "cfk_sketch_codes: codes must be 8-byte aligned");
  if (perm_ids != nullptr) {
    if (((uintptr_t)perm_ids & 15u) != 0) return fail(CFK_ERR_INVALID, "cfk_sketch_codes: perm_ids must be 16-byte aligned");
    sketch_codes_banked_kernel<<<(unsigned)blocks_for(n_units, 8), 256, 0, (cudaStream_t)stream>>>(unit_ptr, ids, n_units,
                                                                                                   (uint2*)codes, (uint4*)perm_ids);
    CFK_CHECK_LAUNCH("sketch_codes_banked_kernel", 1);
    return CFK_OK;
  }
  sketch_codes_kernel<<<(unsigned)blocks_for(n_units, 8), 256, 0, (cudaStream_t)stream>>>(unit_ptr, ids, n_units, (uint2*)codes);
  CFK_CHECK_LAUNCH("sketch_codes_kernel", 1);
  return CFK_OK;
}

int cfk_pair_sketch(const int64_t* unit_ptr, const uint32_t* ids, const uint16_t* codes, const uint32_t* perm_ids,
                    const uint32_t* unit_last,
                    const int64_t* occ_ptr, const uint32_t* occ, const uint32_t* occ_last, int64_t n_entries, int64_t n_kmers, int64_t a_begin,
                    int64_t a_end, int32_t a_stride, int32_t min_d, int32_t max_d, uint32_t min_cov, uint32_t* cand,
                    int64_t max_cand, int64_t* counters, int32_t n_blocks, cfk_stream_t stream) {
  if (n_entries < 0 || n_entries >= (1ll << 32))
    return fail(CFK_ERR_INVALID, "cfk_pair_sketch: need 0 <= n_entries < 2^32 (32-bit positions in the id array)");
  if (((uintptr_t)codes & 7u) != 0) return fail(CFK_ERR_INVALID, "cfk_pair_sketch: codes must be 8-byte aligned");
  if (n_kmers < 0 || n_kmers >= (1ll << 32) - 1) return fail(CFK_ERR_INVALID, "cfk_pair_sketch: bad n_kmers");
  if (min_d < 0) return fail(CFK_ERR_INVALID, "cfk_pair_sketch: min_d < 0 is not defined by the reference loop");
  if (min_cov < CFK_SKETCH_MIN_COV || min_cov > CFK_SKETCH_MAX_COV)
    return fail(CFK_ERR_INVALID, "cfk_pair_sketch: min_cov outside [CFK_SKETCH_MIN_COV, CFK_SKETCH_MAX_COV]; use cfk_pair_candidates");
  if (a_begin < 0 || a_end > n_kmers || a_stride < 1) return fail(CFK_ERR_INVALID, "cfk_pair_sketch: bad id range");
  if (max_cand < 0 || n_blocks < 1) return fail(CFK_ERR_INVALID, "cfk_pair_sketch: bad sizes");
  if (a_begin >= a_end || max_d < (min_d > 1 ? min_d : 1)) return CFK_OK;
  static unsigned long long attr_done = 0;  // per device (bit = device ordinal)
  const int smem = SK_WARPS * SK_WARP_BYTES;
  {
    cudaError_t e = cfk::ensure_dynamic_smem(pair_sketch_kernel, smem, &attr_done);
    if (e != cudaSuccess) return fail(CFK_ERR_CUDA, "cfk_pair_sketch: cudaFuncSetAttribute", e);
  }
  // codes + ids of the graph beyond ~half of the 126 MB L2 (the all-gathered graph of several GPUs): the kernel asks
  // each pass's code blocks into L2 ahead of its walk (CFK_SKETCH_PREFETCH=0 / 1 forces it off / on)
  int prefetch = 6 * n_entries > (48ll << 20);
  if (const char* env = getenv("CFK_SKETCH_PREFETCH")) prefetch = atoi(env) != 0;
  pair_sketch_kernel<<<(unsigned)n_blocks, SK_WARPS * 32, smem, (cudaStream_t)stream>>>(
      unit_ptr, ids, (const uint2*)codes, perm_ids, unit_last, occ_ptr, occ, occ_last, n_kmers, a_begin, a_end, a_stride, min_d, max_d, min_cov, (uint4*)cand,
      max_cand, counters, prefetch);
  CFK_CHECK_LAUNCH("pair_sketch_kernel", 1);
  return CFK_OK;
}

int cfk_pair_join(const uint32_t* cand, int64_t n_cand, const int64_t* occ_ptr, const uint32_t* occ, const uint32_t* occ_last,
                  const uint32_t* unit_last, int32_t min_d, int32_t max_d, uint32_t min_cov, double rel_threshold,
                  uint32_t* edges, int64_t max_edges, uint8_t* selected, int64_t* counters, cfk_stream_t stream) {
  if (n_cand < 0 || max_edges < 0) return fail(CFK_ERR_INVALID, "cfk_pair_join: bad sizes");
  if (n_cand == 0) return CFK_OK;
  pair_join_kernel<<<(unsigned)blocks_for(n_cand, 128), 128, 0, (cudaStream_t)stream>>>(
      (const uint4*)cand, n_cand, occ_ptr, occ, occ_last, unit_last, min_d, max_d, min_cov, rel_threshold, (uint4*)edges,
      max_edges, selected, counters);
  CFK_CHECK_LAUNCH("pair_join_kernel", 1);
  return CFK_OK;
}

int cfk_flag_indices(const uint8_t* flags, int64_t n, uint32_t* out, int64_t* counters, cfk_stream_t stream) {
  if (n < 0) return fail(CFK_ERR_INVALID, "cfk_flag_indices: n < 0");
  if (n == 0) return CFK_OK;
  flag_indices_kernel<<<(unsigned)blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(flags, n, out, counters);
  CFK_CHECK_LAUNCH("flag_indices_kernel", 1);
  return CFK_OK;
}

}  // extern "C"
