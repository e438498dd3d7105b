// docfreq_stream.cu — stage A (document frequency), two-phase form: the default path of cfk_docfreq.
//
// Replaces get_kmer_freqs_from_ncrf_report, scripts/distance_based_kmer_recruitment.py:39-63.
//
//   phase 1  docfreq_emit_kernel   per (read, pass) item: the packed read is brought into shared memory by ONE
//            bulk copy (cp.async.bulk + mbarrier -- the TMA engine, no register round trip), the read's k-mers are
//            de-duplicated in a shared-memory set WITHOUT atomics (claim with plain stores, verify after a
//            barrier), and one 8-byte record per DISTINCT k-mer of the read -- bit 63 = "occurs more than once in
//            this read" -- is appended to one of DE_PARTS hash partitions in HBM (block-wide counting scatter:
//            records of one partition leave the SM as one contiguous run).
//   phase 2  docfreq_apply_kernel  walks the partitions in order.  partition = top bits of mix64(k-mer) and the home
//            slot of the global table is monotone in the same hash, so all updates of one partition fall into one
//            1/DE_PARTS window of the table: the CAS claim and the counter add hit L2, not DRAM.
//
// The table, its slot format and everything downstream (band filter, multi-GPU exchange) are those of cfk.cu.
#include "cfk_common.cuh"

namespace {

using namespace cfk;

#ifndef CFK_DE_THREADS
#define CFK_DE_THREADS 1024
#endif
#ifndef CFK_DE_FILL_PCT
#define CFK_DE_FILL_PCT 72   /* planned load of the per-read set, percent (4-slot buckets) */
#endif
constexpr int DE_THREADS = CFK_DE_THREADS;
constexpr int DE_WARPS = DE_THREADS / 32;
constexpr int DE_PER = 4;                   // consecutive k-mer starts per lane
constexpr int DE_CHUNK = 32 * DE_PER;       // k-mer starts per warp step
constexpr int DE_PART_BITS = CFK_DOCFREQ_PART_BITS;
constexpr int DE_PARTS = 1 << DE_PART_BITS;
constexpr int DE_SMEM_WORDS = 57344;        // dynamic shared memory of the block (224 KB)
constexpr int DE_HIST_WORDS = 2 * DE_WARPS * DE_PARTS / 2;  // two u16 [warp][partition] matrices
constexpr int DE_BASE_WORDS = 2 * DE_PARTS;                 // int64 first output index per partition
constexpr int DE_DATA_WORDS = DE_SMEM_WORDS - DE_HIST_WORDS - DE_BASE_WORDS;  // [ read words | set ]
constexpr int DE_MIN_SET = 16384;
constexpr uint32_t DE_MULTI = 0x80000000u;
static_assert(DE_PARTS <= DE_THREADS && DE_PARTS <= 256, "one thread per partition in the scatter; keys of match.any");
static_assert(DE_DATA_WORDS > 2 * DE_MIN_SET, "shared memory budget");

struct DeGeometry {
  uint32_t n_words;    // words staged (0: the read stays in global memory)
  uint32_t n_buckets;  // 4-slot buckets of the set
  uint32_t fill;       // k-mers planned per pass
};

__host__ __device__ __forceinline__ DeGeometry de_geometry(int64_t len) {
  DeGeometry g;
  const int64_t nw = (((len + 15) >> 4) + 3 + 3) & ~(int64_t)3;  // + the 3-word extraction window, multiple of 4
  g.n_words = (nw <= DE_DATA_WORDS - DE_MIN_SET) ? (uint32_t)nw : 0u;
  g.n_buckets = ((uint32_t)DE_DATA_WORDS - g.n_words) >> 2;
  g.fill = (uint32_t)((uint64_t)g.n_buckets * 4u * CFK_DE_FILL_PCT / 100);
  return g;
}

__global__ void docfreq_emit_plan_kernel(const int64_t* __restrict__ read_len, const int32_t* __restrict__ order,
                                         int64_t n_reads, int k, int32_t* __restrict__ n_pass) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_reads) return;
  const int64_t len = read_len[order[i]], nk = len - k + 1;
  int32_t np = 0;
  if (nk > 0) {
    const DeGeometry g = de_geometry(len);
    np = (int32_t)((nk + g.fill - 1) / g.fill);
  }
  n_pass[i] = np;
}

template <bool IN_SMEM>
__device__ __forceinline__ uint32_t de_word(const uint32_t* words, uint32_t i) {
  if (IN_SMEM) return words[i];
  return __ldg(words + i);
}

// k-mer starting at base q; words[] must be readable up to word (q >> 4) + 2
template <bool IN_SMEM>
__device__ __forceinline__ uint64_t de_kmer_at(const uint32_t* words, uint32_t q, int k) {
  const uint32_t w = q >> 4, sh = (q & 15u) << 1;
  uint64_t bits = ((uint64_t)de_word<IN_SMEM>(words, w) | ((uint64_t)de_word<IN_SMEM>(words, w + 1) << 32)) >> sh;
  if (sh) bits |= (uint64_t)de_word<IN_SMEM>(words, w + 2) << (64 - sh);
  uint64_t r = __brevll(bits);  // base q + j sits at bits 2j..2j+1; the k-mer wants base q on top
  r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);
  return r >> (64 - 2 * k);
}

__device__ __forceinline__ uint32_t de_fold(uint64_t kmer) { return (uint32_t)(kmer ^ (kmer >> 32)); }
__device__ __forceinline__ uint32_t de_mix(uint32_t f) {
  uint32_t g = f * 0x85EBCA6Bu;
  g ^= g >> 13;
  g *= 0xC2B2AE35u;
  g ^= g >> 16;
  return g;
}

// ---- the per-read set ---------------------------------------------------------------------------------------
// Slot (32 bit): bit 31 = the k-mer occurs again in this read; bits [pos_bits-1:0] = position of the occurrence that
// owns the slot + 1 (0 = empty, all ones = dead); the bits between = fingerprint of the k-mer.  Buckets of 4 slots
// (one 16-byte load), linear probing over buckets.
//
// CLAIM phase (all positions): walk the chain; at the first empty slot store the own tag with a PLAIN store and go
// on.  Racing claims of one slot overwrite each other: the loser's k-mer is simply not in the set yet.
// VERIFY phase (after a barrier: every claim has landed): walk the chain again.  The first slot holding the k-mer is
// its canonical entry.  Own tag -> this occurrence is the owner.  Somebody else's -> set the "again" bit (an
// idempotent plain store) and keep walking: an own tag further down is a stale second entry (its claimer had walked
// past a slot that a racing claim filled with this k-mer afterwards) and is marked dead.  An empty slot before any
// match: the claim was lost -> claim again with atomicCAS (rare; all writes to EMPTY slots in this phase are CAS).
// Slots never return to empty, so a chain has no holes and "first match from home" is the same slot for every walker.
template <bool IN_SMEM, bool VERIFY>
__device__ __forceinline__ void de_phase(const uint32_t* words, uint32_t* set, uint32_t nb, int64_t nk, int k, uint32_t pass,
                                         uint32_t n_pass, int64_t* counters) {
  const uint64_t mask = (1ull << (2 * k)) - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pos_bits = 64 - __clzll((unsigned long long)(nk + 1));  // pos + 1 <= nk < 2^pos_bits - 1
  const uint32_t pos_mask = (1u << pos_bits) - 1u;                  // nk < 2^30 (checked on the host)
  const uint32_t fp_mask = 0x7FFFFFFFu & ~pos_mask;
  for (int64_t chunk0 = (int64_t)warp * DE_CHUNK; chunk0 < nk; chunk0 += (int64_t)DE_WARPS * DE_CHUNK) {
    const int64_t base = chunk0 + lane * DE_PER;
    if (base >= nk) continue;
    const uint32_t p0 = (uint32_t)base;  // multiple of 4: offset 0, 4, 8 or 12 inside its word
    const int npos = (int)min((int64_t)DE_PER, nk - base);
    // 48-base window from the word of p0: offset + DE_PER - 1 + k - 1 <= 12 + 3 + 30 < 48
    const uint32_t w0 = p0 >> 4;
    uint64_t win_lo = (uint64_t)de_word<IN_SMEM>(words, w0) | ((uint64_t)de_word<IN_SMEM>(words, w0 + 1) << 32);
    uint32_t win_hi = de_word<IN_SMEM>(words, w0 + 2);
    if (const uint32_t sh0 = (p0 & 15u) << 1) {
      win_lo = (win_lo >> sh0) | ((uint64_t)win_hi << (64 - sh0));
      win_hi >>= sh0;
    }
    uint64_t kmer = 0;
    const int s0 = 2 * (k - 1);
    if (s0) {  // the first k - 1 bases in one go, then roll
      uint64_t r = __brevll(win_lo);
      r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);
      kmer = r >> (64 - s0);
      win_lo = (win_lo >> s0) | ((uint64_t)win_hi << (64 - s0));
      win_hi = s0 < 32 ? (win_hi >> s0) : 0u;
    }
#pragma unroll
    for (int j = 0; j < DE_PER; ++j) {
      kmer = ((kmer << 2) | (win_lo & 3u)) & mask;
      win_lo = (win_lo >> 2) | ((uint64_t)win_hi << 62);
      win_hi >>= 2;
      if (j >= npos) continue;
      const uint32_t f = de_fold(kmer);
      if (n_pass > 1 && __umulhi(f * 0x9E3779B1u, n_pass) != pass) continue;
      const uint32_t g = de_mix(f);
      const uint32_t mine = (g << pos_bits) & fp_mask;
      const uint32_t fresh = (p0 + (uint32_t)j + 1u) | mine;
      uint32_t b = __umulhi(g, nb);
      bool matched = false;  // VERIFY: the canonical entry was somebody else's (now looking for a stale own entry)
      bool done = false;
      uint32_t probes = 0;
      for (; probes < nb && !done; ++probes) {
        const uint4 v4 = *reinterpret_cast<const uint4*>(set + 4 * b);
        const uint32_t vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (done) break;
          uint32_t v = vv[s];
          if (!VERIFY) {
            if (v == 0) {
              set[4 * b + s] = fresh;
              done = true;
            } else if (((v ^ mine) & fp_mask) == 0 && de_kmer_at<IN_SMEM>(words, (v & pos_mask) - 1u, k) == kmer) {
              done = true;  // already there (perhaps only for now: the verify phase decides)
            }
          } else {
            if (v == 0) {
              if (matched) {
                done = true;
                break;
              }
              v = atomicCAS(set + 4 * b + s, 0u, fresh);
              if (v == 0) {  // the lost claim, made good
                done = true;
                break;
              }
            }
            if ((v & ~DE_MULTI) == fresh) {
              if (matched) set[4 * b + s] = v | pos_mask;  // stale second entry of this k-mer: dead
              done = true;
            } else if (!matched && ((v ^ mine) & fp_mask) == 0 && (v & pos_mask) != pos_mask &&
                       de_kmer_at<IN_SMEM>(words, (v & pos_mask) - 1u, k) == kmer) {
              if (!(v & DE_MULTI)) set[4 * b + s] = v | DE_MULTI;
              matched = true;
            }
          }
        }
        if (++b == nb) b = 0;
      }
      if (!done) counters[1] = 1;  // cannot happen: a pass is planned for <= CFK_DE_FILL_PCT % load
    }
  }
}

// ticket -> (index into order[], pass, passes of that read); index -1 when the items are used up
__device__ __forceinline__ void de_fetch(const int64_t* __restrict__ item_ptr, int64_t n_reads, int64_t n_items,
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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Final scan of the set: every live slot becomes one record, appended to its hash partition.  One step = one bucket per
// thread; the records of a step are ranked per partition (match.any inside the warp, a u16 [warp][partition] matrix
// across warps) so that each partition receives ONE contiguous run per step and one atomicAdd on its cursor.
template <bool IN_SMEM>
__device__ __forceinline__ void de_scan_emit(const uint32_t* words, const uint32_t* set, uint32_t nb, int64_t nk, int k,
                                             uint16_t* hist, int64_t* s_base, uint64_t* __restrict__ records,
                                             int64_t part_cap, int64_t* cursors, int64_t* counters) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pos_bits = 64 - __clzll((unsigned long long)(nk + 1));
  const uint32_t pos_mask = (1u << pos_bits) - 1u;
  const unsigned lt = (1u << lane) - 1u;
  int which = 0;
  for (uint32_t b0 = 0; b0 < nb; b0 += DE_THREADS, which ^= 1) {
    uint16_t* hh = hist + which * (DE_WARPS * DE_PARTS);
    const uint32_t b = b0 + threadIdx.x;
    uint4 v4 = make_uint4(0, 0, 0, 0);
    if (b < nb) v4 = *reinterpret_cast<const uint4*>(set + 4 * b);
    const uint32_t vv[4] = {v4.x, v4.y, v4.z, v4.w};
    uint64_t rec[4];
    uint32_t where[4];  // partition << 16 | rank inside (warp, partition); 0xFFFFFFFF: no record
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const uint32_t v = vv[s];
      const bool live = v != 0 && (v & pos_mask) != pos_mask;
      uint32_t part = DE_PARTS + lane;
      rec[s] = 0;
      if (live) {
        const uint64_t kmer = de_kmer_at<IN_SMEM>(words, (v & pos_mask) - 1u, k);
        part = (uint32_t)(mix64(kmer) >> (64 - DE_PART_BITS));
        rec[s] = kmer | ((uint64_t)(v >> 31) << 63);
      }
      const unsigned peers = __match_any_sync(FULL, part);
      const int leader = __ffs(peers) - 1;
      uint32_t first = 0;
      if (live && lane == leader) {
        first = hh[warp * DE_PARTS + part];
        hh[warp * DE_PARTS + part] = (uint16_t)(first + __popc(peers));
      }
      first = __shfl_sync(FULL, first, leader);
      where[s] = live ? ((part << 16) | (first + __popc(peers & lt))) : 0xFFFFFFFFu;
      __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < DE_PARTS) {
      uint32_t run = 0;
#pragma unroll 4
      for (int w = 0; w < DE_WARPS; ++w) {
        const uint32_t c = hh[w * DE_PARTS + threadIdx.x];
        hh[w * DE_PARTS + threadIdx.x] = (uint16_t)run;
        run += c;
      }
      long long g = 0;
      if (run) {
        g = (long long)atomicAdd((unsigned long long*)(cursors + threadIdx.x), (unsigned long long)run);
        if (g + run > part_cap) counters[0] = 1;  // partition buffer full: the host falls back (never silent)
      }
      s_base[threadIdx.x] = g;
      uint16_t* other = hist + (which ^ 1) * (DE_WARPS * DE_PARTS);
#pragma unroll 4
      for (int w = 0; w < DE_WARPS; ++w) other[w * DE_PARTS + threadIdx.x] = 0;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (where[s] == 0xFFFFFFFFu) continue;
      const uint32_t part = where[s] >> 16;
      const int64_t dst = s_base[part] + hh[warp * DE_PARTS + part] + (where[s] & 0xFFFFu);
      if (dst < part_cap) records[(int64_t)part * part_cap + dst] = rec[s];
    }
  }
}

__global__ void __launch_bounds__(DE_THREADS, 1)
docfreq_emit_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ read_off,
                    const int64_t* __restrict__ read_len, const int32_t* __restrict__ order,
                    const int64_t* __restrict__ item_ptr, int64_t n_reads, int k, uint64_t* __restrict__ records,
                    int64_t part_cap, int64_t* cursors, int64_t* counters) {
  extern __shared__ __align__(16) uint32_t de_smem[];
  uint32_t* data = de_smem;                                                    // [ read words | set ]
  uint16_t* hist = reinterpret_cast<uint16_t*>(de_smem + DE_DATA_WORDS);       // 2 x [warp][partition]
  int64_t* s_base = reinterpret_cast<int64_t*>(de_smem + DE_DATA_WORDS + DE_HIST_WORDS);
  __shared__ __align__(8) uint64_t s_mbar;
  __shared__ long long s_read[2];
  __shared__ uint32_t s_pass[2], s_npass[2];
  const int64_t n_items = __ldg(item_ptr + n_reads);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    de_fetch(item_ptr, n_reads, n_items, counters, &s_read[0], &s_pass[0], &s_npass[0]);
  }
  uint32_t parity = 0;
  for (int cur = 0;; cur ^= 1) {
    __syncthreads();  // item `cur` is published; everybody is done with the previous item's shared memory
    const int64_t idx = s_read[cur];
    if (idx < 0) break;
    const uint32_t pass = s_pass[cur], n_pass = s_npass[cur];
    const int64_t r = order[idx];
    const int64_t len = read_len[r], nk = len - k + 1;
    const uint32_t* gwords = packed + (read_off[r] >> 4);  // every read starts on a 64-base boundary
    const DeGeometry g = de_geometry(len);
    uint32_t* set = data + g.n_words;
    // a single-pass read gets a set sized for its own k-mers (less to clear and to scan)
    const uint32_t nb = (n_pass > 1) ? g.n_buckets
                                     : (uint32_t)min((int64_t)g.n_buckets, max((int64_t)512, (nk * 100 / CFK_DE_FILL_PCT) / 4 + 2));
    const uint32_t real_quads = (uint32_t)(((len + 15) >> 4) + 3) >> 2;  // 16-byte pieces of the read's own 64-base blocks
    if (g.n_words && threadIdx.x == 0) {
      // the read travels global -> shared as one bulk copy of the TMA engine; the mbarrier counts its bytes
      const uint32_t bytes = real_quads * 16u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_mbar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(data)),
                   "l"(gwords), "r"(bytes), "r"(smem_u32(&s_mbar))
                   : "memory");
    }
    for (uint32_t i = threadIdx.x; i < nb; i += DE_THREADS) *reinterpret_cast<uint4*>(set + 4 * i) = make_uint4(0, 0, 0, 0);
    for (uint32_t i = threadIdx.x; i < (uint32_t)DE_HIST_WORDS; i += DE_THREADS) de_smem[DE_DATA_WORDS + i] = 0;
    if (g.n_words) {
      for (uint32_t i = real_quads * 4u + threadIdx.x; i < g.n_words; i += DE_THREADS) data[i] = 0;  // the window's overhang
      asm volatile(
          "{\n\t"
          ".reg .pred P1;\n\t"
          "DE_WAIT:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
          "@P1 bra DE_DONE;\n\t"
          "bra DE_WAIT;\n\t"
          "DE_DONE:\n\t"
          "}" ::"r"(smem_u32(&s_mbar)),
          "r"(parity)
          : "memory");
      parity ^= 1u;
    }
    __syncthreads();
    // the next item's ticket and its search run behind the other warps' work on this one
    if (threadIdx.x == 0) de_fetch(item_ptr, n_reads, n_items, counters, &s_read[cur ^ 1], &s_pass[cur ^ 1], &s_npass[cur ^ 1]);
    if (g.n_words) {
      de_phase<true, false>(data, set, nb, nk, k, pass, n_pass, counters);
      __syncthreads();
      de_phase<true, true>(data, set, nb, nk, k, pass, n_pass, counters);
      __syncthreads();
      de_scan_emit<true>(data, set, nb, nk, k, hist, s_base, records, part_cap, cursors, counters);
    } else {
      de_phase<false, false>(gwords, set, nb, nk, k, pass, n_pass, counters);
      __syncthreads();
      de_phase<false, true>(gwords, set, nb, nk, k, pass, n_pass, counters);
      __syncthreads();
      de_scan_emit<false>(gwords, set, nb, nk, k, hist, s_base, records, part_cap, cursors, counters);
    }
  }
}

// ---- phase 2 ------------------------------------------------------------------------------------------------
// Blocks take chunks of DA_CHUNK records in partition order from one ticket counter, so at any moment the whole grid
// works inside one or two neighbouring partitions = one or two 1/DE_PARTS windows of the table.  Per record: the CAS
// claim of the home slot is the probe (four in flight per thread), then ONE 64-bit add on the counter word
// { n_reads low, n_multi high }: + 1 and, for a "more than once in this read" record, + 2^32.
constexpr int DA_THREADS = 256;
constexpr int DA_BLOCKS_PER_SM = 4;
constexpr int DA_PER = 4;
constexpr int DA_CHUNK = DA_THREADS * DA_PER;

__device__ __forceinline__ int64_t da_upsert_from(uint64_t* table, int64_t cap, uint64_t key, int64_t slot) {
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

__global__ void __launch_bounds__(DA_THREADS, DA_BLOCKS_PER_SM)
docfreq_apply_kernel(const uint64_t* __restrict__ records, int64_t part_cap, const int64_t* __restrict__ cursors, int k,
                     uint64_t* table, int64_t cap, int64_t* counters) {
  __shared__ long long s_pref[DE_PARTS + 1];  // chunks before partition p
  __shared__ long long s_ticket;
  for (int p = threadIdx.x; p < DE_PARTS; p += DA_THREADS) {
    const long long n = min((long long)__ldg(cursors + p), (long long)part_cap);
    s_pref[p + 1] = (n + DA_CHUNK - 1) / DA_CHUNK;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    s_pref[0] = 0;
    for (int p = 1; p <= DE_PARTS; ++p) {
      run += s_pref[p];
      s_pref[p] = run;
    }
  }
  __syncthreads();
  const long long n_chunks = s_pref[DE_PARTS];
  const uint64_t mask = (k < 32) ? ((1ull << (2 * k)) - 1) : ~0ull;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = (long long)atomicAdd((unsigned long long*)(counters + 3), 1ull);
    __syncthreads();
    const long long t = s_ticket;
    if (t >= n_chunks) break;
    int lo = 0, hi = DE_PARTS;  // s_pref[lo] <= t < s_pref[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_pref[mid] > t) hi = mid; else lo = mid;
    }
    const int64_t n_part = min((long long)__ldg(cursors + lo), (long long)part_cap);
    const int64_t first = (int64_t)(t - s_pref[lo]) * DA_CHUNK;
    const uint64_t* src = records + (int64_t)lo * part_cap + first;
    const int n = (int)min((int64_t)DA_CHUNK, n_part - first);
    uint64_t key[DA_PER], old[DA_PER], inc[DA_PER];
    int64_t slot[DA_PER];
#pragma unroll
    for (int j = 0; j < DA_PER; ++j) {
      const int i = threadIdx.x + j * DA_THREADS;
      key[j] = EMPTY;
      if (i < n) {
        const uint64_t rec = __ldcs(src + i);  // read once: do not let the stream push the table window out of L2
        key[j] = rec & mask;
        inc[j] = 1ull + ((rec >> 63) << 32);
        slot[j] = home_slot(mix64(key[j]), cap);
        old[j] = atomicCAS((unsigned long long*)(table + 2 * slot[j]), (unsigned long long)EMPTY,
                           (unsigned long long)key[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < DA_PER; ++j) {
      if (key[j] == EMPTY) continue;
      int64_t sl = slot[j];
      if (old[j] != EMPTY && old[j] != key[j]) sl = da_upsert_from(table, cap, key[j], sl + 1);
      if (sl < 0) counters[0] = 1;
      else atomicAdd((unsigned long long*)(table + 2 * sl + 1), (unsigned long long)inc[j]);
    }
  }
}

}  // namespace

extern "C" {

int cfk_docfreq_parts(void) { return DE_PARTS; }

int cfk_docfreq_emit_plan(const int64_t* read_len, const int32_t* order, int64_t n_reads, int k, int32_t* n_pass,
                          cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_docfreq_emit_plan: k must be in [1, 31]");
  if (n_reads < 0) return fail(CFK_ERR_INVALID, "cfk_docfreq_emit_plan: bad sizes");
  if (n_reads == 0) return CFK_OK;
  docfreq_emit_plan_kernel<<<(unsigned)blocks_for(n_reads, 256), 256, 0, (cudaStream_t)stream>>>(read_len, order, n_reads, k,
                                                                                                  n_pass);
  CFK_CHECK_LAUNCH("docfreq_emit_plan_kernel", 1);
  return CFK_OK;
}

int cfk_docfreq_emit(const uint32_t* packed, const int64_t* read_off, const int64_t* read_len, const int32_t* order,
                     const int64_t* item_ptr, int64_t n_reads, int k, uint64_t* records, int64_t part_cap,
                     int64_t* cursors, int64_t* counters, int32_t n_blocks, cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_docfreq_emit: k must be in [1, 31]");
  if (part_cap < 1 || n_reads < 0 || n_blocks < 1) return fail(CFK_ERR_INVALID, "cfk_docfreq_emit: bad sizes");
  if (n_reads == 0) return CFK_OK;
  static unsigned long long attr_done = 0;
  const int smem = DE_SMEM_WORDS * 4;
  {
    cudaError_t e = ensure_dynamic_smem(docfreq_emit_kernel, smem, &attr_done);
    if (e != cudaSuccess) return fail(CFK_ERR_CUDA, "cfk_docfreq_emit: cudaFuncSetAttribute", e);
  }
  docfreq_emit_kernel<<<(unsigned)n_blocks, DE_THREADS, smem, (cudaStream_t)stream>>>(
      packed, read_off, read_len, order, item_ptr, n_reads, k, records, part_cap, cursors, counters);
  CFK_CHECK_LAUNCH("docfreq_emit_kernel", 1);
  return CFK_OK;
}

int cfk_docfreq_apply(const uint64_t* records, int64_t part_cap, const int64_t* cursors, int k, uint64_t* table, int64_t cap,
                      int64_t* counters, int32_t n_blocks, cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_docfreq_apply: k must be in [1, 31]");
  if (part_cap < 1 || cap < 1 || n_blocks < 1) return fail(CFK_ERR_INVALID, "cfk_docfreq_apply: bad sizes");
  docfreq_apply_kernel<<<(unsigned)n_blocks * DA_BLOCKS_PER_SM, DA_THREADS, 0, (cudaStream_t)stream>>>(records, part_cap, cursors, k, table,
                                                                                       cap, counters);
  CFK_CHECK_LAUNCH("docfreq_apply_kernel", 1);
  return CFK_OK;
}

}  // extern "C"
