// docfreq_stream.cu — stage A (document frequency), two-phase form: the default path of the engine.
//
// Replaces get_kmer_freqs_from_ncrf_report + the band of get_rare_kmers,
// scripts/distance_based_kmer_recruitment.py:39-63 and :74-79.
//
//   phase 1  docfreq_emit_kernel   per (read, pass) item: the packed read is brought into shared memory by ONE
//            bulk copy (cp.async.bulk + mbarrier -- the TMA engine, no register round trip), the read's k-mers are
//            de-duplicated in a shared-memory set WITHOUT atomics (claim with plain stores, verify after a
//            barrier), and one 8-byte record per DISTINCT k-mer of the read -- bit 63 = "occurs more than once in
//            this read" -- is appended to one of n_parts hash partitions in HBM.  The set is ordered by the same
//            hash as the partitions, so the final scan of the set meets the partitions in order and neighbouring
//            records share a cursor bump and an L2 line.
//   phase 2  docfreq_count_kernel  one block per partition: the partition's records stream through a
//            shared-memory table (again claim / verify, atomics only for the few k-mers seen in several reads),
//            n_reads / n_multi are final when the partition ends, so the band filter of get_rare_kmers runs right
//            there; the complete (k-mer, n_reads, n_multi) table is written out (dense, no empty slots) only
//            when the caller asks for it.
//
// No global hash table, no table initialisation pass, no separate select pass: HBM sees the packed reads once,
// the records once out and once in, and the results.  Partitions are also the unit of the multi-GPU exchange
// (records of partition range g go to rank g; phase 2 then takes one record run per source rank).
#include "cfk_common.cuh"

namespace {

using namespace cfk;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t pick4(const uint4& v, int s) { return s == 0 ? v.x : s == 1 ? v.y : s == 2 ? v.z : v.w; }

// ---- the shared-memory sets of both phases: 4-slot buckets (one 16-byte load), linear probing over buckets --------
// Slot (32 bit): bits [id_bits-1:0] = id of the entry's owner + 1 (0 = empty slot), the bits above up to bit 30 =
// fingerprint of the key, bit 31 = a flag of the user.  An empty slot is claimed with atomicCAS, so a key is in the
// set exactly once; a walker that loses the CAS looks at the same bucket again (the winner may hold its key).
// Slots never return to empty and a writer takes the FIRST empty slot it sees: a probe chain has no holes.

// ================================================================================================================
// phase 1
// ================================================================================================================
#ifndef CFK_DE_FILL_PCT
#define CFK_DE_FILL_PCT 85   /* planned load of the per-read set, percent (4-slot buckets) */
#endif
#ifndef CFK_DE_THREADS
#define CFK_DE_THREADS 1024
#endif
constexpr int DE_THREADS = CFK_DE_THREADS;
constexpr int DE_WARPS = DE_THREADS / 32;
#ifndef CFK_DE_PER
#define CFK_DE_PER 4
#endif
constexpr int DE_PER = CFK_DE_PER;          // consecutive k-mer starts per lane (2 or 4: the window holds 14 + 3 + 30 bases)
constexpr int DE_CHUNK = 32 * DE_PER;       // k-mer starts per warp step
constexpr int DE_BLOCKS = 1024 / DE_THREADS;  // blocks sharing an SM (and its shared memory)
constexpr int DE_DATA_WORDS = DE_BLOCKS == 1 ? 57344 : 57344 / DE_BLOCKS - 512;  // dynamic shared memory of a block: [ read words | set | queues ]
constexpr int DE_MIN_SET = 16384 / DE_BLOCKS;
constexpr uint32_t DE_MULTI = 0x80000000u;
constexpr int DE_Q = 64;                          // queue entries per warp
constexpr int DE_Q_WORDS = DE_WARPS * DE_Q * 2;   // the queues sit behind the set
static_assert(DE_DATA_WORDS > 2 * DE_MIN_SET, "shared memory budget");

struct DeGeometry {
  uint32_t n_words;    // words staged (0: the read stays in global memory)
  uint32_t n_buckets;  // 4-slot buckets of the set
  uint32_t fill;       // k-mers planned per pass
};

__host__ __device__ __forceinline__ DeGeometry de_geometry(int64_t len) {
  DeGeometry g;
  const int64_t nw = (((len + 15) >> 4) + 3 + 3) & ~(int64_t)3;  // + the 3-word extraction window, multiple of 4
  g.n_words = (nw <= DE_DATA_WORDS - DE_Q_WORDS - DE_MIN_SET) ? (uint32_t)nw : 0u;
  g.n_buckets = ((uint32_t)DE_DATA_WORDS - DE_Q_WORDS - g.n_words) >> 2;
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

// Inside phase 1 a k-mer is handled in the order the packed read holds it: base q + j at bits 2j..2j+1 ("raw").  That
// is a bijection of the k-mer, so hashing and comparing raw forms is exact; only a record that leaves the SM is turned
// into the library's key form (first base most significant).
// raw k-mer starting at base q; words[] must be readable up to word (q >> 4) + 2
template <bool IN_SMEM>
__device__ __forceinline__ uint64_t de_raw_at(const uint32_t* words, uint32_t q, uint64_t mask) {
  const uint32_t w = q >> 4, sh = (q & 15u) << 1;
  const uint32_t a = de_word<IN_SMEM>(words, w), b = de_word<IN_SMEM>(words, w + 1), c = de_word<IN_SMEM>(words, w + 2);
  const uint32_t lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);  // sh < 32
  return ((uint64_t)lo | ((uint64_t)hi << 32)) & mask;
}

__device__ __forceinline__ uint64_t de_key_of_raw(uint64_t raw, int k) {
  uint64_t r = __brevll(raw);
  r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);
  return r >> (64 - 2 * k);
}

// The one hash of phase 1.  g orders everything: pass = floor(g * n_pass / 2^32), set bucket = the fraction of that
// product scaled to the set, partition = floor(g * n_parts / 2^32); the fingerprint takes g's low bits.
__device__ __forceinline__ uint32_t de_hash(uint64_t kmer) {
  uint32_t g = (uint32_t)(kmer ^ (kmer >> 32)) * 0x85EBCA6Bu;
  g ^= g >> 13;
  g *= 0xC2B2AE35u;
  g ^= g >> 16;
  return g;
}

// ---- one walk per k-mer, probes issued from a warp-private queue --------------------------------------------------
// A k-mer start that belongs to this pass becomes a queue item { position | flags, bucket to look at next }.  Whenever
// 32 items wait, the warp takes them, one per lane: re-extracts the k-mer, looks at one bucket, and either finishes
// (the k-mer is there: set "again"; or an empty slot was claimed with atomicCAS) or puts the item back with its next
// bucket.  Every probe runs on full warps, whatever share of the positions the pass selects and however long single
// chains get.
template <bool IN_SMEM>
__device__ __forceinline__ int de_drain(const uint32_t* words, uint32_t* set, uint32_t nb, uint64_t mask, int pos_bits,
                                        uint2* q, int qn) {
  const int lane = threadIdx.x & 31;
  const uint32_t pos_mask = (1u << pos_bits) - 1u;
  const uint32_t fp_mask = 0x7FFFFFFFu & ~pos_mask;
  const int n = min(32, qn), first = qn - n;
  bool more = false;
  uint2 it = make_uint2(0, 0);
  if (lane < n) {
    it = q[first + lane];
    const uint32_t fresh = it.x;  // position + 1 | fingerprint
    uint32_t* bucket = set + 4 * it.y;
    const uint4 v4 = *reinterpret_cast<const uint4*>(bucket);
    unsigned e = 0, m = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const uint32_t x = pick4(v4, s);
      if (x == 0) e |= 1u << s;
      else if (((x ^ fresh) & fp_mask) == 0) m |= 1u << s;
    }
    bool done = false;
    if (m) {
      const uint64_t raw = de_raw_at<IN_SMEM>(words, (fresh & pos_mask) - 1u, mask);
      do {
        const int s = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t x = pick4(v4, s);
        if (de_raw_at<IN_SMEM>(words, (x & pos_mask) - 1u, mask) == raw) {
          if (!(x & DE_MULTI)) bucket[s] = x | DE_MULTI;  // the k-mer occurs again in this read (idempotent)
          done = true;
          break;
        }
      } while (m);
    }
    if (!done) {
      if (e) {
        done = atomicCAS(bucket + (__ffs(e) - 1), 0u, fresh) == 0u;  // lost: look at this bucket again
      } else {
        it.y = (it.y + 1 == nb) ? 0u : it.y + 1;
      }
    }
    more = !done;
  }
  const unsigned mm = __ballot_sync(FULL, more);
  __syncwarp();
  if (more) q[first + __popc(mm & ((1u << lane) - 1u))] = it;
  __syncwarp();
  return first + __popc(mm);
}

template <bool IN_SMEM>
__device__ __forceinline__ void de_insert_q(const uint32_t* words, uint32_t* set, uint32_t nb, int64_t nk, int k, uint32_t pass,
                                            uint32_t n_pass, uint2* queues, int64_t* counters) {
  const uint64_t mask = (1ull << (2 * k)) - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pos_bits = 64 - __clzll((unsigned long long)(nk + 1));  // pos + 1 <= nk < 2^pos_bits - 1
  const uint32_t fp_mask = 0x7FFFFFFFu & ~((1u << pos_bits) - 1u);  // nk < 2^30 (checked on the host)
  uint2* q = queues + warp * DE_Q;
  const unsigned lt = (1u << lane) - 1u;
  int qn = 0;
  for (int64_t chunk0 = (int64_t)warp * DE_CHUNK; chunk0 < nk; chunk0 += (int64_t)DE_WARPS * DE_CHUNK) {
    const int64_t base = chunk0 + lane * DE_PER;
    const uint32_t p0 = (uint32_t)base;  // multiple of DE_PER
    const int npos = base < nk ? (int)min((int64_t)DE_PER, nk - base) : 0;
    uint32_t home[DE_PER], fresh[DE_PER];
    unsigned act = 0;
    if (npos > 0) {
      // 48-base window from the word of p0: offset + DE_PER - 1 + k <= 12 + 3 + 31 < 48
      const uint32_t w0 = p0 >> 4;
      uint64_t win_lo = (uint64_t)de_word<IN_SMEM>(words, w0) | ((uint64_t)de_word<IN_SMEM>(words, w0 + 1) << 32);
      uint32_t win_hi = de_word<IN_SMEM>(words, w0 + 2);
      if (const uint32_t sh0 = (p0 & 15u) << 1) {
        win_lo = (win_lo >> sh0) | ((uint64_t)win_hi << (64 - sh0));
        win_hi >>= sh0;
      }
#pragma unroll
      for (int j = 0; j < DE_PER; ++j) {
        const uint32_t g = de_hash(win_lo & mask);  // the raw k-mer at p0 + j
        win_lo = (win_lo >> 2) | ((uint64_t)win_hi << 62);
        win_hi >>= 2;
        const uint64_t t = (uint64_t)g * n_pass;
        home[j] = __umulhi((uint32_t)t, nb);
        fresh[j] = (p0 + (uint32_t)j + 1u) | ((g << pos_bits) & fp_mask);
        if (j < npos && (uint32_t)(t >> 32) == pass) act |= 1u << j;
      }
    }
#pragma unroll
    for (int j = 0; j < DE_PER; ++j) {
      const bool a = (act >> j) & 1u;
      const unsigned m = __ballot_sync(FULL, a);
      if (m == 0) continue;
      if (a) q[qn + __popc(m & lt)] = make_uint2(fresh[j], home[j]);
      qn += __popc(m);
      __syncwarp();
      while (qn >= 32) qn = de_drain<IN_SMEM>(words, set, nb, mask, pos_bits, q, qn);
    }
  }
  uint32_t rounds = 0;
  while (qn > 0) {
    qn = de_drain<IN_SMEM>(words, set, nb, mask, pos_bits, q, qn);
    if (++rounds > 4 * nb + 64) {  // cannot happen: a pass is planned for <= CFK_DE_FILL_PCT % load
      counters[1] = 1;
      break;
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

// Final scan of the set, one bucket per thread and step: every live slot becomes one record in the partition of
// its k-mer.  The set is in hash order, so the (up to four) records of a bucket mostly share a partition: one
// cursor bump per run of equal partitions.
template <bool IN_SMEM>
__device__ __forceinline__ void de_scan_emit(const uint32_t* words, const uint32_t* set, uint32_t nb, int64_t nk, int k,
                                             uint64_t* __restrict__ records, uint32_t part_cap, uint32_t n_parts,
                                             uint32_t* cursors, int64_t* counters) {
  const uint64_t mask = (1ull << (2 * k)) - 1;
  const int pos_bits = 64 - __clzll((unsigned long long)(nk + 1));
  const uint32_t pos_mask = (1u << pos_bits) - 1u;
  const bool small = (uint64_t)n_parts * part_cap < (1ull << 32);
  for (uint32_t b = threadIdx.x; b < nb; b += DE_THREADS) {
    const uint4 v4 = *reinterpret_cast<const uint4*>(set + 4 * b);
    uint64_t rec[4];
    uint32_t part[4], idx[4];
    unsigned live = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const uint32_t v = pick4(v4, s);
      part[s] = 0xFFFFFFFFu;
      rec[s] = 0;
      if (v != 0) {
        const uint64_t raw = de_raw_at<IN_SMEM>(words, (v & pos_mask) - 1u, mask);
        part[s] = __umulhi(de_hash(raw), n_parts);
        rec[s] = raw | ((uint64_t)(v >> 31) << 63);  // records keep the raw order: phase 2 turns distinct k-mers into keys
        live |= 1u << s;
      }
    }
    if (!live) continue;
    // runs of equal partitions among neighbouring live slots: the leader bumps the cursor for the run
    const bool f1 = part[1] == part[0] && (live & 3u) == 3u;
    const bool f2 = part[2] == part[1] && (live & 6u) == 6u;
    const bool f3 = part[3] == part[2] && (live & 12u) == 12u;
    const uint32_t len0 = 1u + (f1 ? 1u + (f2 ? 1u + (f3 ? 1u : 0u) : 0u) : 0u);
    const uint32_t len1 = 1u + (f2 ? 1u + (f3 ? 1u : 0u) : 0u);
    const uint32_t len2 = 1u + (f3 ? 1u : 0u);
    idx[0] = idx[1] = idx[2] = idx[3] = 0;
    if (live & 1u) idx[0] = atomicAdd(cursors + part[0], len0);
    if ((live & 2u) && !f1) idx[1] = atomicAdd(cursors + part[1], len1);
    if ((live & 4u) && !f2) idx[2] = atomicAdd(cursors + part[2], len2);
    if ((live & 8u) && !f3) idx[3] = atomicAdd(cursors + part[3], 1u);
    if (f1) idx[1] = idx[0] + 1;
    if (f2) idx[2] = idx[1] + 1;
    if (f3) idx[3] = idx[2] + 1;
    bool over = false;
    if (small) {  // the whole record array has fewer than 2^32 entries: 32-bit offsets
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (!((live >> s) & 1u)) continue;
        if (idx[s] < part_cap) records[part[s] * part_cap + idx[s]] = rec[s];
        else over = true;
      }
    } else {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (!((live >> s) & 1u)) continue;
        if (idx[s] < part_cap) records[(uint64_t)part[s] * part_cap + idx[s]] = rec[s];
        else over = true;
      }
    }
    if (over) counters[0] = 1;  // partition buffer full: the host falls back (never silent)
  }
}

__global__ void __launch_bounds__(DE_THREADS, DE_BLOCKS)
docfreq_emit_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ read_off,
                    const int64_t* __restrict__ read_len, const int32_t* __restrict__ order,
                    const int64_t* __restrict__ item_ptr, int64_t n_reads, int k, uint64_t* __restrict__ records,
                    int64_t part_cap, uint32_t n_parts, uint32_t* cursors, int64_t* counters) {
  extern __shared__ __align__(16) uint32_t de_smem[];
  uint32_t* data = de_smem;  // [ read words | set ]
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
    uint2* queues = reinterpret_cast<uint2*>(data + (DE_DATA_WORDS - DE_Q_WORDS));
    if (g.n_words) {
      de_insert_q<true>(data, set, nb, nk, k, pass, n_pass, queues, counters);
      __syncthreads();
      de_scan_emit<true>(data, set, nb, nk, k, records, (uint32_t)part_cap, n_parts, cursors, counters);
    } else {
      de_insert_q<false>(gwords, set, nb, nk, k, pass, n_pass, queues, counters);
      __syncthreads();
      de_scan_emit<false>(gwords, set, nb, nk, k, records, (uint32_t)part_cap, n_parts, cursors, counters);
    }
  }
}

// ================================================================================================================
// phase 2
// ================================================================================================================
// One block per UNIT of `group` neighbouring partitions (tickets; neighbouring partitions are neighbouring hash
// ranges, so their union is a partition too).  Shared memory holds the unit's table: keys[] (64 bit: the record that
// claimed the slot with atomicCAS -- k-mer in bits 0..61, bit 63 = that read held it more than once; all ones =
// empty) and one 32-bit word per slot for the k-mers seen in more than one read: low half = further reads, high
// half = how many of those held the k-mer more than once.  The owner's read is implicit (n_reads = 1 + low half,
// n_multi = bit 63 + high half).  Records stream through registers, four in flight per thread, with no barrier
// between the unit's first record and its last; buckets are two keys (one 16-byte load), linear probing over
// buckets.  A unit with more distinct k-mers than the table takes is done again partition by partition.
#ifndef CFK_CN_THREADS
#define CFK_CN_THREADS 256
#endif
constexpr int CN_THREADS = CFK_CN_THREADS;  // 1024 / CN_THREADS blocks share an SM
constexpr int CN_PER = 4;
constexpr int CN_CHUNK = CN_THREADS * CN_PER;  // records in flight per round
constexpr int CN_DCAP = CFK_DOCFREQ_PART_DISTINCT;  // distinct k-mers a unit may hold
constexpr int CN_NB = CN_DCAP * 2 / 3;         // 2-key buckets: at most 75 % load
constexpr uint64_t CN_KEY = 0x3FFFFFFFFFFFFFFFull;
constexpr int CN_SMEM_BYTES = CN_NB * 2 * 12;
constexpr int CN_MAX_PROBES = 96;              // buckets one record may visit before the unit is declared full
static_assert(CN_SMEM_BYTES <= 227 * 1024 / (1024 / CN_THREADS) - 1024, "shared memory budget");

__device__ __forceinline__ uint32_t cn_hash(uint64_t key) {
  uint32_t x = (uint32_t)key * 0x9E3779B1u + (uint32_t)(key >> 32) * 0x85EBCA77u;
  x ^= x >> 15;
  x *= 0x2C1B3C6Du;
  x ^= x >> 13;
  return x;
}

// exclusive prefix sum of one int per thread across the block; also returns the total (same for every thread)
__device__ __forceinline__ int cn_block_scan(int v, int* total, int* s_warp /* [33] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = lane < CN_THREADS / 32 ? s_warp[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;
    if (lane == 31) s_warp[32] = wi;
  }
  __syncthreads();
  const int res = s_warp[warp] + incl - v;
  *total = s_warp[32];
  __syncthreads();
  return res;
}

struct CountArgs {
  const uint64_t* records;
  const uint32_t* cursors;
  const int64_t* offsets;  // not NULL: records of (source, partition) start at offsets[source * n_parts + partition] (dense runs)
  int64_t part_cap, n_parts, src_stride;
  int32_t n_src, group, k;
  uint32_t lo, hi, max_nonuniq;
  uint64_t* rare_keys;
  uint32_t* rare_nreads;
  uint32_t* rare_nmulti;
  int64_t max_rare;
  uint4* dense;
  int64_t max_dense;
  int64_t* counters;
};

__device__ __forceinline__ int64_t cn_run_len(const CountArgs& A, int32_t src, int64_t p) {
  const int64_t n = (int64_t)__ldg(A.cursors + (int64_t)src * A.n_parts + p);
  return A.offsets != nullptr ? n : min(n, A.part_cap);
}

struct CountSmem {
  uint64_t* keys;
  uint32_t* cnt;
  int* s_warp;
  int* s_abort;
  long long* s_base;
};

// partitions [p0, p1) as one unit, restricted to the records whose hash has `sub` in its low bits (n_sub a power of
// two; 0, 1 = all records); false: more distinct k-mers than the table holds (nothing was written)
__device__ bool cn_unit(const CountArgs& A, const CountSmem& S, int64_t p0, int64_t p1, uint32_t sub, uint32_t n_sub) {
  const int lane = threadIdx.x & 31;
  int64_t n_unit = 0;  // records of the unit
  for (int64_t p = p0; p < p1; ++p)
    for (int32_t src = 0; src < A.n_src; ++src)
      n_unit += cn_run_len(A, src, p);
  // the table is sized for the unit: at most 75 % full even if every record is a new k-mer
  const uint32_t nb = (uint32_t)min((int64_t)CN_NB, max((int64_t)64, n_unit * 2 / 3 + 1));  // (also right for a sub-range)
  for (uint32_t i = threadIdx.x; i < nb; i += CN_THREADS) {
    reinterpret_cast<uint4*>(S.keys)[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    reinterpret_cast<uint2*>(S.cnt)[i] = make_uint2(0, 0);
  }
  if (threadIdx.x == 0) *S.s_abort = 0;
  __syncthreads();
  for (int64_t p = p0; p < p1; ++p) {
    for (int32_t src = 0; src < A.n_src; ++src) {
      const int64_t n = cn_run_len(A, src, p);
      const uint64_t* base = A.offsets != nullptr ? A.records + __ldg(A.offsets + (int64_t)src * A.n_parts + p)
                                                  : A.records + (int64_t)src * A.src_stride + p * A.part_cap;
      for (int64_t c0 = 0; c0 < n; c0 += CN_CHUNK) {
        if (*(volatile int*)S.s_abort) break;  // no barrier in here: every thread still reaches the one below
        uint64_t rec[CN_PER];
        uint32_t b[CN_PER];
        unsigned pend = 0;
#pragma unroll
        for (int j = 0; j < CN_PER; ++j) {
          const int64_t i = c0 + threadIdx.x + j * CN_THREADS;
          rec[j] = 0;
          if (i < n) {
            rec[j] = __ldcs(base + i);
            const uint32_t x = cn_hash(rec[j] & CN_KEY);
            b[j] = __umulhi(x, nb);
            if ((x & (n_sub - 1u)) == sub) pend |= 1u << j;
          }
        }
        // one walk per record: the k-mer is there -> its read goes to the slot's counter; else claim an empty slot
        uint32_t rounds = 0;
        while (pend) {
#pragma unroll
          for (int j = 0; j < CN_PER; ++j) {
            if (!((pend >> j) & 1u)) continue;
            const uint4 v4 = reinterpret_cast<const uint4*>(S.keys)[b[j]];
            const uint64_t key = rec[j] & CN_KEY;
            const uint64_t k0 = (uint64_t)v4.x | ((uint64_t)v4.y << 32), k1 = (uint64_t)v4.z | ((uint64_t)v4.w << 32);
            int hit = -1;  // slot of the bucket that holds the k-mer
            if ((k0 & CN_KEY) == key) hit = 0;
            else if (k0 == EMPTY) {
              const uint64_t old = atomicCAS((unsigned long long*)(S.keys + 2 * b[j]), (unsigned long long)EMPTY,
                                             (unsigned long long)rec[j]);
              if (old == EMPTY) {
                pend &= ~(1u << j);  // first read with this k-mer: the slot's owner
                continue;
              }
              if ((old & CN_KEY) == key) hit = 0;
            }
            if (hit < 0) {
              if ((k1 & CN_KEY) == key) hit = 1;
              else if (k1 == EMPTY) {
                const uint64_t old = atomicCAS((unsigned long long*)(S.keys + 2 * b[j] + 1), (unsigned long long)EMPTY,
                                               (unsigned long long)rec[j]);
                if (old == EMPTY) {
                  pend &= ~(1u << j);
                  continue;
                }
                if ((old & CN_KEY) == key) hit = 1;
              }
            }
            if (hit >= 0) {
              const uint32_t old = atomicAdd(S.cnt + 2 * b[j] + hit, 1u + (uint32_t)(rec[j] >> 63 << 16));
              if ((old & 0xFFFFu) == 0xFFFFu) A.counters[0] = 2;  // 65536 further reads: the 16-bit halves are exhausted
              pend &= ~(1u << j);
            } else {
              b[j] = (b[j] + 1 == nb) ? 0u : b[j] + 1;
            }
          }
          if (++rounds > CN_MAX_PROBES) {  // the table is (nearly) full
            *S.s_abort = 1;
            break;
          }
        }
      }
    }
  }
  __syncthreads();
  if (*S.s_abort) return false;  // block-uniform: read behind the barrier, not written after it
  // ---- output: every occupied slot is one distinct k-mer with its final counts
  const uint32_t n_slots = 2 * nb;
  int n_live = 0;
  for (uint32_t s0 = 0; s0 < n_slots; s0 += CN_THREADS) {  // block-uniform trip count
    const uint32_t s = s0 + threadIdx.x;
    const uint64_t r = s < n_slots ? S.keys[s] : EMPTY;
    bool take = false;
    uint32_t nr = 0, nm = 0;
    if (r != EMPTY) {
      ++n_live;
      const uint32_t cw = S.cnt[s];
      nr = 1u + (cw & 0xFFFFu);
      nm = (uint32_t)(r >> 63) + (cw >> 16);
      take = A.rare_keys != nullptr && nm <= A.max_nonuniq && nr >= A.lo && nr <= A.hi;
    }
    const unsigned m = __ballot_sync(FULL, take);
    if (m == 0) continue;
    long long at = 0;
    if (lane == __ffs(m) - 1) at = (long long)atomicAdd((unsigned long long*)(A.counters + 4), (unsigned long long)__popc(m));
    at = __shfl_sync(FULL, at, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u));
    if (take && at < A.max_rare) {
      A.rare_keys[at] = de_key_of_raw(r & CN_KEY, A.k);
      if (A.rare_nreads != nullptr) A.rare_nreads[at] = nr;
      if (A.rare_nmulti != nullptr) A.rare_nmulti[at] = nm;
    }
  }
  {
    int total = 0;
    int rank = cn_block_scan(n_live, &total, S.s_warp);
    if (threadIdx.x == 0) *S.s_base = (long long)atomicAdd((unsigned long long*)(A.counters + 5), (unsigned long long)total);
    if (A.dense != nullptr) {
      __syncthreads();
      long long at = *S.s_base + rank;
      for (uint32_t s = threadIdx.x; s < n_slots; s += CN_THREADS) {
        const uint64_t r = S.keys[s];
        if (r == EMPTY) continue;
        const uint32_t cw = S.cnt[s];
        const uint64_t key = de_key_of_raw(r & CN_KEY, A.k);
        if (at < A.max_dense)
          A.dense[at] = make_uint4((uint32_t)key, (uint32_t)(key >> 32), 1u + (cw & 0xFFFFu), (uint32_t)(r >> 63) + (cw >> 16));
        ++at;
      }
    }
  }
  __syncthreads();  // the table is cleared by the next unit
  return true;
}

__global__ void __launch_bounds__(CN_THREADS, 1024 / CN_THREADS) docfreq_count_kernel(const CountArgs A) {
  extern __shared__ __align__(16) uint32_t cn_smem[];
  __shared__ int s_warp[33];
  __shared__ long long s_ticket, s_base;
  __shared__ int s_abort;
  CountSmem S;
  S.keys = reinterpret_cast<uint64_t*>(cn_smem);
  S.cnt = cn_smem + 4 * CN_NB;
  S.s_warp = s_warp;
  S.s_abort = &s_abort;
  S.s_base = &s_base;
  const int64_t n_units = (A.n_parts + A.group - 1) / A.group;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = (long long)atomicAdd((unsigned long long*)(A.counters + 3), 1ull);
    __syncthreads();
    const int64_t u = s_ticket;
    if (u >= n_units) break;
    const int64_t p0 = u * A.group, p1 = min(p0 + (int64_t)A.group, A.n_parts);
    if (cn_unit(A, S, p0, p1, 0u, 1u)) continue;
    // too many distinct k-mers for one table: partition by partition, and a partition that still does not fit by
    // halves of its hash range (every record is read once per attempt; the results do not depend on the split)
    if (threadIdx.x == 0) atomicAdd((unsigned long long*)(A.counters + 6), 1ull);
    bool ok = true;
    for (int64_t p = p0; p < p1 && ok; ++p) {
      __syncthreads();
      if (p1 - p0 > 1 && cn_unit(A, S, p, p + 1, 0u, 1u)) continue;
      uint32_t st_sub[8], st_n[8];
      int top = 0;
      st_sub[top] = 0u, st_n[top++] = 2u;
      st_sub[top] = 1u, st_n[top++] = 2u;
      while (top > 0 && ok) {
        --top;
        const uint32_t sub = st_sub[top], n_sub = st_n[top];
        __syncthreads();
        if (cn_unit(A, S, p, p + 1, sub, n_sub)) continue;
        if (n_sub >= 64u) {
          ok = false;
        } else {
          st_sub[top] = sub, st_n[top++] = 2u * n_sub;
          st_sub[top] = sub + n_sub, st_n[top++] = 2u * n_sub;
        }
      }
    }
    if (ok) continue;
    if (threadIdx.x == 0) A.counters[0] = 1;  // one partition holds more distinct k-mers than planned: the host falls back
  }
}

// records[p * part_cap + i], i < cursors[p]  ->  out[offsets[p] + i]: the partitions back to back (the send buffer of
// the multi-GPU exchange; offsets = exclusive scan of the cursors).  One block per partition.
__global__ void __launch_bounds__(256) records_pack_kernel(const uint64_t* __restrict__ records, int64_t part_cap,
                                                           const uint32_t* __restrict__ cursors,
                                                           const int64_t* __restrict__ offsets, int64_t n_parts,
                                                           uint64_t* __restrict__ out) {
  for (int64_t p = blockIdx.x; p < n_parts; p += gridDim.x) {
    const int64_t n = min((int64_t)__ldg(cursors + p), part_cap);
    const uint64_t* src = records + p * part_cap;
    uint64_t* dst = out + __ldg(offsets + p);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldcs(src + i);
  }
}

}  // namespace

extern "C" {

int cfk_docfreq_part_target(void) { return 4608; }
int cfk_docfreq_part_distinct(void) { return CN_DCAP; }

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
                     int64_t n_parts, uint32_t* cursors, int64_t* counters, int32_t n_blocks, cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_docfreq_emit: k must be in [1, 31]");
  if (part_cap < 1 || part_cap > 0x7FFFFFFF || n_parts < 1 || n_parts > 0x7FFFFFFF || n_reads < 0 || n_blocks < 1)
    return fail(CFK_ERR_INVALID, "cfk_docfreq_emit: bad sizes");
  if (n_reads == 0) return CFK_OK;
  static unsigned long long attr_done = 0;
  const int smem = DE_DATA_WORDS * 4;
  {
    cudaError_t e = ensure_dynamic_smem(docfreq_emit_kernel, smem, &attr_done);
    if (e != cudaSuccess) return fail(CFK_ERR_CUDA, "cfk_docfreq_emit: cudaFuncSetAttribute", e);
  }
  docfreq_emit_kernel<<<(unsigned)n_blocks * DE_BLOCKS, DE_THREADS, smem, (cudaStream_t)stream>>>(
      packed, read_off, read_len, order, item_ptr, n_reads, k, records, part_cap, (uint32_t)n_parts, cursors, counters);
  CFK_CHECK_LAUNCH("docfreq_emit_kernel", 1);
  return CFK_OK;
}

int cfk_records_pack(const uint64_t* records, int64_t part_cap, const uint32_t* cursors, const int64_t* offsets,
                     int64_t n_parts, uint64_t* out, cfk_stream_t stream) {
  if (part_cap < 1 || n_parts < 0) return fail(CFK_ERR_INVALID, "cfk_records_pack: bad sizes");
  if (n_parts == 0) return CFK_OK;
  const int64_t grid = n_parts < 148 * 16 ? n_parts : 148 * 16;
  records_pack_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(records, part_cap, cursors, offsets, n_parts, out);
  CFK_CHECK_LAUNCH("records_pack_kernel", 1);
  return CFK_OK;
}

int cfk_docfreq_count_parts(const uint64_t* records, int64_t part_cap, const uint32_t* cursors, const int64_t* offsets,
                            int64_t n_parts, int32_t n_src, int64_t src_stride, int32_t group, int k, uint32_t lo, uint32_t hi,
                            uint32_t max_nonuniq, uint64_t* rare_keys, uint32_t* rare_nreads, uint32_t* rare_nmulti,
                            int64_t max_rare, uint64_t* dense, int64_t max_dense, int64_t* counters, int32_t n_blocks,
                            cfk_stream_t stream) {
  if (k < 1 || k > 31) return fail(CFK_ERR_INVALID, "cfk_docfreq_count_parts: k must be in [1, 31]");
  if (part_cap < 1 || n_parts < 0 || n_src < 1 || group < 1 || n_blocks < 1 || max_rare < 0 || max_dense < 0)
    return fail(CFK_ERR_INVALID, "cfk_docfreq_count_parts: bad sizes");
  if (n_parts == 0) return CFK_OK;
  static unsigned long long attr_done = 0;
  {
    cudaError_t e = ensure_dynamic_smem(docfreq_count_kernel, CN_SMEM_BYTES, &attr_done);
    if (e != cudaSuccess) return fail(CFK_ERR_CUDA, "cfk_docfreq_count_parts: cudaFuncSetAttribute", e);
  }
  CountArgs A;
  A.records = records;
  A.cursors = cursors;
  A.offsets = offsets;
  A.part_cap = part_cap;
  A.n_parts = n_parts;
  A.src_stride = src_stride;
  A.n_src = n_src;
  A.group = group;
  A.k = k;
  A.lo = lo;
  A.hi = hi;
  A.max_nonuniq = max_nonuniq;
  A.rare_keys = rare_keys;
  A.rare_nreads = rare_nreads;
  A.rare_nmulti = rare_nmulti;
  A.max_rare = max_rare;
  A.dense = reinterpret_cast<uint4*>(dense);
  A.max_dense = max_dense;
  A.counters = counters;
  const int64_t n_units = (n_parts + group - 1) / group;
  const int64_t want = (int64_t)n_blocks * (1024 / CN_THREADS);
  const int64_t grid = n_units < want ? n_units : want;
  docfreq_count_kernel<<<(unsigned)grid, CN_THREADS, CN_SMEM_BYTES, (cudaStream_t)stream>>>(A);
  CFK_CHECK_LAUNCH("docfreq_count_kernel", 1);
  return CFK_OK;
}

}  // extern "C"
