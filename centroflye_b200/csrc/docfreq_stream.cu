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

// ---- the shared-memory set: 4-slot buckets (one 16-byte load), linear probing over buckets --------------------
// Slot (32 bit): bits [id_bits-1:0] = id of the entry's owner + 1 (0 = empty slot, all ones = dead entry), the bits
// above up to bit 30 = fingerprint of the key, bit 31 = a flag of the user.  `fresh` = id + 1 | fingerprint.
// Slots never return to empty and a writer takes the FIRST empty slot it sees, so a probe chain has no holes and
// "first entry holding the key, walking from home" is the same slot for every walker: the canonical entry.
//
// CLAIM  (all keys, plain stores): walk the chain; stop at an entry holding the key, else store `fresh` into the
//        first empty slot.  Racing claims of one slot overwrite each other: the loser's key is simply not in the
//        set yet; racing claims of one KEY may leave it twice.
// VERIFY (after a barrier): walk again.  Canonical entry is mine -> owner.  Somebody else's -> on_match() (once),
//        then keep walking: an own entry further down is a stale second entry of the key and is marked dead.
//        An empty slot before any match: the claim was lost -> claim again with atomicCAS (rare).
//        A key that stored nothing in CLAIM (`claimed` false) cannot have a stale entry and stops at the match.
// Return: 1 done (claim: the key was there), 3 done (claim: stored), 0 go on with the next bucket, 2 look at this
// bucket again (a CAS lost against a newcomer).
template <class Eq>
__device__ __forceinline__ int set_claim_step(const uint4 v4, uint32_t* bucket, uint32_t fresh, uint32_t fp_mask, Eq eq) {
  const uint32_t mine = fresh & fp_mask;
  unsigned e = 0, m = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const uint32_t x = pick4(v4, s);
    if (x == 0) e |= 1u << s;
    else if (((x ^ mine) & fp_mask) == 0) m |= 1u << s;
  }
  while (m) {
    const int s = __ffs(m) - 1;
    m &= m - 1;
    if (eq(pick4(v4, s))) return 1;
  }
  if (e) {
    bucket[__ffs(e) - 1] = fresh;
    return 3;
  }
  return 0;
}

template <class Eq, class OnMatch>
__device__ __forceinline__ int set_verify_step(const uint4 v4, uint32_t* bucket, uint32_t fresh, uint32_t id_mask,
                                               uint32_t fp_mask, bool claimed, bool& matched, int& own_slot, Eq eq,
                                               OnMatch on_match) {
  const uint32_t mine = fresh & fp_mask;
  unsigned o = 0, e = 0, m = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const uint32_t x = pick4(v4, s);
    if (x == 0) e |= 1u << s;
    else if ((x & 0x7FFFFFFFu) == fresh) o |= 1u << s;
    else if (((x ^ mine) & fp_mask) == 0 && (x & id_mask) != id_mask) m |= 1u << s;
  }
  if (!matched) {
    const unsigned stop = o | e;
    unsigned mm = stop ? (m & ((stop & (0u - stop)) - 1u)) : m;  // candidates in front of the own entry / the chain's end
    while (mm) {
      const int s = __ffs(mm) - 1;
      mm &= mm - 1;
      const uint32_t x = pick4(v4, s);
      if (eq(x)) {
        matched = true;
        on_match(s, x);
        if (!claimed) return 1;
        break;
      }
    }
  }
  if (o) {
    const int s = __ffs(o) - 1;
    if (matched) bucket[s] = pick4(v4, s) | id_mask;  // stale second entry of this key: dead
    else own_slot = s;
    return 1;
  }
  if (e) {
    if (matched) return 1;
    const int s = __ffs(e) - 1;
    const uint32_t old = atomicCAS(bucket + s, 0u, fresh);  // the lost claim, made good
    if (old == 0) {
      own_slot = s;
      return 1;
    }
    return 2;
  }
  return 0;
}

// ================================================================================================================
// phase 1
// ================================================================================================================
#ifndef CFK_DE_FILL_PCT
#define CFK_DE_FILL_PCT 72   /* planned load of the per-read set, percent (4-slot buckets) */
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
constexpr int DE_DATA_WORDS = 57344;        // dynamic shared memory of the block (224 KB): [ read words | set ]
constexpr int DE_MIN_SET = 16384;
constexpr uint32_t DE_MULTI = 0x80000000u;
#ifndef CFK_DE_MODE
#define CFK_DE_MODE 1   /* 1: one walk per k-mer, empty slots claimed with atomicCAS, probes issued from a warp queue;
                           0: claim with plain stores, barrier, verify (no shared-memory atomics) */
#endif
constexpr int DE_Q = 64;                                              // queue entries per warp (mode 1)
constexpr int DE_Q_WORDS = CFK_DE_MODE == 1 ? DE_WARPS * DE_Q * 2 : 0;  // the queues sit behind the set
static_assert(DE_DATA_WORDS > 2 * DE_MIN_SET, "shared memory budget");

struct DeGeometry {
  uint32_t n_words;    // words staged (0: the read stays in global memory)
  uint32_t bm_words;   // "stored something in the claim phase" bits, one per k-mer start
  uint32_t n_buckets;  // 4-slot buckets of the set
  uint32_t fill;       // k-mers planned per pass
};

__host__ __device__ __forceinline__ DeGeometry de_geometry(int64_t len) {
  DeGeometry g;
  const int64_t nw = (((len + 15) >> 4) + 3 + 3) & ~(int64_t)3;  // + the 3-word extraction window, multiple of 4
  const int64_t bw = (((len + DE_CHUNK - 1) / DE_CHUNK) * DE_PER + 3) & ~(int64_t)3;
  g.bm_words = (CFK_DE_MODE == 0 && bw <= DE_DATA_WORDS / 4) ? (uint32_t)bw : 0u;  // 0: no bitmap (every key counts as "stored")
  g.n_words = (nw + g.bm_words <= DE_DATA_WORDS - DE_Q_WORDS - DE_MIN_SET) ? (uint32_t)nw : 0u;
  g.n_buckets = ((uint32_t)DE_DATA_WORDS - DE_Q_WORDS - g.n_words - g.bm_words) >> 2;
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

// The one hash of phase 1.  g orders everything: pass = floor(g * n_pass / 2^32), set bucket = the fraction of that
// product scaled to the set, partition = floor(g * n_parts / 2^32); the fingerprint takes g's low bits.
__device__ __forceinline__ uint32_t de_hash(uint64_t kmer) {
  uint32_t g = (uint32_t)(kmer ^ (kmer >> 32)) * 0x85EBCA6Bu;
  g ^= g >> 13;
  g *= 0xC2B2AE35u;
  g ^= g >> 16;
  return g;
}

template <bool IN_SMEM, bool VERIFY>
__device__ __forceinline__ void de_phase(const uint32_t* words, uint32_t* bm, uint32_t* set, uint32_t nb, int64_t nk, int k,
                                         uint32_t pass, uint32_t n_pass, int64_t* counters) {
  const uint64_t mask = (1ull << (2 * k)) - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pos_bits = 64 - __clzll((unsigned long long)(nk + 1));  // pos + 1 <= nk < 2^pos_bits - 1
  const uint32_t pos_mask = (1u << pos_bits) - 1u;                  // nk < 2^30 (checked on the host)
  const uint32_t fp_mask = 0x7FFFFFFFu & ~pos_mask;
  for (int64_t chunk0 = (int64_t)warp * DE_CHUNK; chunk0 < nk; chunk0 += (int64_t)DE_WARPS * DE_CHUNK) {
    const int64_t base = chunk0 + lane * DE_PER;
    const uint32_t p0 = (uint32_t)base;  // multiple of 4: offset 0, 4, 8 or 12 inside its word
    const int npos = base < nk ? (int)min((int64_t)DE_PER, nk - base) : 0;
    uint64_t km[DE_PER];
    uint32_t fresh[DE_PER], bkt[DE_PER];
    unsigned pend = 0;
    if (npos > 0) {
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
        km[j] = kmer;
        const uint32_t g = de_hash(kmer);
        const uint64_t t = (uint64_t)g * n_pass;
        bkt[j] = __umulhi((uint32_t)t, nb);
        fresh[j] = (p0 + (uint32_t)j + 1u) | ((g << pos_bits) & fp_mask);
        if (j < npos && (uint32_t)(t >> 32) == pass) pend |= 1u << j;
      }
    }
    unsigned matched = 0, stored = VERIFY ? 0xFFFFFFFFu : 0u;
    uint32_t* bm_chunk = bm + (chunk0 / DE_CHUNK) * DE_PER;  // this warp step's DE_PER words: bit = lane
    if (VERIFY && bm != nullptr) {
      stored = 0;
#pragma unroll
      for (int j = 0; j < DE_PER; ++j) stored |= ((bm_chunk[j] >> lane) & 1u) << j;
    }
    uint32_t rounds = 0;
    while (__any_sync(FULL, pend != 0)) {
#pragma unroll
      for (int j = 0; j < DE_PER; ++j) {
        if (!((pend >> j) & 1u)) continue;
        const uint4 v4j = *reinterpret_cast<const uint4*>(set + 4 * bkt[j]);
        const uint64_t kmer = km[j];
        auto eq = [&](uint32_t x) { return de_kmer_at<IN_SMEM>(words, (x & pos_mask) - 1u, k) == kmer; };
        uint32_t* bucket = set + 4 * bkt[j];
        int r;
        if (!VERIFY) {
          r = set_claim_step(v4j, bucket, fresh[j], fp_mask, eq);
          if (r == 3) {
            stored |= 1u << j;
            r = 1;
          }
        } else {
          bool mt = (matched >> j) & 1u;
          int own = -1;
          r = set_verify_step(v4j, bucket, fresh[j], pos_mask, fp_mask, (stored >> j) & 1u, mt, own, eq, [&](int s, uint32_t x) {
            if (!(x & DE_MULTI)) bucket[s] = x | DE_MULTI;  // the k-mer occurs again in this read (idempotent)
          });
          if (mt) matched |= 1u << j;
        }
        if (r == 1) pend &= ~(1u << j);
        else if (r == 0) bkt[j] = (bkt[j] + 1 == nb) ? 0u : bkt[j] + 1;
      }
      if (++rounds > 2 * nb + 64) {  // cannot happen: a pass is planned for <= CFK_DE_FILL_PCT % load
        counters[1] = 1;
        break;
      }
    }
    if (!VERIFY && bm != nullptr) {
#pragma unroll
      for (int j = 0; j < DE_PER; ++j) {
        const unsigned w = __ballot_sync(FULL, (stored >> j) & 1u);
        if (lane == 0) bm_chunk[j] = w;
      }
    }
  }
}

// ---- mode 1: one walk per k-mer, probes issued from a warp-private queue -----------------------------------------
// A k-mer start that belongs to this pass becomes a queue item { position | flags, bucket to look at next }.  Whenever
// 32 items wait, the warp takes them, one per lane: re-extracts the k-mer, looks at one bucket, and either finishes
// (the k-mer is there: set "again"; or an empty slot was claimed with atomicCAS) or puts the item back with its next
// bucket.  Every probe runs on full warps, whatever share of the positions the pass selects and however long single
// chains get.
template <bool IN_SMEM>
__device__ __forceinline__ int de_drain(const uint32_t* words, uint32_t* set, uint32_t nb, int k, uint32_t n_pass, int pos_bits,
                                        uint2* q, int qn, int64_t* counters) {
  const int lane = threadIdx.x & 31;
  const uint32_t pos_mask = (1u << pos_bits) - 1u;
  const uint32_t fp_mask = 0x7FFFFFFFu & ~pos_mask;
  const int n = min(32, qn), first = qn - n;
  bool more = false;
  uint2 it = make_uint2(0, 0);
  if (lane < n) {
    it = q[first + lane];
    const uint32_t pos = it.x;
    const uint64_t kmer = de_kmer_at<IN_SMEM>(words, pos, k);
    const uint32_t g = de_hash(kmer);
    const uint32_t fresh = (pos + 1u) | ((g << pos_bits) & fp_mask);
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
    while (m) {
      const int s = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t x = pick4(v4, s);
      if (de_kmer_at<IN_SMEM>(words, (x & pos_mask) - 1u, k) == kmer) {
        if (!(x & DE_MULTI)) bucket[s] = x | DE_MULTI;  // the k-mer occurs again in this read (idempotent)
        done = true;
        break;
      }
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
  uint2* q = queues + warp * DE_Q;
  const unsigned lt = (1u << lane) - 1u;
  int qn = 0;
  for (int64_t chunk0 = (int64_t)warp * DE_CHUNK; chunk0 < nk; chunk0 += (int64_t)DE_WARPS * DE_CHUNK) {
    const int64_t base = chunk0 + lane * DE_PER;
    const uint32_t p0 = (uint32_t)base;  // multiple of DE_PER
    const int npos = base < nk ? (int)min((int64_t)DE_PER, nk - base) : 0;
    uint32_t home[DE_PER];
    unsigned act = 0;
    if (npos > 0) {
      // 48-base window from the word of p0: offset + DE_PER - 1 + k - 1 <= 14 + 3 + 30 < 48
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
        const uint64_t t = (uint64_t)de_hash(kmer) * n_pass;
        home[j] = __umulhi((uint32_t)t, nb);
        if (j < npos && (uint32_t)(t >> 32) == pass) act |= 1u << j;
      }
    }
#pragma unroll
    for (int j = 0; j < DE_PER; ++j) {
      const bool a = (act >> j) & 1u;
      const unsigned m = __ballot_sync(FULL, a);
      if (m == 0) continue;
      if (a) q[qn + __popc(m & lt)] = make_uint2(p0 + (uint32_t)j, home[j]);
      qn += __popc(m);
      __syncwarp();
      while (qn >= 32) qn = de_drain<IN_SMEM>(words, set, nb, k, n_pass, pos_bits, q, qn, counters);
    }
  }
  uint32_t rounds = 0;
  while (qn > 0) {
    qn = de_drain<IN_SMEM>(words, set, nb, k, n_pass, pos_bits, q, qn, counters);
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
                                             uint64_t* __restrict__ records, int64_t part_cap, uint32_t n_parts,
                                             uint32_t* cursors, int64_t* counters) {
  const int pos_bits = 64 - __clzll((unsigned long long)(nk + 1));
  const uint32_t pos_mask = (1u << pos_bits) - 1u;
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
      if (v != 0 && (v & pos_mask) != pos_mask) {
        const uint64_t kmer = de_kmer_at<IN_SMEM>(words, (v & pos_mask) - 1u, k);
        part[s] = __umulhi(de_hash(kmer), n_parts);
        rec[s] = kmer | ((uint64_t)(v >> 31) << 63);
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
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (!((live >> s) & 1u)) continue;
      if ((int64_t)idx[s] < part_cap) records[(int64_t)part[s] * part_cap + idx[s]] = rec[s];
      else over = true;
    }
    if (over) counters[0] = 1;  // partition buffer full: the host falls back (never silent)
  }
}

__global__ void __launch_bounds__(DE_THREADS, 1)
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
    uint32_t* bm = g.bm_words ? data + g.n_words : nullptr;
    (void)bm;
    uint32_t* set = data + g.n_words + g.bm_words;
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
#if CFK_DE_MODE == 1
    uint2* queues = reinterpret_cast<uint2*>(data + (DE_DATA_WORDS - DE_Q_WORDS));
    if (g.n_words) {
      de_insert_q<true>(data, set, nb, nk, k, pass, n_pass, queues, counters);
      __syncthreads();
      de_scan_emit<true>(data, set, nb, nk, k, records, part_cap, n_parts, cursors, counters);
    } else {
      de_insert_q<false>(gwords, set, nb, nk, k, pass, n_pass, queues, counters);
      __syncthreads();
      de_scan_emit<false>(gwords, set, nb, nk, k, records, part_cap, n_parts, cursors, counters);
    }
#else
    if (g.n_words) {
      de_phase<true, false>(data, bm, set, nb, nk, k, pass, n_pass, counters);
      __syncthreads();
      de_phase<true, true>(data, bm, set, nb, nk, k, pass, n_pass, counters);
      __syncthreads();
      de_scan_emit<true>(data, set, nb, nk, k, records, part_cap, n_parts, cursors, counters);
    } else {
      de_phase<false, false>(gwords, bm, set, nb, nk, k, pass, n_pass, counters);
      __syncthreads();
      de_phase<false, true>(gwords, bm, set, nb, nk, k, pass, n_pass, counters);
      __syncthreads();
      de_scan_emit<false>(gwords, set, nb, nk, k, records, part_cap, n_parts, cursors, counters);
    }
#endif
  }
}

// ================================================================================================================
// phase 2
// ================================================================================================================
// One block per partition (tickets).  Shared memory: dk[] = the partition's distinct records so far followed by the
// chunk being added, the set (slot id = index into dk + 1) and one 32-bit word per slot for the k-mers seen in more
// than one read: low half = further reads, high half = how many of those held the k-mer more than once.  The
// owner's own read is implicit (n_reads = 1 + low half, n_multi = bit 63 of the owner's record + high half), so the
// ~97 % of k-mers that occur in one read never touch a counter.  Records arrive in chunks of CN_CHUNK; after a
// chunk its owners are compacted to the front of the chunk area (they are the new distinct keys), so a partition
// made long by a k-mer present in every read still needs room for its DISTINCT keys only.
constexpr int CN_THREADS = 1024;
constexpr int CN_PER = 4;
constexpr int CN_CHUNK = CN_THREADS * CN_PER;  // records added per round
constexpr int CN_DCAP = CFK_DOCFREQ_PART_DISTINCT;  // distinct k-mers a partition may hold
constexpr int CN_NB = 2048;                    // 4-slot buckets
constexpr uint32_t CN_ID_MASK = 0x3FFFu;       // 14 bits: index into dk + 1
constexpr uint32_t CN_FP_MASK = 0x7FFFC000u;
constexpr uint64_t CN_KEY = 0x3FFFFFFFFFFFFFFFull;
constexpr int CN_SMEM_BYTES = (CN_DCAP + CN_CHUNK) * 8 + CN_NB * 16 * 2;
static_assert(CN_DCAP + CN_CHUNK + 1 < (int)CN_ID_MASK, "slot id field");
static_assert(CN_DCAP <= CN_NB * 4 * 3 / 4, "the set stays below 75 % load");

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
    const int w = s_warp[lane];
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

struct CountOut {
  uint32_t lo, hi, max_nonuniq;
  uint64_t* rare_keys;
  uint32_t* rare_nreads;
  uint32_t* rare_nmulti;
  int64_t max_rare;
  uint4* dense;
  int64_t max_dense;
};

__global__ void __launch_bounds__(CN_THREADS, 1)
docfreq_count_kernel(const uint64_t* __restrict__ records, int64_t part_cap, const uint32_t* __restrict__ cursors,
                     int64_t n_parts, int32_t n_src, int64_t src_stride, CountOut out, int64_t* counters) {
  extern __shared__ __align__(16) uint32_t cn_smem[];
  uint64_t* dk = reinterpret_cast<uint64_t*>(cn_smem);
  uint32_t* set = cn_smem + 2 * (CN_DCAP + CN_CHUNK);
  uint32_t* cnt = set + 4 * CN_NB;
  __shared__ int s_warp[33];
  __shared__ long long s_ticket, s_base;
  __shared__ int s_abort;
  const int lane = threadIdx.x & 31;
  for (;;) {
    if (threadIdx.x == 0) {
      s_ticket = (long long)atomicAdd((unsigned long long*)(counters + 3), 1ull);
      s_abort = 0;
    }
    for (uint32_t i = threadIdx.x; i < 2u * CN_NB; i += CN_THREADS)  // set and counters are adjacent
      reinterpret_cast<uint4*>(set)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const int64_t p = s_ticket;
    if (p >= n_parts) break;
    int nd = 0;  // distinct records at the front of dk
    bool aborted = false;  // block-uniform: s_abort is only read right behind a barrier
    for (int32_t src = 0; src < n_src && !aborted; ++src) {
      const int64_t n = min((int64_t)__ldg(cursors + (int64_t)src * n_parts + p), part_cap);
      const uint64_t* base = records + (int64_t)src * src_stride + p * part_cap;
      for (int64_t c0 = 0; c0 < n; c0 += CN_CHUNK) {
        const int c = (int)min((int64_t)CN_CHUNK, n - c0);
        const bool last = (c0 + CN_CHUNK >= n) && (src + 1 == n_src);
        uint64_t rec[CN_PER];
        uint32_t fresh[CN_PER], bkt[CN_PER];
        unsigned pend0 = 0, stored = 0;
#pragma unroll
        for (int j = 0; j < CN_PER; ++j) {
          const int i = threadIdx.x + j * CN_THREADS;
          rec[j] = 0;
          if (i < c) {
            rec[j] = __ldcs(base + c0 + i);
            dk[nd + i] = rec[j];
            const uint32_t x = cn_hash(rec[j] & CN_KEY);
            bkt[j] = __umulhi(x, (uint32_t)CN_NB);
            fresh[j] = (uint32_t)(nd + i + 1) | ((x << 14) & CN_FP_MASK);
            pend0 |= 1u << j;
          }
        }
        __syncthreads();
        // ---- claim
        {
          unsigned pend = pend0;
          uint32_t b[CN_PER];
#pragma unroll
          for (int j = 0; j < CN_PER; ++j) b[j] = bkt[j];
          uint32_t rounds = 0;
          while (__any_sync(FULL, pend != 0)) {
#pragma unroll
            for (int j = 0; j < CN_PER; ++j) {
              if (!((pend >> j) & 1u)) continue;
              const uint4 v4j = *reinterpret_cast<const uint4*>(set + 4 * b[j]);
              const uint64_t key = rec[j] & CN_KEY;
              auto eq = [&](uint32_t x) { return (dk[(x & CN_ID_MASK) - 1u] & CN_KEY) == key; };
              const int r = set_claim_step(v4j, set + 4 * b[j], fresh[j], CN_FP_MASK, eq);
              if (r == 3) stored |= 1u << j;
              if (r != 0) pend &= ~(1u << j);
              else b[j] = (b[j] + 1 == CN_NB) ? 0u : b[j] + 1;
            }
            if (++rounds > 2 * CN_NB + 64) {  // the set is full: more distinct k-mers than planned
              s_abort = 1;
              break;
            }
          }
        }
        __syncthreads();
        aborted = s_abort != 0;
        // ---- verify: owners learn their slot, the others add their read to the canonical entry's counter
        int own[CN_PER];
        unsigned owner = 0;
        if (!aborted) {
          unsigned pend = pend0, matched = 0;
          uint32_t b[CN_PER];
#pragma unroll
          for (int j = 0; j < CN_PER; ++j) {
            b[j] = bkt[j];
            own[j] = -1;
          }
          uint32_t rounds = 0;
          while (__any_sync(FULL, pend != 0)) {
#pragma unroll
            for (int j = 0; j < CN_PER; ++j) {
              if (!((pend >> j) & 1u)) continue;
              const uint4 v4j = *reinterpret_cast<const uint4*>(set + 4 * b[j]);
              const uint64_t key = rec[j] & CN_KEY;
              const uint32_t inc = 1u + (uint32_t)(rec[j] >> 63 << 16);
              auto eq = [&](uint32_t x) { return (dk[(x & CN_ID_MASK) - 1u] & CN_KEY) == key; };
              bool mt = (matched >> j) & 1u;
              int os = -1;
              const uint32_t bj = b[j];
              const int r = set_verify_step(v4j, set + 4 * bj, fresh[j], CN_ID_MASK, CN_FP_MASK, (stored >> j) & 1u, mt, os, eq,
                                            [&](int s, uint32_t) {
                                              const uint32_t old = atomicAdd(cnt + 4 * bj + s, inc);
                                              if ((old & 0xFFFFu) == 0xFFFFu) counters[0] = 2;  // 65536 further reads: 16-bit halves exhausted
                                            });
              if (mt) matched |= 1u << j;
              if (os >= 0) {
                own[j] = (int)(4 * bj) + os;
                owner |= 1u << j;
              }
              if (r == 1) pend &= ~(1u << j);
              else if (r == 0) b[j] = (b[j] + 1 == CN_NB) ? 0u : b[j] + 1;
            }
            if (++rounds > 4 * CN_NB + 64) {
              s_abort = 1;
              break;
            }
          }
        }
        __syncthreads();
        aborted = s_abort != 0;
        if (aborted) break;
        if (!last) {
          // ---- the chunk's owners move to the front of the chunk area: they are the new distinct keys
          int total = 0;
          int rank = cn_block_scan(__popc(owner), &total, s_warp);
#pragma unroll
          for (int j = 0; j < CN_PER; ++j) {
            if (!((owner >> j) & 1u)) continue;
            const int at = nd + rank++;
            if (at < CN_DCAP + CN_CHUNK) {
              dk[at] = rec[j];
              set[own[j]] = (uint32_t)(at + 1) | (fresh[j] & CN_FP_MASK);
            }
          }
          nd += total;
          __syncthreads();
          if (nd > CN_DCAP) {  // block-uniform
            aborted = true;
            break;
          }
        }
      }
    }
    if (aborted) {
      if (threadIdx.x == 0) counters[0] = 1;  // more distinct k-mers in one partition than planned: the host falls back
      __syncthreads();
      continue;
    }
    // ---- output: every live slot is one distinct k-mer with its final counts
    constexpr int OUT_PER = 4 * CN_NB / CN_THREADS;
    uint64_t okey[OUT_PER];
    uint32_t onr[OUT_PER], onm[OUT_PER];
    int n_live = 0;
#pragma unroll
    for (int j = 0; j < OUT_PER; ++j) {
      const uint32_t s = threadIdx.x + j * CN_THREADS;
      const uint32_t v = set[s];
      okey[j] = EMPTY;
      if (v != 0 && (v & CN_ID_MASK) != CN_ID_MASK) {
        const uint64_t r = dk[(v & CN_ID_MASK) - 1u];
        const uint32_t cw = cnt[s];
        okey[j] = r & CN_KEY;
        onr[j] = 1u + (cw & 0xFFFFu);
        onm[j] = (uint32_t)(r >> 63) + (cw >> 16);
        ++n_live;
      }
    }
    if (out.rare_keys != nullptr) {
#pragma unroll
      for (int j = 0; j < OUT_PER; ++j) {
        const bool take = okey[j] != EMPTY && onm[j] <= out.max_nonuniq && onr[j] >= out.lo && onr[j] <= out.hi;
        const unsigned m = __ballot_sync(FULL, take);
        if (m == 0) continue;
        long long at = 0;
        if (lane == __ffs(m) - 1) at = (long long)atomicAdd((unsigned long long*)(counters + 4), (unsigned long long)__popc(m));
        at = __shfl_sync(FULL, at, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u));
        if (take && at < out.max_rare) {
          out.rare_keys[at] = okey[j];
          if (out.rare_nreads != nullptr) out.rare_nreads[at] = onr[j];
          if (out.rare_nmulti != nullptr) out.rare_nmulti[at] = onm[j];
        }
      }
    }
    if (out.dense != nullptr) {
      int total = 0;
      int rank = cn_block_scan(n_live, &total, s_warp);
      if (threadIdx.x == 0) s_base = (long long)atomicAdd((unsigned long long*)(counters + 5), (unsigned long long)total);
      __syncthreads();
      const long long at0 = s_base;
#pragma unroll
      for (int j = 0; j < OUT_PER; ++j) {
        if (okey[j] == EMPTY) continue;
        const long long at = at0 + rank++;
        if (at < out.max_dense)
          out.dense[at] = make_uint4((uint32_t)okey[j], (uint32_t)(okey[j] >> 32), onr[j], onm[j]);
      }
    }
    __syncthreads();  // the table is cleared at the top of the loop
  }
}

}  // namespace

extern "C" {

int cfk_docfreq_part_target(void) { return CN_DCAP * 3 / 4; }

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
  docfreq_emit_kernel<<<(unsigned)n_blocks, DE_THREADS, smem, (cudaStream_t)stream>>>(
      packed, read_off, read_len, order, item_ptr, n_reads, k, records, part_cap, (uint32_t)n_parts, cursors, counters);
  CFK_CHECK_LAUNCH("docfreq_emit_kernel", 1);
  return CFK_OK;
}

int cfk_docfreq_count_parts(const uint64_t* records, int64_t part_cap, const uint32_t* cursors, int64_t n_parts,
                            int32_t n_src, int64_t src_stride, uint32_t lo, uint32_t hi, uint32_t max_nonuniq,
                            uint64_t* rare_keys, uint32_t* rare_nreads, uint32_t* rare_nmulti, int64_t max_rare,
                            uint64_t* dense, int64_t max_dense, int64_t* counters, int32_t n_blocks,
                            cfk_stream_t stream) {
  if (part_cap < 1 || n_parts < 0 || n_src < 1 || n_blocks < 1 || max_rare < 0 || max_dense < 0)
    return fail(CFK_ERR_INVALID, "cfk_docfreq_count_parts: bad sizes");
  if (n_parts == 0) return CFK_OK;
  static unsigned long long attr_done = 0;
  {
    cudaError_t e = ensure_dynamic_smem(docfreq_count_kernel, CN_SMEM_BYTES, &attr_done);
    if (e != cudaSuccess) return fail(CFK_ERR_CUDA, "cfk_docfreq_count_parts: cudaFuncSetAttribute", e);
  }
  CountOut out;
  out.lo = lo;
  out.hi = hi;
  out.max_nonuniq = max_nonuniq;
  out.rare_keys = rare_keys;
  out.rare_nreads = rare_nreads;
  out.rare_nmulti = rare_nmulti;
  out.max_rare = max_rare;
  out.dense = reinterpret_cast<uint4*>(dense);
  out.max_dense = max_dense;
  const int64_t grid = n_parts < n_blocks ? n_parts : n_blocks;
  docfreq_count_kernel<<<(unsigned)grid, CN_THREADS, CN_SMEM_BYTES, (cudaStream_t)stream>>>(
      records, part_cap, cursors, n_parts, n_src, src_stride, out, counters);
  CFK_CHECK_LAUNCH("docfreq_count_kernel", 1);
  return CFK_OK;
}

}  // extern "C"
