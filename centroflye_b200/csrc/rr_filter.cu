// rr_filter.cu — the read recruitment pre-filter (second half).
//
// Replaces the two edlibAlign calls of scripts/read_recruitment/rr.cpp:73-90: a read is kept when the unit, or its
// reverse complement, occurs somewhere in the read within `threshold` edits (edlib's HW mode: the whole unit against
// any infix of the read, EDLIB_TASK_DISTANCE with k = threshold; editDistance != -1 <=> best infix distance <= k).
//
// One thread per (read, strand): Myers' bit-vector algorithm in Hyyro's block form, the unit as the pattern cut into
// NW 64-bit words whose vertical deltas (Pv, Mv) stay in registers for the whole read; the match masks Peq[symbol][word]
// of both strands sit in shared memory.  Per text character the thread walks the NW blocks once, carrying the
// horizontal delta; the bottom-row score starts at |unit| (free start in the text: the top row is all zeros) and its
// running minimum is the infix distance.  Bits above |unit| in the last word never feed back into lower bits (carries
// only travel upwards), so the bottom row's delta is read at bit (|unit| - 1) mod 64 and no wildcard padding is needed.
// A strand stops as soon as its minimum is within the threshold (the caller only wants to know whether it is) unless
// exact distances are asked for.  Characters are compared as bytes, like edlib: a read symbol that the unit does not
// contain matches nothing.
#include "cfk_common.cuh"

namespace {

using namespace cfk;

constexpr int RR_THREADS = 128;
constexpr int RR_MAX_SYMBOLS = 8;  // distinct bytes of the unit (ACGT in practice) + slot 0 = "not in the unit"

struct RrArgs {
  const uint8_t* text;        // all reads, every read starts on a 16-byte boundary
  const int64_t* read_off;    // byte offset of read r
  const int64_t* read_len;
  const int32_t* order;       // items are dealt longest read first
  int64_t n_reads;
  const uint64_t* peq;        // [2 strands][RR_MAX_SYMBOLS][NW]
  const uint8_t* sym_of;      // [256] byte -> symbol slot (0 = not in the unit)
  int32_t m;                  // unit length
  int32_t threshold;          // < 0: no limit (edlib's k = -1)
  int32_t exact;              // != 0: never stop early, out_dist holds the exact infix distance
  int32_t* out_dist;          // [2 * n_reads] (read, strand), or NULL
  uint8_t* out_keep;          // [n_reads]
};

template <int NW>
__global__ void __launch_bounds__(RR_THREADS) rr_filter_kernel(const RrArgs A) {
  __shared__ uint64_t s_peq[2 * RR_MAX_SYMBOLS * NW];
  __shared__ uint8_t s_sym[256];
  for (int i = threadIdx.x; i < 2 * RR_MAX_SYMBOLS * NW; i += RR_THREADS) s_peq[i] = A.peq[i];
  for (int i = threadIdx.x; i < 256; i += RR_THREADS) s_sym[i] = A.sym_of[i];
  __syncthreads();
  const int64_t item = (int64_t)blockIdx.x * RR_THREADS + threadIdx.x;
  if (item >= 2 * A.n_reads) return;
  const int64_t r = A.order[item >> 1];
  const int strand = (int)(item & 1);
  const uint64_t* peq = s_peq + strand * RR_MAX_SYMBOLS * NW;
  const int64_t n = A.read_len[r];
  const uint4* text = reinterpret_cast<const uint4*>(A.text + A.read_off[r]);
  const int last_bit = (A.m - 1) & 63;
  uint64_t Pv[NW], Mv[NW];
#pragma unroll
  for (int j = 0; j < NW; ++j) {
    Pv[j] = ~0ull;
    Mv[j] = 0ull;
  }
  int score = A.m, best = A.m;
  const bool limited = A.threshold >= 0 && !A.exact;
  for (int64_t t0 = 0; t0 < n; t0 += 16) {
    const uint4 chunk = __ldg(text + (t0 >> 4));
    const int cnt = (int)min((int64_t)16, n - t0);
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
      uint32_t word = wi == 0 ? chunk.x : wi == 1 ? chunk.y : wi == 2 ? chunk.z : chunk.w;
#pragma unroll 1
      for (int c = 4 * wi; c < min(4 * wi + 4, cnt); ++c, word >>= 8) {
        const uint64_t* eqrow = peq + (int)s_sym[word & 0xFFu] * NW;
        int hin = 0;  // free start in the text: the top row is all zeros
#pragma unroll
        for (int j = 0; j < NW; ++j) {
          uint64_t Eq = eqrow[j];
          const uint64_t pv = Pv[j], mv = Mv[j];
          const uint64_t Xv = Eq | mv;
          if (hin < 0) Eq |= 1ull;
          const uint64_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
          uint64_t Ph = mv | ~(Xh | pv);
          uint64_t Mh = pv & Xh;
          const int bit = (j == NW - 1) ? last_bit : 63;
          const int hout = (int)((Ph >> bit) & 1ull) - (int)((Mh >> bit) & 1ull);
          Ph <<= 1;
          Mh <<= 1;
          if (hin < 0) Mh |= 1ull;
          else if (hin > 0) Ph |= 1ull;
          Pv[j] = Mh | ~(Xv | Ph);
          Mv[j] = Ph & Xv;
          hin = hout;
        }
        score += hin;  // the last block's delta at the unit's last row
        best = min(best, score);
      }
    }
    if (limited && best <= A.threshold) break;
  }
  if (A.out_dist != nullptr) A.out_dist[2 * r + strand] = best;
  if (A.threshold < 0 || best <= A.threshold) A.out_keep[r] = 1;
}

template <int NW>
int rr_launch(const RrArgs& A, cudaStream_t st) {
  const int64_t items = 2 * A.n_reads;
  rr_filter_kernel<NW><<<(unsigned)blocks_for(items, RR_THREADS), RR_THREADS, 0, st>>>(A);
  return NW;
}

}  // namespace

extern "C" {

int cfk_rr_max_symbols(void) { return RR_MAX_SYMBOLS; }

/* words per strand of the peq table for a unit of m bases, or -1 if the unit is too long */
int cfk_rr_words(int32_t m) {
  if (m < 1) return -1;
  const int need = (m + 63) / 64;
  const int sizes[] = {4, 9, 17, 33, 52};
  for (int s : sizes)
    if (need <= s) return s;
  return -1;
}

int cfk_rr_filter(const uint8_t* text, const int64_t* read_off, const int64_t* read_len, const int32_t* order, int64_t n_reads,
                  const uint64_t* peq, const uint8_t* sym_of, int32_t m, int32_t threshold, int32_t exact, int32_t* out_dist,
                  uint8_t* out_keep, cfk_stream_t stream) {
  const int nw = cfk_rr_words(m);
  if (nw < 0) return fail(CFK_ERR_INVALID, "cfk_rr_filter: the unit must have 1..3328 bases");
  if (n_reads < 0 || n_reads >= (1ll << 30)) return fail(CFK_ERR_INVALID, "cfk_rr_filter: bad sizes");
  if (((uintptr_t)text & 15u) != 0) return fail(CFK_ERR_INVALID, "cfk_rr_filter: text must be 16-byte aligned");
  if (n_reads == 0) return CFK_OK;
  RrArgs A{text, read_off, read_len, order, n_reads, peq, sym_of, m, threshold, exact, out_dist, out_keep};
  cudaStream_t st = (cudaStream_t)stream;
  switch (nw) {
    case 4: rr_launch<4>(A, st); break;
    case 9: rr_launch<9>(A, st); break;
    case 17: rr_launch<17>(A, st); break;
    case 33: rr_launch<33>(A, st); break;
    default: rr_launch<52>(A, st); break;
  }
  CFK_CHECK_LAUNCH("rr_filter_kernel", 1);
  return CFK_OK;
}

}  // extern "C"
