// ncrf_ingest.cpp — native NCRF report ingestion (host side of the recruitment path).
//
// One pass from the report text to the flat arrays the device consumes: record selection
// (scripts/ncrf_parser.py:61-118), strand flip (utils/bio.py:27-29), gap removal + 2-bit packing
// (what distance_based_kmer_recruitment.py:47-53 slides over) and unit segmentation
// (scripts/ncrf_parser.py:28-59, as the linear scan of centroflye_b200/ncrf_parser.py) — without
// building a Python object per record.  Results are bit-identical to
// ingest.batch_from_report(NCRF_Report(path)) + ingest.units_from_report(...), which the golden
// fixtures pin to the reference's regex parser (tests/test_ncrf_native.py).
//
// Plain host C++ (std::thread over records); linked into libcfk.so so the C ABI stays one library.
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/cfk.h"

namespace {

thread_local char g_ingest_err[512] = "";

inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }  // str.strip() / \s on ASCII
inline bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }

struct Span {
  const char* p = nullptr;
  size_t n = 0;
};

struct Record {
  Span id, r_al, m_al, motif;
  int64_t r_len = 0, r_al_len = 0, r_st = 0, r_en = 0, m_al_len = 0, score = 0;
  char strand = '+';
};

// \s+ then (\d+): returns false if either is missing or the number overflows int64
bool ws_number(const char*& q, const char* end, int64_t& out, bool need_ws = true) {
  const char* s = q;
  while (s < end && is_space((unsigned char)*s)) ++s;
  if (need_ws && s == q) return false;
  if (s >= end || !is_digit((unsigned char)*s)) return false;
  uint64_t v = 0;
  while (s < end && is_digit((unsigned char)*s)) {
    if (v > (uint64_t)INT64_MAX / 10) return false;
    v = v * 10 + (uint64_t)(*s - '0');
    if (v > (uint64_t)INT64_MAX) return false;
    ++s;
  }
  out = (int64_t)v;
  q = s;
  return true;
}

bool literal(const char*& q, const char* end, const char* lit) {
  const size_t n = strlen(lit);
  if ((size_t)(end - q) < n || memcmp(q, lit, n) != 0) return false;
  q += n;
  return true;
}

// \s+(.+)$
bool ws_rest(const char*& q, const char* end, Span& out) {
  const char* s = q;
  while (s < end && is_space((unsigned char)*s)) ++s;
  if (s == q || s >= end) return false;
  out.p = s;
  out.n = (size_t)(end - s);
  q = end;
  return true;
}

// ^([^ ]+)\s+(\d+)\s+(\d+)bp\s+(\d+)-(\d+)\s+(.+)$   (ncrf_parser.py:74), including the backtracking of group 1
bool parse_first(const char* b, const char* e, Record& r) {
  const char* sp = (const char*)memchr(b, ' ', (size_t)(e - b));
  const char* id_end = sp ? sp : e;
  for (; id_end > b; --id_end) {
    if (id_end >= e || !is_space((unsigned char)*id_end)) continue;  // group 1 must be followed by \s
    const char* q = id_end;
    Record t = r;
    if (ws_number(q, e, t.r_len) && ws_number(q, e, t.r_al_len) && literal(q, e, "bp") && ws_number(q, e, t.r_st) &&
        literal(q, e, "-") && ws_number(q, e, t.r_en, false) && ws_rest(q, e, t.r_al)) {
      t.id.p = b;
      t.id.n = (size_t)(id_end - b);
      r = t;
      return true;
    }
  }
  return false;
}

// ^([^+-]+)([+-])\s+(\d+)bp\s+score=(\d+)\s+(.+)$   (ncrf_parser.py:75)
bool parse_second(const char* b, const char* e, Record& r) {
  const char* q = b;
  while (q < e && *q != '+' && *q != '-') ++q;
  if (q == b || q >= e) return false;
  r.motif.p = b;
  r.motif.n = (size_t)(q - b);
  r.strand = *q++;
  return ws_number(q, e, r.m_al_len) && literal(q, e, "bp") && [&] {
    const char* s = q;
    while (s < e && is_space((unsigned char)*s)) ++s;
    if (s == q) return false;
    q = s;
    return literal(q, e, "score=");
  }() && ws_number(q, e, r.score, false) && ws_rest(q, e, r.m_al);
}

inline char complement(char c) {
  switch (c) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'G': return 'C';
    case 'C': return 'G';
    case 'a': return 't';
    case 't': return 'a';
    case 'g': return 'c';
    case 'c': return 'g';
    default: return c;  // '-' and everything else pass through (utils/bio.py:27-29)
  }
}

inline int code_of(char c) {
  switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
  }
}

// byte -> 2-bit code of the base it contributes to the forward-oriented gap-free read row: 0..3, GAP, or BAD
constexpr uint8_t LUT_GAP = 0xFE, LUT_BAD = 0xFF;
struct Luts {
  uint8_t code_fwd[256], code_rc[256];   // read row, as stored / reverse-complemented
  uint8_t upper_fwd[256], upper_rc[256];  // motif row, upper-cased (after complementing for '-' records)
  Luts() {
    for (int c = 0; c < 256; ++c) {
      const int f = code_of((char)c), r = code_of(complement((char)c));
      code_fwd[c] = c == '-' ? LUT_GAP : (f < 0 ? LUT_BAD : (uint8_t)f);
      code_rc[c] = c == '-' ? LUT_GAP : (r < 0 ? LUT_BAD : (uint8_t)r);
      auto up = [](char ch) { return (uint8_t)((ch >= 'a' && ch <= 'z') ? ch - 'a' + 'A' : ch); };
      upper_fwd[c] = up((char)c);
      upper_rc[c] = up(complement((char)c));
    }
  }
};
const Luts g_luts;

struct PerRead {
  int64_t len = 0;                 // gap-free length
  std::vector<int64_t> bounds;     // gap-free unit boundaries (may be empty)
  std::string error;
};

}  // namespace

struct cfk_ncrf {
  // the report text: the file mapped read-only (no copy, no zero-filled buffer: 0.1 s of a 320 MB report), or -- what
  // cannot be mapped (pipes, empty files) -- read into `text`
  const char* data = nullptr;
  size_t size = 0;
  void* mapped = nullptr;
  std::vector<char> text;
  ~cfk_ncrf() {
    if (mapped) munmap(mapped, size);
  }
  std::vector<Record> kept;         // insertion order of the reference's dict
  std::vector<PerRead> per;
  std::vector<int64_t> read_off;    // bases, 64-aligned
  int64_t n_words = 0;
  int64_t n_bases = 0;
  int64_t n_units = 0;
  int64_t ids_bytes = 0;
  int n_per_match = 1;
  int n_threads = 1;
  int64_t n_pairs = 0;              // records seen in the file
};

namespace {

int ingest_fail(int code, const std::string& what) {
  snprintf(g_ingest_err, sizeof(g_ingest_err), "%s", what.c_str());
  return code;
}

template <class F>
void parallel_for(int64_t n, int threads, F&& body) {
  if (threads <= 1 || n <= 1) {
    for (int64_t i = 0; i < n; ++i) body(i);
    return;
  }
  std::atomic<int64_t> next{0};
  std::vector<std::thread> pool;
  const int t = (int)std::min<int64_t>(threads, n);
  for (int w = 0; w < t; ++w)
    pool.emplace_back([&] {
      for (;;) {
        const int64_t i = next.fetch_add(1);
        if (i >= n) break;
        body(i);
      }
    });
  for (auto& th : pool) th.join();
}

// forward-oriented view of an alignment row: column c of the (possibly reverse-complemented) row
struct Row {
  const char* p;
  size_t n;
  bool rc;
  inline char at(size_t c) const { return rc ? complement(p[n - 1 - c]) : p[c]; }
};

// unit boundaries of one record in gap-free read offsets (centroflye_b200/ncrf_parser.py motif_unit_columns +
// ingest.units_from_report)
void segment(const Record& rec, int n_per_match, PerRead& out) {
  const bool rc = rec.strand == '-';
  const Row m{rec.m_al.p, rec.m_al.n, rc}, r{rec.r_al.p, rec.r_al.n, rc};
  const size_t motif_len = rec.motif.n;
  std::string pattern;
  pattern.reserve(motif_len * (size_t)n_per_match);
  for (int i = 0; i < n_per_match; ++i) pattern.append(rec.motif.p, motif_len);
  if (pattern.empty()) return;
  // gap-free upper-cased motif row + the column of each of its symbols
  std::string flat;
  std::vector<uint32_t> cols;
  flat.reserve(m.n);
  cols.reserve(m.n);
  flat.resize(m.n);
  cols.resize(m.n);
  {
    const uint8_t* up = rc ? g_luts.upper_rc : g_luts.upper_fwd;
    size_t nf = 0;
    for (size_t c = 0; c < m.n; ++c) {
      const uint8_t ch = up[(uint8_t)(rc ? m.p[m.n - 1 - c] : m.p[c])];
      flat[nf] = (char)ch;
      cols[nf] = (uint32_t)c;
      nf += ch != '-';
    }
    flat.resize(nf);
    cols.resize(nf);
  }
  std::vector<int64_t> coords;
  size_t last_end = 0;
  for (size_t from = 0; from + pattern.size() <= flat.size();) {
    const void* hit = memmem(flat.data() + from, flat.size() - from, pattern.data(), pattern.size());
    if (!hit) break;
    const size_t q = (size_t)((const char*)hit - flat.data());
    coords.push_back((int64_t)cols[q]);
    last_end = q + pattern.size();
    from = last_end;
  }
  if (coords.empty()) return;
  coords.push_back(last_end < cols.size() ? (int64_t)cols[last_end] : (int64_t)m.n);
  const double slack = (double)motif_len * 0.2;  // ncrf_parser.py:49-52, same float64 expressions
  const int64_t r_cols = (int64_t)r.n;
  if ((double)coords.front() > slack) coords.insert(coords.begin(), 0);
  if ((double)coords.back() < (double)r_cols - slack) coords.push_back(r_cols);
  // alignment columns -> gap-free read offsets
  out.bounds.resize(coords.size());
  size_t ci = 0;
  int64_t before = 0;
  for (size_t c = 0; c <= r.n && ci < coords.size(); ++c) {
    while (ci < coords.size() && coords[ci] == (int64_t)c) out.bounds[ci++] = before;
    if (c < r.n && r.at(c) != '-') ++before;
  }
  if (ci < coords.size()) {
    out.bounds.clear();
    out.error = "unit boundary behind the end of the read row (r_al shorter than m_al)";
  }
}

}  // namespace

extern "C" {

const char* cfk_ncrf_last_error(void) { return g_ingest_err; }

int cfk_ncrf_open(const char* path, int64_t min_record_len, int32_t n_per_match, int32_t n_threads, cfk_ncrf_t** out) {
  if (!path || !out || n_per_match < 1) return ingest_fail(CFK_ERR_INVALID, "cfk_ncrf_open: bad arguments");
  *out = nullptr;
  std::unique_ptr<cfk_ncrf> ctx(new cfk_ncrf);
  {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return ingest_fail(CFK_ERR_INVALID, std::string("cfk_ncrf_open: cannot open ") + path);
    struct stat st;
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
      void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m != MAP_FAILED) {
        ctx->mapped = m;
        ctx->data = (const char*)m;
        ctx->size = (size_t)st.st_size;
        madvise(m, ctx->size, MADV_WILLNEED);
      }
    }
    if (!ctx->mapped) {  // not mappable: read to the end
      char buf[1 << 16];
      for (;;) {
        const ssize_t got = read(fd, buf, sizeof(buf));
        if (got < 0) {
          close(fd);
          return ingest_fail(CFK_ERR_INVALID, std::string("cfk_ncrf_open: short read of ") + path);
        }
        if (got == 0) break;
        ctx->text.insert(ctx->text.end(), buf, buf + got);
      }
      ctx->data = ctx->text.data();
      ctx->size = ctx->text.size();
    }
    close(fd);
  }
  ctx->n_per_match = n_per_match;
  ctx->n_threads = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());

  // lines: strip, drop empty and '#' lines (ncrf_parser.py:65-68), then pair them up
  std::vector<Span> lines;
  const char* p = ctx->data;
  const char* end = p + ctx->size;
  while (p < end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    const char* b = p;
    const char* e = le;
    while (b < e && is_space((unsigned char)*b)) ++b;
    while (e > b && is_space((unsigned char)e[-1])) --e;
    if (e > b && *b != '#') lines.push_back(Span{b, (size_t)(e - b)});
    p = nl ? nl + 1 : end;
  }
  if (lines.size() % 2) return ingest_fail(CFK_ERR_INVALID, "cfk_ncrf_open: dangling record line (odd number of lines)");
  ctx->n_pairs = (int64_t)(lines.size() / 2);

  std::unordered_map<std::string, size_t> slot;  // r_id -> index in kept
  for (size_t i = 0; i < lines.size(); i += 2) {
    Record rec;
    if (!parse_first(lines[i].p, lines[i].p + lines[i].n, rec))
      return ingest_fail(CFK_ERR_INVALID, "cfk_ncrf_open: malformed read line of record " + std::to_string(i / 2));
    if (!parse_second(lines[i + 1].p, lines[i + 1].p + lines[i + 1].n, rec))
      return ingest_fail(CFK_ERR_INVALID, "cfk_ncrf_open: malformed motif line of record " + std::to_string(i / 2));
    std::string id(rec.id.p, rec.id.n);
    auto it = slot.find(id);
    if (it != slot.end() && ctx->kept[it->second].r_al_len >= rec.r_al_len) continue;  // ncrf_parser.py:91
    if (rec.r_al_len < min_record_len) continue;                                        // :93
    if (rec.strand == '-') {                                                            // :96-100 (rows are flipped lazily)
      const int64_t st = rec.r_len - rec.r_en, en = rec.r_len - rec.r_st;
      rec.r_st = st;
      rec.r_en = en;
    }
    if (it != slot.end()) {
      ctx->kept[it->second] = rec;  // a dict keeps the position of the first insertion
    } else {
      slot.emplace(std::move(id), ctx->kept.size());
      ctx->kept.push_back(rec);
    }
  }

  const int64_t R = (int64_t)ctx->kept.size();
  ctx->per.resize((size_t)R);
  parallel_for(R, ctx->n_threads, [&](int64_t i) {
    const Record& rec = ctx->kept[(size_t)i];
    PerRead& pr = ctx->per[(size_t)i];
    int64_t len = 0;
    for (size_t c = 0; c < rec.r_al.n; ++c) len += rec.r_al.p[c] != '-';
    pr.len = len;
    segment(rec, ctx->n_per_match, pr);
  });
  ctx->read_off.resize((size_t)R);
  int64_t off = 0;
  for (int64_t i = 0; i < R; ++i) {
    const PerRead& pr = ctx->per[(size_t)i];
    if (!pr.error.empty())
      return ingest_fail(CFK_ERR_INVALID, "cfk_ncrf_open: record " + std::string(ctx->kept[(size_t)i].id.p, ctx->kept[(size_t)i].id.n) +
                                              ": " + pr.error);
    ctx->read_off[(size_t)i] = off;
    off += (pr.len + 63) / 64 * 64;
    ctx->n_bases += pr.len;
    ctx->n_units += pr.bounds.empty() ? 0 : (int64_t)pr.bounds.size() - 1;
    ctx->ids_bytes += (int64_t)ctx->kept[(size_t)i].id.n + 1;
  }
  ctx->n_words = (off + 64) / 16;  // one spare 64-base line behind the last read (ingest.pack_reads)
  *out = ctx.release();
  return CFK_OK;
}

int64_t cfk_ncrf_n_records(const cfk_ncrf_t* ctx) { return ctx ? (int64_t)ctx->kept.size() : -1; }
int64_t cfk_ncrf_n_seen(const cfk_ncrf_t* ctx) { return ctx ? ctx->n_pairs : -1; }
int64_t cfk_ncrf_n_words(const cfk_ncrf_t* ctx) { return ctx ? ctx->n_words : -1; }
int64_t cfk_ncrf_n_bases(const cfk_ncrf_t* ctx) { return ctx ? ctx->n_bases : -1; }
int64_t cfk_ncrf_n_units(const cfk_ncrf_t* ctx) { return ctx ? ctx->n_units : -1; }
int64_t cfk_ncrf_ids_bytes(const cfk_ncrf_t* ctx) { return ctx ? ctx->ids_bytes : -1; }

int cfk_ncrf_export(const cfk_ncrf_t* ctx, uint32_t* packed_h, int64_t* read_off_h, int64_t* read_len_h,
                    int64_t* read_unit_ptr_h, int64_t* unit_off_h, int32_t* unit_len_h, int32_t* unit_read_h,
                    char* ids_h, int64_t* fields_h) {
  if (!ctx || !packed_h || !read_off_h || !read_len_h || !read_unit_ptr_h)
    return ingest_fail(CFK_ERR_INVALID, "cfk_ncrf_export: bad arguments");
  const int64_t R = (int64_t)ctx->kept.size();
  memset(packed_h, 0, (size_t)ctx->n_words * 4);
  int64_t u = 0;
  char* idp = ids_h;
  for (int64_t i = 0; i < R; ++i) {
    const PerRead& pr = ctx->per[(size_t)i];
    const Record& rec = ctx->kept[(size_t)i];
    read_off_h[i] = ctx->read_off[(size_t)i];
    read_len_h[i] = pr.len;
    read_unit_ptr_h[i] = u;
    const int64_t nu = pr.bounds.empty() ? 0 : (int64_t)pr.bounds.size() - 1;
    for (int64_t j = 0; j < nu; ++j) {
      if (unit_off_h) unit_off_h[u + j] = ctx->read_off[(size_t)i] + pr.bounds[(size_t)j];
      if (unit_len_h) unit_len_h[u + j] = (int32_t)(pr.bounds[(size_t)j + 1] - pr.bounds[(size_t)j]);
      if (unit_read_h) unit_read_h[u + j] = (int32_t)i;
    }
    u += nu;
    if (idp) {
      memcpy(idp, rec.id.p, rec.id.n);
      idp += rec.id.n;
      *idp++ = '\n';
    }
    if (fields_h) {
      int64_t* fld = fields_h + 8 * i;
      fld[0] = rec.r_len; fld[1] = rec.r_al_len; fld[2] = rec.r_st; fld[3] = rec.r_en;
      fld[4] = rec.strand == '-' ? -1 : 1; fld[5] = rec.m_al_len; fld[6] = rec.score; fld[7] = (int64_t)rec.r_al.n;
    }
  }
  read_unit_ptr_h[R] = u;
  // gap removal + 2-bit packing; every read owns whole 64-base (4-word) lines, so reads never share a word
  std::atomic<int64_t> bad_read{-1};
  std::vector<int64_t> bad_col((size_t)std::max<int64_t>(R, 1), -1);
  parallel_for(R, ctx->n_threads, [&](int64_t i) {
    const Record& rec = ctx->kept[(size_t)i];
    const Row r{rec.r_al.p, rec.r_al.n, rec.strand == '-'};
    uint32_t* w = packed_h + (ctx->read_off[(size_t)i] >> 4);
    uint32_t acc = 0;
    int64_t nb = 0;
    const uint8_t* lut = r.rc ? g_luts.code_rc : g_luts.code_fwd;
    for (size_t c = 0; c < r.n; ++c) {
      const uint8_t code = lut[(uint8_t)(r.rc ? r.p[r.n - 1 - c] : r.p[c])];
      if (code == LUT_GAP) continue;
      if (code == LUT_BAD) {
        bad_col[(size_t)i] = nb;
        int64_t expect = -1;
        bad_read.compare_exchange_strong(expect, i);
        return;
      }
      acc |= (uint32_t)code << ((nb & 15) << 1);
      if ((++nb & 15) == 0) {
        *w++ = acc;
        acc = 0;
      }
    }
    if (nb & 15) *w = acc;
  });
  if (bad_read.load() >= 0) {
    int64_t first = -1;  // report the first offending record in record order, like the sequential host path
    for (int64_t i = 0; i < R && first < 0; ++i)
      if (bad_col[(size_t)i] >= 0) first = i;
    const Record& rec = ctx->kept[(size_t)first];
    return ingest_fail(CFK_ERR_INVALID, "non-ACGT symbol at offset " + std::to_string(bad_col[(size_t)first]) + " of record " +
                                            std::string(rec.id.p, rec.id.n) +
                                            ": the 2-bit device path only accepts upper-case A/C/G/T");
  }
  return CFK_OK;
}

void cfk_ncrf_close(cfk_ncrf_t* ctx) { delete ctx; }

}  // extern "C"
