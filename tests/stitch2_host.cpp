// TEST INFRASTRUCTURE ONLY: CPU harness around faucet_b200/csrc/stitch2_walk.cuh.
//
// The thread-per-record junction walk of the CUDA stitch (s2_lookup_line / s2_line / s2_subread) is
// __host__ __device__ code written against an environment type.  This file supplies a host
// environment -- flag planes computed here from a plain Bloom array, an open-addressing junction
// table in ordinary memory -- and runs the records one after the other (= a window of one record per
// round), so tests/test_stitch2_host.py can hold that code to the oracle without a GPU.
// Clean inputs only (complete 2- / 4-line records, lines of at most S2_POS_CAP k-mer positions).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../faucet_b200/csrc/kmer.cuh"
#include "../faucet_b200/csrc/pair_filter_host.hpp"
#include "stitch2_walk.cuh"

using namespace faucet;

namespace {

struct HostBits {
  const uint8_t* bits; uint64_t mask; int n_hash;
  bool contains(uint64_t c) const {  // Bloom::oldContains -> contains(h0, h1)
    uint64_t h = hash0(c) & mask, h1 = hash1(c) & mask;
    for (int i = 0; i < n_hash; i++, h = (h + h1) & mask)
      if (!(bits[h >> 3] & (1u << (h & 7)))) return false;
    return true;
  }
};

struct Planes {
  std::vector<uint32_t> inval, packed, fp;
  std::vector<uint32_t> seq_start, seq_end;
};

static bool jcheck(const HostBits& b, uint64_t x, int k, int j) {  // JChecker::jcheck, utils/JChecker.cpp:51-80
  if (j == 0) return true;
  const uint64_t mask = kmer_mask(k);
  for (uint32_t c = 0; c < 4; c++) {
    const uint64_t y = ext_fwd(x, c, mask);
    if (b.contains(canon(y, revcomp(y, k))) && jcheck(b, y, k, j - 1)) return true;
  }
  return false;
}

static void build_planes(const char* text, size_t n, int fastq, int k, int j, const HostBits& bloom, Planes& P) {
  const size_t words = (n + 31) / 32 + 4;
  P.inval.assign(words, 0xffffffffu);
  P.packed.assign((n + 15) / 16 + 8, 0u);
  P.fp.assign(words * FP_STRIDE, 0u);
  for (size_t p = 0; p < n; p++) P.packed[p >> 4] |= nt_code((uint8_t)text[p]) << (30 - 2 * (p & 15));
  // records: the sequence line is line 1 of every group of 4 (fastq) or 2 (fasta) lines
  const int per = fastq ? 4 : 2;
  size_t line_start = 0;
  long line_no = 0;
  for (size_t p = 0; p <= n; p++) {
    if (p == n || text[p] == '\n') {
      if (p == n && line_start == n) break;
      if (line_no % per == 1) {
        P.seq_start.push_back((uint32_t)line_start);
        P.seq_end.push_back((uint32_t)p);
        for (size_t q = line_start; q < p; q++)
          if (nt_valid((uint8_t)text[q])) P.inval[q >> 5] &= ~(1u << (q & 31));
      }
      line_no++;
      line_start = p + 1;
    }
  }
  auto inv = [&](size_t p) { return (P.inval[p >> 5] >> (p & 31)) & 1u; };
  auto setp = [&](int plane, size_t p) { P.fp[(p >> 5) * FP_STRIDE + plane] |= 1u << (p & 31); };
  const uint64_t mask = kmer_mask(k);
  size_t run = 0;  // valid bases ending at p
  for (size_t e = 0; e < n; e++) {
    run = inv(e) ? 0 : run + 1;
    if (run < (size_t)k) continue;
    const size_t p = e + 1 - k;
    uint64_t fwd = 0;
    for (int i = 0; i < k; i++) fwd = (fwd << 2) | nt_code((uint8_t)text[p + i]);
    const uint64_t rc = revcomp(fwd, k);
    if (!bloom.contains(canon(fwd, rc))) continue;
    setp(FP_V, p);
    for (int d = 0; d < 2; d++) {  // testForJunction (src/ReadScanner.cpp:36-56): d = 0 FORWARD, 1 BACKWARD
      uint32_t real;
      if (d == 0) { if (e + 1 >= n || inv(e + 1)) continue; real = nt_code((uint8_t)text[e + 1]); }
      else { if (p == 0 || inv(p - 1)) continue; real = nt_comp(nt_code((uint8_t)text[p - 1])); }
      const uint64_t base = d == 0 ? fwd : rc;
      unsigned cnt = 0; bool junc = false;
      for (uint32_t c = 0; c < 4 && !junc; c++) {
        if (c == real) continue;
        const uint64_t y = ext_fwd(base, c, mask);
        if (!bloom.contains(canon(y, revcomp(y, k)))) continue;
        cnt++;
        if (jcheck(bloom, y, k, j)) junc = true;
      }
      if (junc) setp(d == 0 ? FP_JF : FP_JB, p);
      if (cnt & 1) setp(d == 0 ? FP_CF0 : FP_CB0, p);
      if (cnt & 2) setp(d == 0 ? FP_CF1 : FP_CB1, p);
    }
  }
}

struct HostEnv {
  const Planes& P;
  int k, j, spacer;
  bool pairs, want_ext;
  unsigned long long st[S2_COUNTERS];
  // junction table
  uint64_t cap;
  std::vector<uint64_t> keys, stamps;
  std::vector<uint32_t> recs;  // 16 u32 per slot: dist[5], link mask, cov[4]
  bool special = false;
  unsigned long long stamp_next = 0;
  uint64_t n_entries = 0;
  // outputs
  HostBloom* spf = nullptr;
  std::vector<uint64_t> ext;
  uint32_t rec = 0, part = 0;
  std::vector<uint64_t> ext_buf;
  bool changed = false;  // the current record created a junction or raised a stored distance (diagnostic)

  HostEnv(const Planes& p, uint64_t cap_) : P(p), cap(cap_), keys(cap_ + 1, S2_KEY_EMPTY), stamps(cap_ + 1, 0), recs((cap_ + 1) * 16, 0) {
    for (auto& x : st) x = 0;
  }
  static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
  }
  uint32_t inval_word(uint32_t w) const { return P.inval[w]; }
  uint32_t fp_word(int p, uint32_t w) const { return P.fp[(size_t)w * FP_STRIDE + p]; }
  uint32_t packed_word(uint32_t w) const { return P.packed[w]; }
  uint64_t tbl_home(uint64_t key) const { return mix64(key) & (cap - 1); }
  uint64_t tbl_next(uint64_t h) const { return (h + 1) & (cap - 1); }
  uint64_t tbl_key(uint64_t h) const { return keys[h]; }
  int find(uint64_t key) const {
    if (key == S2_KEY_EMPTY) return special ? (int)cap : -1;
    uint64_t h = tbl_home(key);
    while (true) {
      if (keys[h] == key) return (int)h;
      if (keys[h] == S2_KEY_EMPTY) return -1;
      h = tbl_next(h);
    }
  }
  int insert(uint64_t key, bool* created) {
    if (key == S2_KEY_EMPTY) {
      *created = !special;
      if (*created) { special = true; keys[cap] = key; n_entries++; }
      return (int)cap;
    }
    uint64_t h = tbl_home(key);
    while (true) {
      if (keys[h] == S2_KEY_EMPTY) { keys[h] = key; *created = true; changed = true; n_entries++; return (int)h; }
      if (keys[h] == key) { *created = false; return (int)h; }
      h = tbl_next(h);
    }
  }
  void stamp(int slot) { stamps[slot] = stamp_next++; }
  uint32_t dist_peek(int slot, int idx) const { return recs[(size_t)slot * 16 + idx]; }
  void add_cov(int slot, int nt) { recs[(size_t)slot * 16 + 6 + nt]++; }
  void update(int slot, int idx, int length) {
    uint32_t& d = recs[(size_t)slot * 16 + idx];
    const uint32_t v = (uint32_t)length & 0xffu;
    if (v > d) { d = v; changed = true; }
  }
  void link(int slot, int idx) { recs[(size_t)slot * 16 + 5] |= 1u << idx; }
  uint32_t dist_now(int slot, int idx) const { return recs[(size_t)slot * 16 + idx]; }
  void spf_pair(uint64_t k1, uint64_t k2) { spf->add_pair(k1, k2, k); }
  bool aborted() const { return false; }
  void ext_flush() {
    ext.push_back(((uint64_t)rec << 32) | ((uint64_t)(part & 0xffffu) << 16) | ext_buf.size());
    ext.insert(ext.end(), ext_buf.begin(), ext_buf.end());
    ext_buf.clear();
    part++;
  }
  void ext_push(uint64_t kmer) {
    ext_buf.push_back(kmer);
    if (ext_buf.size() == (size_t)S2_EXT) ext_flush();
  }
};

}  // namespace

extern "C" {

struct s2h_rec { uint64_t kmer; uint8_t dist[5], cov[4], linked[5], pad[2]; };  // == fo_junction_rec
struct s2h_stats { uint64_t n_junctions, nb_jcheck_kmer, nb_no_juncs, nb_processed, nb_skipped, reads_no_errors, reads_processed, unambiguous_reads; };

// returns 0, or -1 when a line is too long for the thread path
int s2h_scan(const char* text, size_t n, int fastq, int paired, int no_cleaning, int k, int j, int spacer,
             const uint8_t* bloo2, int log2_tai, int n_hash, uint8_t* short_pf, int spf_log2, int spf_nh,
             uint8_t* long_pf, int lpf_log2, int lpf_nh, s2h_rec** recs_out, uint64_t* n_out, s2h_stats* stats,
             uint64_t* quiet_profile /* NULL, or 20 counters: records per 5 % of the stream that changed visible state */) {
  HostBits bloom{bloo2, (1ull << log2_tai) - 1, n_hash};
  Planes P;
  build_planes(text, n, fastq, k, j, bloom, P);
  uint64_t cap = 1024;
  while (cap < 8 * (uint64_t)P.seq_start.size() + 1024) cap <<= 1;
  HostEnv e(P, cap);
  e.k = k; e.j = j; e.spacer = spacer;
  HostBloom spf;
  LongPairFilter lpf;
  e.pairs = !no_cleaning && short_pf != nullptr;
  if (e.pairs) { spf.bits = short_pf; spf.mask = (1ull << spf_log2) - 1; spf.n_hash = spf_nh; e.spf = &spf; }
  if (long_pf && paired && !no_cleaning) lpf.init(long_pf, lpf_log2, lpf_nh, k);
  e.want_ext = lpf.enabled();
  LineState L;
  for (size_t r = 0; r < P.seq_start.size(); r++) {
    const uint32_t ls = P.seq_start[r], le = P.seq_end[r];
    const int n_pos = (int)(le - ls) - k + 1;
    if (n_pos > S2_POS_CAP) return -1;
    s2_lookup_line(e, L, ls, n_pos > 0 ? n_pos : 0);
    e.rec = (uint32_t)r; e.part = 0;
    e.stamp_next = (unsigned long long)r << 20;
    e.changed = false;
    const bool quiet = s2_is_quiet(e, L, ls, le);  // read-only dry run first: it must predict the real walk
    s2_line(e, L, ls, le);
    if (quiet == e.changed) return -2;
    if (quiet_profile && e.changed) quiet_profile[r * 20 / P.seq_start.size()]++;
    if (e.want_ext && !e.ext_buf.empty()) e.ext_flush();
  }
  if (e.want_ext) lpf.process_batch(e.ext.data(), e.ext.size(), (uint32_t)P.seq_start.size(), 0);
  // creation order = stamp order
  std::vector<std::pair<uint64_t, uint64_t>> order;
  for (uint64_t h = 0; h <= cap; h++)
    if (h == cap ? e.special : e.keys[h] != S2_KEY_EMPTY) order.push_back({e.stamps[h], h});
  std::sort(order.begin(), order.end());
  s2h_rec* out = (s2h_rec*)calloc(order.size() + 1, sizeof(s2h_rec));
  for (size_t i = 0; i < order.size(); i++) {
    const uint64_t h = order[i].second;
    const uint32_t* rr = &e.recs[h * 16];
    out[i].kmer = e.keys[h];
    for (int f = 0; f < 5; f++) { out[i].dist[f] = (uint8_t)rr[f]; out[i].linked[f] = (rr[5] >> f) & 1u; }
    for (int f = 0; f < 4; f++) out[i].cov[f] = (uint8_t)(rr[6 + f] > 255u ? 255u : rr[6 + f]);
  }
  *recs_out = out; *n_out = order.size();
  stats->n_junctions = order.size();
  stats->nb_jcheck_kmer = e.st[S2_JCHECK]; stats->nb_no_juncs = e.st[S2_NOJUNC]; stats->nb_processed = e.st[S2_PROCESSED];
  stats->nb_skipped = e.st[S2_SKIPPED]; stats->reads_no_errors = e.st[S2_NOERR]; stats->unambiguous_reads = e.st[S2_UNAMBIG];
  stats->reads_processed = P.seq_start.size();
  return 0;
}
void s2h_free(void* p) { free(p); }
}
