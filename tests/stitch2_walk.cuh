// Pass 2, stream-order part, formulated for ONE THREAD per record (stitch2.cuh runs it on the GPU).
//
// scan_forward (src/ReadScanner.cpp:112-231) visits half-steps one by one and asks three questions at
// each: is the oriented k-mer a known junction (JunctionMap lookup, :67), has the spacer distance
// elapsed (:72), does testForJunction fire (:76)?  The third answer is a pure function of the text and
// is already a bit plane (scan_flags); the first one is a bit plane too once the line's keys have been
// looked up at the start of the round (the table cannot change under an executing record except
// through the record itself, see stitch.cuh); the second is a threshold.  So "advance to the next
// junction" is a find-first-set over (J | K) planes clipped by the spacer threshold, and the counters
// the reference keeps per visited half-step (NbProcessed, NbJCheckKmer) are range popcounts.  A record
// then costs a handful of junction events instead of ~200 half-step iterations, which is what makes
// one thread per record (instead of one warp) affordable.
//
// Everything here is __host__ __device__ and written against an environment type E, so the same code
// TEST ASSET since round 2 (its CUDA kernel was retired): runs in the CPU harness of the test-suite
// (tests/stitch2_host.cpp), where it is checked against the oracle.
#pragma once
#include <stdint.h>

#include "../faucet_b200/csrc/kmer.cuh"

namespace faucet {

constexpr int S2_POS_CAP = 128;         // k-mer positions per line on the thread path (longer lines: warp path)
constexpr int S2_KW = S2_POS_CAP / 32;  // words per K plane
constexpr int S2_PARK = 8;              // known junctions of a line whose slot / skip distance are parked
constexpr int S2_ROW = 32;              // u32 per record in the reservation rows: [0] = count, [1..31] = slots
constexpr int S2_EXT = 8;               // real-extension k-mers staged per thread before a chunk is flushed
constexpr unsigned long long S2_KEY_EMPTY = ~0ull;

// flag planes written by scan_flags_kernel: word w of plane i = fplanes[8 * w + i], bit b <-> byte offset 32 w + b
enum { FP_V = 0, FP_JF, FP_JB, FP_CF0, FP_CF1, FP_CB0, FP_CB1, FP_STRIDE = 8 };
enum { S2_JCHECK = 0, S2_NOJUNC, S2_PROCESSED, S2_SKIPPED, S2_NOERR, S2_UNAMBIG, S2_COUNTERS };

FHD int s2_popc(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
FHD int s2_ctz(uint32_t x) {  // x != 0
#ifdef __CUDA_ARCH__
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
FHD int s2_clz(uint32_t x) {  // x != 0
#ifdef __CUDA_ARCH__
  return __clz((int)x);
#else
  return __builtin_clz(x);
#endif
}

// per-record state that lives from the lookups (phase 1) to the walk (phase 2)
struct LineState {
  uint32_t kf[S2_KW], kb[S2_KW];  // K planes: bit r <=> the FORWARD / BACKWARD key at position ls + r is a junction
  uint32_t pslot[S2_PARK];        // parked: table slot ...
  uint32_t pinfo[S2_PARK];        // ... of the key at line half-step t = 2 r + dir, and its dist[fwdIdx] when the round started: t << 8 | dist
  uint32_t n_park;                // hits seen (only the first S2_PARK are parked)
  uint32_t pstale;                // bit i: the line has touched that junction since (the parked distance may be stale)
};
FHD int s2_parked(const LineState& L) { return L.n_park < (uint32_t)S2_PARK ? (int)L.n_park : S2_PARK; }

// ---- bit-plane views ------------------------------------------------------------------------------
template <class E, int P>
struct FpView {
  const E& e;
  FHD uint32_t word(uint32_t w) const { return e.fp_word(P, w); }
};
template <class E>
struct InvView {
  const E& e;
  FHD uint32_t word(uint32_t w) const { return e.inval_word(w); }
};
struct KView {
  const uint32_t* p;
  FHD uint32_t word(uint32_t w) const { return p[w]; }
};

// first position in [from, to) whose bit equals WANT, or `to`
template <bool WANT, class V>
FHD uint32_t s2_next_bit(const V& v, uint32_t from, uint32_t to) {
  if (from >= to) return to;
  uint32_t w = from >> 5;
  const uint32_t wl = (to - 1) >> 5;
  uint32_t x = v.word(w);
  if (!WANT) x = ~x;
  x &= ~0u << (from & 31);
  while (true) {
    if (x) {
      const uint32_t b = (w << 5) + (uint32_t)s2_ctz(x);
      return b < to ? b : to;
    }
    if (w == wl) return to;
    x = v.word(++w);
    if (!WANT) x = ~x;
  }
}
// highest position in [s, pos) whose bit equals `want`, or -1
template <class V>
FHD long long s2_prev_bit(const V& v, uint32_t s, uint32_t pos, bool want) {
  if (pos <= s) return -1;
  const uint32_t p = pos - 1, ws = s >> 5;
  uint32_t w = p >> 5;
  uint32_t word = v.word(w);
  if (!want) word = ~word;
  if ((p & 31) != 31) word &= (2u << (p & 31)) - 1u;
  while (true) {
    if (w == ws) word &= ~((1u << (s & 31)) - 1u);
    if (word) return ((long long)w << 5) + 31 - s2_clz(word);
    if (w == ws) return -1;
    w--;
    word = v.word(w);
    if (!want) word = ~word;
  }
}
template <class V>
FHD int s2_popc_range(const V& v, uint32_t from, uint32_t to) {
  if (from >= to) return 0;
  const uint32_t w0 = from >> 5, wl = (to - 1) >> 5;
  int c = 0;
  for (uint32_t w = w0; w <= wl; w++) {
    uint32_t x = v.word(w);
    if (w == w0) x &= ~0u << (from & 31);
    if (w == wl && (to & 31)) x &= (1u << (to & 31)) - 1u;
    c += s2_popc(x);
  }
  return c;
}
template <class V>
FHD uint32_t s2_bit(const V& v, uint32_t p) { return (v.word(p >> 5) >> (p & 31)) & 1u; }

// k-mer starting at byte offset p of the big-endian 2-bit plane (16 bases per u32)
template <class E>
FHD uint64_t s2_kmer_at(const E& e, uint32_t p, int k) {
  const uint32_t w = p >> 4, o = 2 * (p & 15);
  const uint64_t hi = ((uint64_t)e.packed_word(w) << 32) | e.packed_word(w + 1);
  const uint64_t lo = (uint64_t)e.packed_word(w + 2) << 32;
  const uint64_t x = o ? ((hi << o) | (lo >> (64 - o))) : hi;
  return x >> (64 - 2 * k);
}
template <class E>
FHD uint32_t s2_code_at(const E& e, uint32_t p) { return (e.packed_word(p >> 4) >> (30 - 2 * (p & 15))) & 3u; }

// ---- phase 1: which keys of the line are junctions right now ------------------------------------------
// Looks up the FORWARD and BACKWARD key of every k-mer position of the line [ls, ls + n_pos + k - 1)
// and records the answers as bit planes; the first S2_PARK hits also park their slot and skip distance.
// Positions whose window holds a non-base produce an arbitrary key: a chance hit there sets a bit that
// no walk ever reads (the walk only visits positions of valid sub-reads).
template <class E>
FHD void s2_lookup_line(const E& e, LineState& L, uint32_t ls, int n_pos) {
  const int k = e.k;
  const uint64_t mask = kmer_mask(k);
  for (int i = 0; i < S2_KW; i++) L.kf[i] = L.kb[i] = 0;
  L.pstale = 0; L.n_park = 0;
  if (n_pos <= 0) return;
  uint64_t f = s2_kmer_at(e, ls, k), r = revcomp(f, k);
  constexpr int U = 4;
  for (int base = 0; base < n_pos; base += U) {
    uint64_t kk[2 * U], hh[2 * U], got[2 * U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      kk[2 * u] = f; kk[2 * u + 1] = r;
      const uint32_t c = s2_code_at(e, ls + base + u + k);  // (one base past the line for the last position: unused)
      f = ext_fwd(f, c, mask);
      r = ext_rc(r, c, k);
    }
#pragma unroll
    for (int u = 0; u < 2 * U; u++) {  // all first probes in flight together
      hh[u] = e.tbl_home(kk[u]);
      got[u] = e.tbl_key(hh[u]);
    }
#pragma unroll
    for (int u = 0; u < 2 * U; u++) {
      const int pos = base + (u >> 1);
      if (pos >= n_pos) break;
      const int dir = (u & 1) ? 0 : 1;  // even entries hold the forward k-mer = the FORWARD key
      long long slot = -1;
      if (kk[u] == S2_KEY_EMPTY) slot = e.find(kk[u]);
      else {
        while (got[u] != kk[u] && got[u] != S2_KEY_EMPTY) { hh[u] = e.tbl_next(hh[u]); got[u] = e.tbl_key(hh[u]); }
        if (got[u] == kk[u]) slot = (long long)hh[u];
      }
      if (slot < 0) continue;
      (dir ? L.kf : L.kb)[pos >> 5] |= 1u << (pos & 31);
      if (L.n_park < (uint32_t)S2_PARK) {
        // dist[fwdIdx]: facing forward fwdIdx = the read's next base, facing backward fwdIdx = 4 (utils/ReadKmer.cpp:95-100)
        const int idx = dir ? (int)s2_code_at(e, ls + pos + k) : 4;
        L.pslot[L.n_park] = (uint32_t)slot;
        L.pinfo[L.n_park] = ((uint32_t)(2 * pos + dir) << 8) | (e.dist_peek((int)slot, idx) & 0xffu);
      }
      L.n_park++;
    }
  }
}

// a key this line just created: every other half-step of the line with that key must see it
template <class E>
FHD void s2_publish(const E& e, LineState& L, uint32_t ls, int n_pos, uint64_t key) {
  const int k = e.k;
  const uint64_t mask = kmer_mask(k), rk = revcomp(key, k);  // revcomp(f) == key  <=>  f == revcomp(key)
  uint64_t f = s2_kmer_at(e, ls, k);
  for (int r = 0; r < n_pos; r++) {
    if (f == key) L.kf[r >> 5] |= 1u << (r & 31);
    if (f == rk) L.kb[r >> 5] |= 1u << (r & 31);
    f = ext_fwd(f, s2_code_at(e, ls + r + k), mask);
  }
}

// sum of the j-checked-alternate counts (scan_flags bits 3-6) over the half-steps [ta, tb) of the
// sub-read that starts at byte offset s0
template <class E>
FHD unsigned s2_cnt_sum(const E& e, uint32_t s0, int ta, int tb) {
  if (ta >= tb) return 0;
  const uint32_t f0 = s0 + (uint32_t)(ta >> 1), f1 = s0 + (uint32_t)(tb >> 1);              // t = 2 pos + 1 in [ta, tb)
  const uint32_t b0 = s0 + (uint32_t)((ta + 1) >> 1), b1 = s0 + (uint32_t)((tb + 1) >> 1);  // t = 2 pos     in [ta, tb)
  return (unsigned)(s2_popc_range(FpView<E, FP_CF0>{e}, f0, f1) + 2 * s2_popc_range(FpView<E, FP_CF1>{e}, f0, f1) +
                    s2_popc_range(FpView<E, FP_CB0>{e}, b0, b1) + 2 * s2_popc_range(FpView<E, FP_CB1>{e}, b0, b1));
}

// scan_forward (src/ReadScanner.cpp:112-231) on the valid sub-read at byte offset s0, `len` bases, of the
// line that starts at ls and has n_pos k-mer positions
template <class E>
FHD void s2_subread(E& e, LineState& L, uint32_t ls, int n_pos, uint32_t s0, int len) {
  const int k = e.k, j = e.j;
  const uint64_t mask = kmer_mask(k);
  const int rel0 = (int)(s0 - ls);
  const int tested_end = 2 * len - 2 * k + 1 - 2 * j;  // distToEnd > 2j  <=>  tp < tested_end
  const int spc_span = 2 * e.spacer - 1;
  int tp = 2 * j + 1, last_junc_pos = 0;
  bool have_last = false, have_fb = false, have_lf = false;
  int last_tp = 0, last_fwd_idx = 0, rev_pos = 0, for_pos = 0, last_slot = -1;
  uint64_t fb_ext = 0, lf_ext = 0, v_prev1 = 0, v_prev2 = 0;  // v[n-1], v[n-2] of this sub-read's result list
  uint32_t n_out = 0;

  auto push_out = [&](uint64_t real_ext) {
    // pairs (v[i], v[i+2]) once the list has more than two entries; a list that ends with exactly two
    // entries is handled after the loop (:208-225)
    if (e.pairs && n_out >= 2) e.spf_pair(v_prev2, real_ext);
    v_prev2 = v_prev1; v_prev1 = real_ext; n_out++;
    if (e.want_ext) e.ext_push(real_ext);
  };
  auto touch = [&](int slot) {  // whatever was parked about this junction may be stale from now on
    for (int i = 0; i < s2_parked(L); i++)
      if (L.pslot[i] == (uint32_t)slot) L.pstale |= 1u << i;
  };

  while (tp < tested_end) {
    // ---- find_next_junction (:61-86)
    const int endF = tested_end >> 1, endB = (tested_end + 1) >> 1;  // positions whose F / B half-step is < tested_end
    const int pF0 = tp >> 1, pB0 = (tp + 1) >> 1;                    // first position whose F / B half-step is >= tp
    int hit = 0x7fffffff;
    if (pF0 < endF) {
      const int a = (int)(s2_next_bit<true>(FpView<E, FP_JF>{e}, s0 + pF0, s0 + endF) - s0);
      const int b = (int)s2_next_bit<true>(KView{L.kf}, (uint32_t)(rel0 + pF0), (uint32_t)(rel0 + endF)) - rel0;
      const int fF = a < b ? a : b;
      if (fF < endF) hit = 2 * fF + 1;
    }
    if (pB0 < endB) {
      const int a = (int)(s2_next_bit<true>(FpView<E, FP_JB>{e}, s0 + pB0, s0 + endB) - s0);
      const int b = (int)s2_next_bit<true>(KView{L.kb}, (uint32_t)(rel0 + pB0), (uint32_t)(rel0 + endB)) - rel0;
      const int fB = a < b ? a : b;
      if (fB < endB && 2 * fB < hit) hit = 2 * fB;
    }
    int ts = last_junc_pos + spc_span;  // first half-step the spacer rule fires at (:72)
    if (ts < tp) ts = tp;
    if (ts < tested_end && ts < hit) hit = ts;
    if (hit == 0x7fffffff) {  // the rest of the tested zone is junction-free
      e.st[S2_PROCESSED] += (unsigned)(tested_end - tp);
      e.st[S2_JCHECK] += s2_cnt_sum(e, s0, tp, tested_end);
      break;
    }
    const int pos = hit >> 1, dir = hit & 1;
    const bool known = (((dir ? L.kf : L.kb)[(rel0 + pos) >> 5] >> ((rel0 + pos) & 31)) & 1u) != 0;
    const bool spc = hit - last_junc_pos >= spc_span;
    // NbJCheckKmer (:46): every half-step that reached testForJunction, the hit one included
    unsigned jc = s2_cnt_sum(e, s0, tp, hit);
    if (!known && !spc)
      jc += dir ? s2_bit(FpView<E, FP_CF0>{e}, s0 + pos) + 2 * s2_bit(FpView<E, FP_CF1>{e}, s0 + pos)
                : s2_bit(FpView<E, FP_CB0>{e}, s0 + pos) + 2 * s2_bit(FpView<E, FP_CB1>{e}, s0 + pos);
    e.st[S2_JCHECK] += jc;
    e.st[S2_PROCESSED] += (unsigned)(hit - tp);
    tp = hit;
    // ---- the junction at half-step tp (:134-192)
    const uint64_t fwd = s2_kmer_at(e, s0 + pos, k);
    const uint64_t key = dir ? fwd : revcomp(fwd, k);
    const int real = dir ? (int)s2_code_at(e, s0 + pos + k) : (int)nt_comp(s2_code_at(e, s0 + pos - 1));
    const int fwd_idx = dir ? real : 4, back_idx = dir ? 4 : real;  // getExtensionIndex (utils/ReadKmer.cpp:95-100)
    int slot = -1, hop = -1;
    bool created = false;
    if (known) {
      const uint32_t tl = (uint32_t)(2 * (rel0 + pos) + dir);
      for (int i = 0; i < s2_parked(L); i++)
        if ((L.pinfo[i] >> 8) == tl) {
          slot = (int)L.pslot[i];
          if (!((L.pstale >> i) & 1u)) hop = (int)(L.pinfo[i] & 0xffu);
          break;
        }
      if (slot < 0) slot = e.find(key);  // beyond the parking space, or published by this line
    }
    if (slot < 0) {
      slot = e.insert(key, &created);
      if (e.aborted()) return;
      if (created) {
        e.stamp(slot);
        s2_publish(e, L, ls, n_pos, key);
      }
    }
    e.add_cov(slot, real);
    if (have_last) {  // directLinkJunctions (utils/JunctionMap.cpp:551-561)
      const int d = tp - last_tp;
      e.update(last_slot, last_fwd_idx, d); e.link(last_slot, last_fwd_idx);
      e.update(slot, back_idx, d); e.link(slot, back_idx);
    } else {
      e.update(slot, back_idx, tp - 2 * j);
    }
    if (e.aborted()) return;
    int dist;
    if (created) dist = 0;  // a zeroed record; back_idx != fwd_idx, so nothing written above shows here
    else if (hop >= 0) dist = hop;
    else dist = (int)e.dist_now(slot, fwd_idx);
    if (dist < 1) dist = 1;
    e.st[S2_PROCESSED] += 1; e.st[S2_SKIPPED] += (unsigned)(dist - 1);
    touch(slot);
    const uint64_t real_ext = ext_fwd(key, (uint32_t)real, mask);
    if (!dir) { if (!have_fb) { have_fb = true; fb_ext = real_ext; rev_pos = pos; } }
    else { if (!have_lf) { have_lf = true; for_pos = pos; } lf_ext = real_ext; }
    push_out(real_ext);
    have_last = true;
    last_junc_pos = tp; last_tp = tp; last_slot = slot; last_fwd_idx = fwd_idx;
    tp += dist;
  }
  if (!have_last) {  // add_fake_junction (:92-104): mid-read, facing forward
    const int pos = len / 2 - k / 2;
    const uint64_t key = s2_kmer_at(e, s0 + pos, k);
    const int real = (int)s2_code_at(e, s0 + pos + k);
    bool created = false;
    e.st[S2_NOJUNC]++;
    const int slot = e.insert(key, &created);
    if (e.aborted()) return;
    if (created) {
      e.stamp(slot);
      s2_publish(e, L, ls, n_pos, key);
    }
    e.add_cov(slot, real);
    const int mtp = 2 * pos + 1;
    e.update(slot, 4, mtp - 2 * j);
    e.update(slot, real, (2 * len - mtp - 2 * k + 1) - 2 * j);
    touch(slot);
    push_out(ext_fwd(key, (uint32_t)real, mask));
  } else {  // :205
    e.update(last_slot, last_fwd_idx, (2 * len - last_tp - 2 * k + 1) - 2 * j);
  }
  if (e.pairs && n_out == 2) {  // :208-218
    if (have_fb && have_lf && !(rev_pos > for_pos)) e.spf_pair(fb_ext, lf_ext);
    if (have_fb != have_lf) e.spf_pair(v_prev2, v_prev1);
  }
}

// scanInputRead (:260-282) + getValidReads (:233-257) for the sequence line [ls, le)
template <class E>
FHD void s2_line(E& e, LineState& L, uint32_t ls, uint32_t le) {
  const int k = e.k, j = e.j;
  const int n_pos = (int)(le - ls) - k + 1;
  uint32_t pos = le;
  while (pos > ls) {  // getUnambiguousReads hands the segments over LAST first (utils/Kmer.cpp:64-80)
    const long long hi = s2_prev_bit(InvView<E>{e}, ls, pos, false);
    if (hi < 0) break;
    const uint32_t ee = (uint32_t)hi + 1;
    const long long lo = s2_prev_bit(InvView<E>{e}, ls, ee, true);
    const uint32_t ss = lo < 0 ? ls : (uint32_t)lo + 1;
    pos = ss;
    const int len = (int)(ee - ss);
    if (len < k || len < k + 2 * j + 1) continue;
    e.st[S2_UNAMBIG]++;
    // getValidReads: maximal runs of >= k Bloom-positive k-mers
    const uint32_t qe = ss + (uint32_t)(len - k + 1);
    uint32_t q = ss;
    while (q < qe) {
      const uint32_t a = s2_next_bit<true>(FpView<E, FP_V>{e}, q, qe);
      if (a >= qe) break;
      const uint32_t b = s2_next_bit<false>(FpView<E, FP_V>{e}, a, qe);
      if ((int)(b - a) >= k) {
        s2_subread(e, L, ls, n_pos, a, (int)(b - a) + k - 1);
        if (e.aborted()) return;
        e.st[S2_NOERR]++;
      }
      q = b;
    }
  }
}

// ---- "would this record change anything a later record can see?" --------------------------------
// A record is QUIET against the current table if its walk creates no junction and raises no stored
// distance: all it does is count coverage and set link flags, which no walk ever reads and which
// commute.  Quiet records therefore commute with each other; only the others ("writers") need the
// exclusive reservations of the schedule.  DryEnv runs the same walk read-only and stops at the first
// visible change.
template <class E>
struct DryEnv {
  const E& e;
  int k, j, spacer;
  bool pairs, want_ext;
  unsigned st[S2_COUNTERS];  // discarded
  bool quiet;
  FHD explicit DryEnv(const E& env) : e(env), k(env.k), j(env.j), spacer(env.spacer), pairs(false), want_ext(false), quiet(true) {
    for (int i = 0; i < S2_COUNTERS; i++) st[i] = 0;
  }
  FHD uint32_t inval_word(uint32_t w) const { return e.inval_word(w); }
  FHD uint32_t fp_word(int p, uint32_t w) const { return e.fp_word(p, w); }
  FHD uint32_t packed_word(uint32_t w) const { return e.packed_word(w); }
  FHD int find(uint64_t key) const { return e.find(key); }
  FHD int insert(uint64_t key, bool* created) {  // a key that is not there yet would be created: a visible change
    *created = false;
    const int s = e.find(key);
    if (s < 0) { quiet = false; return 0; }
    return s;
  }
  FHD void stamp(int) {}
  FHD void add_cov(int, int) {}
  FHD void link(int, int) {}
  FHD void update(int slot, int idx, int length) {
    if (((uint32_t)length & 0xffu) > e.dist_peek(slot, idx)) quiet = false;
  }
  FHD uint32_t dist_now(int slot, int idx) const { return e.dist_peek(slot, idx); }  // nothing of ours to be ordered after
  FHD void spf_pair(uint64_t, uint64_t) {}
  FHD void ext_push(uint64_t) {}
  FHD bool aborted() const { return !quiet; }
};
template <class E>
FHD bool s2_is_quiet(const E& e, LineState& L, uint32_t ls, uint32_t le) {
  DryEnv<E> d(e);
  s2_line(d, L, ls, le);
  L.pstale = 0;  // the dry walk marked what it touched; the real one starts over
  return d.quiet;
}

}  // namespace faucet
