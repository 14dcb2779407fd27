// Pass 2, stream-order part ("stitch"), on the GPU.
//
// ReadScanner::scanReads / scanInputRead / scan_forward / find_next_junction / add_fake_junction
// (src/ReadScanner.cpp:61-359) mutate one JunctionMap in stream order: whether a half-step is a
// junction depends on the junctions earlier reads created (isJunction lookup, :67) and how far the
// cursor then skips depends on the dist[] earlier reads stored (:189-192).  Everything the scan asks
// the Bloom filter is already in the flag bytes of scan_flags_kernel, so what is left is this
// bookkeeping -- and it must come out exactly as if the records had been processed one by one.
//
// Schedule ("windowed deterministic reservations"): a record can only read or
// write junction keys that are k-mers (either orientation) of its own sequence line.  Rounds:
//   phase 1  every record of the window (the deferred ones + the next W new ones, so every unexecuted
//            record with a smaller index is inside it) reserves its keys with atomicMin(record index);
//   phase 2  a record executes iff it holds ALL its reservations, i.e. no earlier unexecuted record
//            shares a key with it; otherwise it is deferred to the next round.
// Two records that execute in the same round share no key, and an executing record shares no key with
// any earlier unexecuted one, so each record observes exactly the sequential state.  The earliest
// record of a window always executes.
//
// Reservations are taken on MINIMIZERS instead of k-mers: key(X) = min over the s-mers inside X of
// h(canonical s-mer), s = min(k,16).  Two lines that share a canonical k-mer share that value, so
// sharing is still detected (conservatively), with ~2/(k-s+2) reservations per k-mer instead of one.
// The reservation array is a plain u32 table indexed by the minimizer hash (collisions only defer).
//
// The junction map itself is an open-addressing table in HBM (key = oriented k-mer, ReadKmer::getKmer;
// 16-byte record = Junction's dist/cov/linked; 8-byte creation stamp = (record index, n-th creation in
// that record)).  Sorting by stamp gives the reference's creation order.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer.cuh"
#include "load.cuh"

namespace faucet {

namespace cg = cooperative_groups;

constexpr int STITCH_THREADS = 256;
constexpr int STITCH_WARPS = STITCH_THREADS / 32;
constexpr unsigned long long KEY_EMPTY = ~0ull;
constexpr uint32_t RES_FREE = 0xffffffffu;
constexpr int EXT_STAGE = 32;  // staged real-extension k-mers per warp before a chunk is flushed

enum { ST_DONE = 0, ST_GROW_TABLE = 1, ST_DRAIN_EXT = 2 };
enum { SS_JCHECK = 0, SS_NOJUNC, SS_PROCESSED, SS_SKIPPED, SS_NOERR, SS_UNAMBIG, SS_ROUNDS, SS_DEFERRED, SS_COUNT };

struct StitchState {             // device-resident; survives kernel launches and batches
  unsigned long long n_entries;  // occupied slots of the junction table
  unsigned long long stats[SS_COUNT];
  unsigned long long ext_used;   // u64 words used in the ext buffer
  unsigned long long need[2];    // upper bound of junction events of the records reserved this round
  unsigned int next;             // next new record of the batch
  unsigned int nd[2];            // deferred-record counts, double-buffered by round parity
  unsigned int W;                // window size (adapts to the deferral rate)
  unsigned int round;
  unsigned int status;
  unsigned int special;          // the key equal to KEY_EMPTY (k = 32, all 'G') is present
};

struct StitchArgs {
  const uint32_t* inval;
  const uint32_t* packed;
  const uint8_t* flags;
  const uint32_t* seq_start;
  const uint32_t* seq_end;
  uint32_t n_recs;               // records in this batch
  unsigned long long rec_base;   // global index of record 0 of this batch
  int k, j, spacer;
  int no_cleaning, paired;
  unsigned long long* keys;      // cap + 1 entries (the last one is the home of the KEY_EMPTY k-mer)
  uint4* recs;
  unsigned long long* stamps;
  unsigned long long cap;        // power of two
  uint32_t* res;                 // reservation table
  uint32_t res_mask;
  uint32_t* deferred[2];         // w_max entries each
  StitchState* st;
  uint32_t* spf;                 // short pair filter on the device (NULL: none)
  unsigned long long spf_mask;
  int spf_nh;
  unsigned long long* ext;       // real-extension chunks for the host-side long pair filter (NULL: none)
  unsigned long long ext_cap;
  uint32_t w_min, w_max;
};

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// h(canonical s-mer starting at byte offset q), s <= 16
__device__ __forceinline__ uint32_t smer_hash(const uint32_t* __restrict__ packed, uint32_t q, int s) {
  uint32_t w0 = __ldg(packed + (q >> 4)), w1 = __ldg(packed + (q >> 4) + 1);
  uint32_t x = __funnelshift_l(w1, w0, 2 * (q & 15)) >> (32 - 2 * s);
  uint32_t r = __brev(x << (32 - 2 * s));                       // reversed bit order, low-aligned
  r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);      // un-swap inside the 2-bit groups
  r ^= (s == 16 ? 0xaaaaaaaau : (0xaaaaaaaau & ((1u << (2 * s)) - 1u)));
  return mix32(x < r ? x : r);
}

// One warp walks the minimizers of line [ls, ls+len) and calls f(slot) once per run of equal values.
// MODE 0: reserve, 1: check (returns false on a foreign reservation), 2: release own reservations.
template <int MODE>
__device__ bool line_reservations(const StitchArgs& a, uint32_t ls, uint32_t len, uint32_t rec, int lane) {
  const int k = a.k, s = k < 16 ? k : 16, w = k - s + 1;
  if (len < (uint32_t)k) return true;
  const uint32_t nk = len - k + 1, ns = len - s + 1;
  bool ok = true;
  uint32_t g0 = lane < (int)ns ? smer_hash(a.packed, ls + lane, s) : 0xffffffffu;
  uint32_t prev_last = 0;  // minimizer of the last position of the previous chunk
  for (uint32_t base = 0; base < nk; base += 32) {
    uint32_t q1 = base + 32 + lane;
    uint32_t g1 = q1 < ns ? smer_hash(a.packed, ls + q1, s) : 0xffffffffu;
    uint32_t m = 0xffffffffu;
    for (int t = 0; t < w; t++) {
      int src = (lane + t) & 31;
      uint32_t v0 = __shfl_sync(0xffffffffu, g0, src), v1 = __shfl_sync(0xffffffffu, g1, src);
      uint32_t v = lane + t < 32 ? v0 : v1;
      m = v < m ? v : m;
    }
    uint32_t left = __shfl_up_sync(0xffffffffu, m, 1);
    if (lane == 0) left = prev_last;
    bool active = base + lane < nk && (base + lane == 0 || m != left);
    if (active) {
      uint32_t* slot = a.res + (m & a.res_mask);
      if (MODE == 0) atomicMin(slot, rec);
      if (MODE == 1 && __ldcg(slot) != rec) ok = false;
      if (MODE == 2 && __ldcg(slot) == rec) __stcg(slot, RES_FREE);
    }
    prev_last = __shfl_sync(0xffffffffu, m, 31);
    g0 = g1;
  }
  return MODE == 1 ? __all_sync(0xffffffffu, ok) : true;
}

// ---- junction table -----------------------------------------------------------------------------
__device__ __forceinline__ long long tbl_find(const StitchArgs& a, uint64_t key) {
  if (key == KEY_EMPTY) return __ldcg(&a.st->special) ? (long long)a.cap : -1;
  uint64_t h = mix64(key) & (a.cap - 1);
  while (true) {
    unsigned long long kk = __ldcg(a.keys + h);
    if (kk == key) return (long long)h;
    if (kk == KEY_EMPTY) return -1;
    h = (h + 1) & (a.cap - 1);
  }
}
// one thread; returns the slot and whether the key was created (JunctionMap::createJunction, zeroed record)
__device__ __forceinline__ long long tbl_insert(const StitchArgs& a, uint64_t key, bool* created) {
  if (key == KEY_EMPTY) {
    *created = atomicExch(&a.st->special, 1u) == 0u;
    if (*created) { a.keys[a.cap] = key; atomicAdd(&a.st->n_entries, 1ull); }
    return (long long)a.cap;
  }
  uint64_t h = mix64(key) & (a.cap - 1);
  while (true) {
    unsigned long long old = atomicCAS(a.keys + h, KEY_EMPTY, (unsigned long long)key);
    if (old == KEY_EMPTY) { *created = true; atomicAdd(&a.st->n_entries, 1ull); return (long long)h; }
    if (old == key) { *created = false; return (long long)h; }
    h = (h + 1) & (a.cap - 1);
  }
}

// 16-byte record image: bytes 0-4 dist, 5-8 cov, 9-13 linked (utils/Junction.h:12-19)
struct RecImg {
  uint32_t w[4];
  __device__ __forceinline__ uint32_t get(int b) const { return (w[b >> 2] >> (8 * (b & 3))) & 0xffu; }
  __device__ __forceinline__ void set(int b, uint32_t v) {
    w[b >> 2] = (w[b >> 2] & ~(0xffu << (8 * (b & 3)))) | (v << (8 * (b & 3)));
  }
  // Junction::update: dist = max(dist, (unsigned char)length)   (utils/Junction.cpp:69-71 + u8 narrowing at the call)
  __device__ __forceinline__ void update(int idx, int length) {
    uint32_t l = (uint32_t)length & 0xffu;
    if (l > get(idx)) set(idx, l);
  }
  __device__ __forceinline__ void add_cov(int nt) {  // Junction::addCoverage, saturating (utils/Junction.cpp:59-67)
    uint32_t c = get(5 + nt);
    if (c != 255) set(5 + nt, c + 1);
  }
  __device__ __forceinline__ void link(int idx) { set(9 + idx, 1); }
};
__device__ __forceinline__ RecImg rec_load(const StitchArgs& a, long long slot) {
  uint4 v = __ldcg(a.recs + slot);
  RecImg r; r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
  return r;
}
__device__ __forceinline__ void rec_store(const StitchArgs& a, long long slot, const RecImg& r) {
  __stcg(a.recs + slot, make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]));
}

// Bloom::addPair on the device copy of the short pair filter (utils/Bloom.cpp:127-140); adds commute
__device__ void spf_add_pair(const StitchArgs& a, uint64_t k1, uint64_t k2) {
  uint64_t e1 = canon(k1, revcomp(k1, a.k)), e2 = canon(k2, revcomp(k2, a.k));
  uint64_t h = hash0(e1 < e2 ? e1 : e2) & a.spf_mask, h1 = hash1(e1 < e2 ? e2 : e1) & a.spf_mask;
  for (int i = 0; i < a.spf_nh; i++, h += h1) {
    h &= a.spf_mask;
    atomicOr(a.spf + (h >> 5), 1u << (h & 31));
  }
}

struct WarpCtx {
  unsigned long long st[SS_COUNT];  // lane 0 only
  unsigned long long stamp;         // next creation stamp of the current record
  unsigned long long* stage;        // shared staging area of this warp (EXT_STAGE entries)
  uint32_t n_stage, part, rec;
};

__device__ void ext_flush(const StitchArgs& a, WarpCtx& c, int lane) {
  // chunk = header {record:32 | part:16 | count:16} + count real-extension k-mers
  uint32_t n = c.n_stage;
  unsigned long long off = 0;
  if (lane == 0) off = atomicAdd(&a.st->ext_used, (unsigned long long)n + 1);
  off = __shfl_sync(0xffffffffu, off, 0);
  if (lane == 0) a.ext[off] = ((unsigned long long)c.rec << 32) | ((unsigned long long)(c.part & 0xffffu) << 16) | n;
  __syncwarp();
  if (lane < (int)n) a.ext[off + 1 + lane] = c.stage[lane];
  __syncwarp();
  c.n_stage = 0;
  c.part++;
}

// scan_forward (src/ReadScanner.cpp:112-231) on the valid sub-read at byte offset s0, `len` bases
__device__ void scan_forward(const StitchArgs& a, WarpCtx& c, uint32_t s0, int len, int lane) {
  const int k = a.k, j = a.j;
  const uint64_t mask = kmer_mask(k);
  const int tested_end = 2 * len - 2 * k + 1 - 2 * j;  // distToEnd > 2j  <=>  tp < tested_end
  int tp = 2 * j + 1, last_junc_pos = 0;
  bool have_last = false, have_fb = false, have_lf = false;
  int last_tp = 0, last_fwd_idx = 0, rev_pos = 0, for_pos = 0;
  long long last_slot = -1;
  uint64_t fb_ext = 0, lf_ext = 0, v_prev1 = 0, v_prev2 = 0;  // v[n-1], v[n-2] of this sub-read's result list
  uint32_t n_out = 0;
  const bool pairs = !a.no_cleaning && a.spf != nullptr;
  const bool want_ext = a.ext != nullptr;

  auto push_out = [&](uint64_t real_ext) {
    if (pairs && n_out >= 2 && lane == 0) spf_add_pair(a, v_prev2, real_ext);  // (v[i], v[i+2]) once the list has > 2 entries
    // the first such pair is (v0, v2), issued when v2 arrives; a list that ends with exactly two
    // entries is handled after the loop (:208-218)
    v_prev2 = v_prev1; v_prev1 = real_ext; n_out++;
    if (want_ext) {
      if (lane == 0) c.stage[c.n_stage] = real_ext;
      c.n_stage++;
      __syncwarp();
      if (c.n_stage == EXT_STAGE - 1) ext_flush(a, c, lane);
    }
  };

  while (true) {
    // ---- find_next_junction (:61-86): 32 half-steps per warp iteration
    bool found = false;
    uint64_t key = 0;
    long long slot = -1;
    while (tp < tested_end) {
      const int t = tp + lane;
      const bool active = t < tested_end;
      bool known = false, spc = false, tst = false;
      uint32_t cnt = 0;
      uint64_t kk = 0;
      long long sl = -1;
      if (active) {
        const int pos = t >> 1, dir = t & 1;
        uint64_t fwd = kmer_at(a.packed, s0 + pos, k);
        kk = dir ? fwd : revcomp(fwd, k);
        sl = tbl_find(a, kk);
        known = sl >= 0;
        spc = t - last_junc_pos >= 2 * a.spacer - 1;
        uint32_t f = a.flags[s0 + pos];
        cnt = dir ? (f >> 3) & 3u : (f >> 5) & 3u;
        tst = dir ? (f & 2u) != 0 : (f & 4u) != 0;
      }
      const uint32_t am = __ballot_sync(0xffffffffu, active);
      const uint32_t hb = __ballot_sync(0xffffffffu, active && (known || spc || tst));
      const int hit = hb ? __ffs(hb) - 1 : 32;
      const uint32_t upto = hit < 32 ? (hit == 31 ? 0xffffffffu : ((2u << hit) - 1u)) : am;
      // NbJCheckKmer (:46): every half-step that reached testForJunction, the hit one included
      uint32_t jc = (active && ((upto >> lane) & 1u) && !known && !spc) ? cnt : 0u;
      jc = __reduce_add_sync(0xffffffffu, jc);
      if (lane == 0) c.st[SS_JCHECK] += jc;
      if (hb) {
        if (lane == 0) c.st[SS_PROCESSED] += hit;
        tp += hit;
        key = __shfl_sync(0xffffffffu, kk, hit);
        slot = __shfl_sync(0xffffffffu, sl, hit);
        found = true;
        break;
      }
      if (lane == 0) c.st[SS_PROCESSED] += __popc(am);
      tp += 32;
    }
    if (!found) break;
    // ---- the junction at half-step tp (:134-192)
    const int pos = tp >> 1, dir = tp & 1;
    const int real = dir ? (int)code_at(a.packed, s0 + pos + k) : (int)nt_comp(code_at(a.packed, s0 + pos - 1));
    const int fwd_idx = dir ? real : 4, back_idx = dir ? 4 : real;  // getExtensionIndex (utils/ReadKmer.cpp:95-100)
    int dist = 0;
    if (lane == 0) {
      if (slot < 0) {
        bool created;
        slot = tbl_insert(a, key, &created);
        if (created) a.stamps[slot] = c.stamp++;
      }
      RecImg r = rec_load(a, slot);
      r.add_cov(real);
      if (have_last) {  // directLinkJunctions (utils/JunctionMap.cpp:551-561)
        const int d = tp - last_tp;
        if (last_slot == slot) {
          r.update(last_fwd_idx, d); r.link(last_fwd_idx);
        } else {
          RecImg q = rec_load(a, last_slot);
          q.update(last_fwd_idx, d); q.link(last_fwd_idx);
          rec_store(a, last_slot, q);
        }
        r.update(back_idx, d); r.link(back_idx);
      } else {
        r.update(back_idx, tp - 2 * j);
      }
      rec_store(a, slot, r);
      dist = (int)r.get(fwd_idx);
      if (dist < 1) dist = 1;
      c.st[SS_PROCESSED] += 1;
      c.st[SS_SKIPPED] += (unsigned long long)(dist - 1);
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    dist = __shfl_sync(0xffffffffu, dist, 0);
    c.stamp = __shfl_sync(0xffffffffu, c.stamp, 0);
    __syncwarp();
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
    const uint64_t key = kmer_at(a.packed, s0 + pos, k);
    const int real = (int)code_at(a.packed, s0 + pos + k);
    if (lane == 0) {
      c.st[SS_NOJUNC]++;
      bool created;
      long long slot = tbl_insert(a, key, &created);
      if (created) a.stamps[slot] = c.stamp++;
      RecImg r = rec_load(a, slot);
      r.add_cov(real);
      const int mtp = 2 * pos + 1;
      r.update(4, mtp - 2 * j);
      r.update(real, (2 * len - mtp - 2 * k + 1) - 2 * j);
      rec_store(a, slot, r);
    }
    c.stamp = __shfl_sync(0xffffffffu, c.stamp, 0);
    __syncwarp();
    push_out(ext_fwd(key, (uint32_t)real, mask));
  } else if (lane == 0) {  // :205
    RecImg r = rec_load(a, last_slot);
    r.update(last_fwd_idx, (2 * len - last_tp - 2 * k + 1) - 2 * j);
    rec_store(a, last_slot, r);
  }
  __syncwarp();
  if (pairs && n_out == 2 && lane == 0) {  // :208-218
    if (have_fb && have_lf && !(rev_pos > for_pos)) spf_add_pair(a, fb_ext, lf_ext);
    if (have_fb != have_lf) spf_add_pair(a, v_prev2, v_prev1);
  }
}

// highest position in [s, pos) whose plane bit equals `want`, or -1; uniform across the warp
__device__ long long find_prev_bit(const uint32_t* __restrict__ plane, uint32_t s, uint32_t pos, bool want) {
  if (pos <= s) return -1;
  const uint32_t p = pos - 1, ws = s >> 5;
  uint32_t w = p >> 5;
  uint32_t word = __ldg(plane + w);
  if (!want) word = ~word;
  if ((p & 31) != 31) word &= (2u << (p & 31)) - 1u;
  while (true) {
    if (w == ws) word &= ~((1u << (s & 31)) - 1u);
    if (word) return ((long long)w << 5) + 31 - __clz(word);
    if (w == ws) return -1;
    w--;
    word = __ldg(plane + w);
    if (!want) word = ~word;
  }
}

// scanInputRead (:260-282) + getValidReads (:233-257) for the sequence line [ls, le)
__device__ void scan_line(const StitchArgs& a, WarpCtx& c, uint32_t ls, uint32_t le, int lane) {
  const int k = a.k, j = a.j;
  uint32_t pos = le;
  while (pos > ls) {  // getUnambiguousReads hands the segments over LAST first (utils/Kmer.cpp:64-80)
    long long hi = find_prev_bit(a.inval, ls, pos, false);
    if (hi < 0) break;
    const uint32_t ee = (uint32_t)hi + 1;
    long long lo = find_prev_bit(a.inval, ls, ee, true);
    const uint32_t ss = lo < 0 ? ls : (uint32_t)lo + 1;
    pos = ss;
    const int L = (int)(ee - ss);
    if (L < k || L < k + 2 * j + 1) continue;
    if (lane == 0) c.st[SS_UNAMBIG]++;
    // getValidReads: maximal runs of >= k Bloom-positive k-mers; npos is a virtual negative position
    const int npos = L - k + 1;
    int run_start = -1;
    for (int base = 0; base <= npos; base += 32) {
      const int i = base + lane;
      const bool v = i < npos && (a.flags[ss + i] & 1u);
      const uint32_t m = __ballot_sync(0xffffffffu, v);
      const int lanes = npos + 1 - base < 32 ? npos + 1 - base : 32;
      int bit = 0;
      while (bit < lanes) {
        if (run_start < 0) {
          uint32_t x = m >> bit;
          if (!x) break;
          bit += __ffs(x) - 1;
          run_start = base + bit;
        } else {
          uint32_t x = (~m) >> bit;
          if (!x) break;
          bit += __ffs(x) - 1;
          if (bit >= lanes) bit = lanes - 1;  // unreachable: bit `lanes-1` of the last chunk is clear
          const int run_len = base + bit - run_start;
          if (run_len >= k) {
            scan_forward(a, c, ss + run_start, run_len + k - 1, lane);
            if (lane == 0) c.st[SS_NOERR]++;
          }
          run_start = -1;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(STITCH_THREADS) stitch_kernel(StitchArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ unsigned long long stage[STITCH_WARPS][EXT_STAGE];
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * STITCH_THREADS + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * STITCH_THREADS) >> 5;
  StitchState* st = a.st;
  uint32_t next = __ldcg(&st->next), W = __ldcg(&st->W), round = __ldcg(&st->round);
  WarpCtx c;
  for (int i = 0; i < SS_COUNT; i++) c.st[i] = 0;
  c.stage = stage[threadIdx.x >> 5];
  c.n_stage = 0; c.part = 0; c.rec = 0; c.stamp = 0;
  uint32_t status = ST_DONE;

  while (true) {
    const int cur = round & 1, nxt = cur ^ 1;
    const uint32_t nd = __ldcg(&st->nd[cur]);
    const uint32_t room = a.n_recs - next;
    const uint32_t n_new = W > nd ? (W - nd < room ? W - nd : room) : 0u;
    const uint32_t n_win = nd + n_new;
    if (n_win == 0) break;
    // n_entries / ext_used only move in phase 2, so this snapshot is the same in every thread
    const unsigned long long entries0 = __ldcg(&st->n_entries), ext0 = __ldcg(&st->ext_used);
    if (gw == 0 && lane == 0) { __stcg(&st->nd[nxt], 0u); __stcg(&st->need[nxt], 0ull); }
    // ---- phase 1: reservations
    unsigned long long need = 0;
    for (uint32_t e = gw; e < n_win; e += n_warps) {
      const uint32_t rec = e < nd ? __ldcg(a.deferred[cur] + e) : next + (e - nd);
      const uint32_t ls = __ldg(a.seq_start + rec), le = __ldg(a.seq_end + rec);
      const uint32_t len = le > ls ? le - ls : 0u;
      line_reservations<0>(a, ls, len, rec, lane);
      need += 2ull * len + 2;
    }
    if (lane == 0 && need) atomicAdd(&st->need[cur], need);
    grid.sync();
    {
      const unsigned long long bound = __ldcg(&st->need[cur]);
      if (entries0 + bound > (a.cap / 4) * 3) { status = ST_GROW_TABLE; break; }
      if (a.ext && ext0 + 2 * bound + n_win > a.ext_cap) { status = ST_DRAIN_EXT; break; }
    }
    // ---- phase 2: execute or defer
    for (uint32_t e = gw; e < n_win; e += n_warps) {
      const uint32_t rec = e < nd ? __ldcg(a.deferred[cur] + e) : next + (e - nd);
      const uint32_t ls = __ldg(a.seq_start + rec), le = __ldg(a.seq_end + rec);
      const uint32_t len = le > ls ? le - ls : 0u;
      const bool mine = line_reservations<1>(a, ls, len, rec, lane);
      line_reservations<2>(a, ls, len, rec, lane);
      if (mine) {
        c.rec = rec; c.part = 0; c.n_stage = 0;
        c.stamp = (a.rec_base + rec) << 20;
        if (len) scan_line(a, c, ls, ls + len, lane);
        if (a.ext && c.n_stage) ext_flush(a, c, lane);
      } else if (lane == 0) {
        a.deferred[nxt][atomicAdd(&st->nd[nxt], 1u)] = rec;
        c.st[SS_DEFERRED]++;
      }
    }
    grid.sync();
    const uint32_t nd_next = __ldcg(&st->nd[nxt]);
    if (nd_next * 4 > n_win) W = W / 2 > a.w_min ? W / 2 : a.w_min;
    else if (nd_next * 10 < n_win && n_win >= W) W = W * 2 < a.w_max ? W * 2 : a.w_max;
    next += n_new;
    round++;
    if (gw == 0 && lane == 0) c.st[SS_ROUNDS]++;
  }
  if (lane == 0)
    for (int i = 0; i < SS_COUNT; i++)
      if (c.st[i]) atomicAdd(&st->stats[i], c.st[i]);
  if (gw == 0 && lane == 0) {
    __stcg(&st->next, next); __stcg(&st->W, W); __stcg(&st->round, round); __stcg(&st->status, status);
  }
}

// ---- table maintenance ---------------------------------------------------------------------------
__global__ void stitch_rehash_kernel(const unsigned long long* __restrict__ okeys, const uint4* __restrict__ orecs,
                                     const unsigned long long* __restrict__ ostamps, unsigned long long ocap,
                                     unsigned long long* keys, uint4* recs, unsigned long long* stamps,
                                     unsigned long long cap) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= ocap;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long key = okeys[i];
    if (i == ocap) {  // the home of the KEY_EMPTY k-mer moves to the new last slot
      keys[cap] = key; recs[cap] = orecs[i]; stamps[cap] = ostamps[i];
      continue;
    }
    if (key == KEY_EMPTY) continue;
    unsigned long long h = mix64(key) & (cap - 1);
    while (atomicCAS(keys + h, KEY_EMPTY, key) != KEY_EMPTY) h = (h + 1) & (cap - 1);
    recs[h] = orecs[i];
    stamps[h] = ostamps[i];
  }
}

struct JunctionOut {  // == faucet_junction_rec (include/faucet_gpu.h)
  unsigned long long kmer;
  uint32_t body[4];   // dist[5] cov[4] linked[5] pad[2]
  unsigned long long stamp;
};

__global__ void stitch_collect_kernel(const unsigned long long* __restrict__ keys, const uint4* __restrict__ recs,
                                      const unsigned long long* __restrict__ stamps, unsigned long long cap,
                                      unsigned int special, JunctionOut* __restrict__ out,
                                      unsigned long long* __restrict__ n_out) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= cap;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long key = keys[i];
    bool occ = i == cap ? special != 0 : key != KEY_EMPTY;
    uint32_t m = __ballot_sync(__activemask(), occ);
    if (!occ) continue;
    // one atomic per warp
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(n_out, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    unsigned long long o = base + __popc(m & ((1u << lane) - 1u));
    uint4 r = recs[i];
    JunctionOut jo;
    jo.kmer = key; jo.body[0] = r.x; jo.body[1] = r.y; jo.body[2] = r.z; jo.body[3] = r.w & 0xffffu; jo.stamp = stamps[i];
    out[o] = jo;
  }
}

}  // namespace faucet
