// Pass 2, order-free part: everything ReadScanner asks the Bloom filter, evaluated for EVERY k-mer
// position at once.  One flag byte per byte offset p that starts a k-mer:
//
//   bit 0      V    bloo2.oldContains(canon(kmer@p))              getValidReads, src/ReadScanner.cpp:240
//   bit 1      JF   testForJunction at half-step (p, FORWARD)     src/ReadScanner.cpp:36-56
//   bit 2      JB   testForJunction at half-step (p, BACKWARD)
//   bits 3-4   CF   how many alternates were j-checked at (p, FORWARD)  (NbJCheckKmer, :46)
//   bits 5-6   CB   same for (p, BACKWARD)
//
// testForJunction is a pure function of (bloo2, oriented k-mer, real next nucleotide); the stitch
// (which half-steps are actually visited, in stream order) consumes these flags and never touches
// the Bloom filter again.
//
// Work shape.  A position with V set has 6 alternate extensions (3 per direction); ~w of them pass
// the first Bloom probe, ~w^n pass all n, and only those need the depth-j check (4^j more queries).
// Doing that per lane leaves ~6 of 32 lanes busy (measured), so the warp works in stages and
// re-packs the survivors of each stage densely over its lanes through shared memory:
//   stage 0  lane = position: k-mer, canonical form, V
//   stage 1  lane = position, 6 rounds: alternate -> canonical -> hash0 -> first probe; survivors queued
//   stage 2  lane = queued survivor: hash1, remaining probes; full members queued as j-check candidates
//   stage 3  lane = (candidate, next nucleotide) for j = 1, lane = candidate (DFS) for j >= 2
//   stage 4  lane = position: fold the per-alternate results in the reference's order
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer.cuh"
#include "load.cuh"

namespace faucet {

constexpr int SCAN_THREADS = 256;
#ifndef FAUCET_SCAN_CTAS
#define FAUCET_SCAN_CTAS 4
#endif
constexpr int SCAN_CTAS_PER_SM = FAUCET_SCAN_CTAS;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_Q = 192;  // 32 positions x 6 alternates
constexpr int MAX_J = 4;     // JChecker scratch arrays hold 1000 k-mers => j <= 4 (utils/JChecker.cpp:93-94)

struct ScanArgs {
  const uint32_t* inval;
  const uint32_t* packed;
  uint32_t n_words;       // the words [w_begin, n_words) of the text (32 byte offsets each) are scanned
  uint32_t w_begin;
  const uint32_t* bloom;  // bloo2, plain reference layout viewed as little-endian u32 words
  uint32_t wmask;         // (tai - 1) >> 5: word index mask (log2_tai <= 37, checked by the session)
  int k, j, n_hash;
  unsigned long long* memo;  // scan_flags_memo_kernel: one word per Bloom member seen so far (see there)
  uint64_t memo_mask;        // entries - 1 (power of two, >= 2^20)
  int memo_qbits;            // 64 - log2(entries): bits of the hash kept in the entry
  unsigned long long* dbg;   // optional counters: [0] lanes that missed, [1] warp passes through the long path, [2] failed inserts
  uint8_t* flags;     // one byte per byte offset
};

// Word holding bit (h mod tai).  Hash values are carried UNMASKED: (h0 + i*h1) mod tai only needs the
// low log2_tai bits of the 64-bit sum, so masking happens once, on the 32-bit word index (one funnel
// shift + one and + one IMAD.WIDE for the address instead of 64-bit shift/mask/add chains).
__device__ __forceinline__ uint32_t bloom_word(const ScanArgs& a, uint64_t h) {
  return __ldg(a.bloom + (__funnelshift_r((uint32_t)h, (uint32_t)(h >> 32), 5) & a.wmask));
}
__device__ __forceinline__ bool bloom_bit(const ScanArgs& a, uint64_t h) {
  return (bloom_word(a, h) >> ((uint32_t)h & 31u)) & 1u;
}

// Bloom::oldContains -> contains(h0,h1), utils/Bloom.h:162-173,242-258 (early exit on a clear bit)
template <int NH>
__device__ __forceinline__ bool bloom_contains(const ScanArgs& a, uint64_t x, uint64_t xrc) {
  const int nh = NH ? NH : a.n_hash;
  uint64_t c = canon(x, xrc);
  uint64_t h = hash0(c);
  if (!bloom_bit(a, h)) return false;
  const uint64_t h1 = hash1(c);
#pragma unroll
  for (int i = 1; i < (NH ? NH : MAX_NHASH); i++) {
    if (i >= nh) break;
    h += h1;
    if (!bloom_bit(a, h)) return false;
  }
  return true;
}

// same answer with every probe in flight at once (no early exit): for k-mers that are mostly members
template <int NH>
__device__ __forceinline__ bool bloom_contains_all(const ScanArgs& a, uint64_t x, uint64_t xrc) {
  const int nh = NH ? NH : a.n_hash;
  const uint64_t c = canon(x, xrc);
  uint64_t h = hash0(c);
  const uint64_t h1 = hash1(c);
  uint32_t ok = 1u;
#pragma unroll
  for (int i = 0; i < (NH ? NH : MAX_NHASH); i++) {
    if (i >= nh) break;
    ok &= bloom_word(a, h) >> ((uint32_t)h & 31u);
    h += h1;
  }
  return ok & 1u;
}

// JChecker::jcheck(kmer_type), utils/JChecker.cpp:51-80: true iff some path of j Bloom-positive
// forward extensions leaves x.  (Depth-first with early exit; the BFS there computes the same bool.)
template <int NH>
__device__ bool jcheck(const ScanArgs& a, uint64_t x, uint64_t xrc, uint64_t mask) {
  if (a.j == 0) return true;
  uint64_t kx[MAX_J], kr[MAX_J];
  int nt[MAX_J];
  int d = 0;
  kx[0] = x; kr[0] = xrc; nt[0] = 0;
  while (d >= 0) {
    if (nt[d] == 4) { d--; continue; }
    uint32_t c = nt[d]++;
    uint64_t y = ext_fwd(kx[d], c, mask), yr = ext_rc(kr[d], c, a.k);
    if (!bloom_contains<NH>(a, y, yr)) continue;
    if (d + 1 == a.j) return true;
    d++;
    kx[d] = y; kr[d] = yr; nt[d] = 0;
  }
  return false;
}

struct ScanQueue {  // per warp
  unsigned long long canon[SCAN_Q];  // canonical form of the queued alternate
  unsigned long long h0[SCAN_Q];     // its first probe position
  uint8_t tag[SCAN_Q];               // bit 0: the canonical form IS the oriented k-mer
  uint8_t res[SCAN_Q];               // bit 0: Bloom member (counted as j-checked), bit 1: passed the j-check
  uint8_t cand[SCAN_Q];              // queue indices of the full members
};

template <int NH>
__global__ void __launch_bounds__(SCAN_THREADS, SCAN_CTAS_PER_SM) scan_flags_kernel(ScanArgs a) {
  __shared__ ScanQueue queues[SCAN_WARPS];
  ScanQueue& q = queues[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t warp = a.w_begin + ((blockIdx.x * SCAN_THREADS + threadIdx.x) >> 5);
  const uint32_t n_warps = (gridDim.x * SCAN_THREADS) >> 5;
  const int nh = NH ? NH : a.n_hash;
  const int k = a.k;
  const uint64_t kbits = k >= 32 ? 0xffffffffull : ((1ull << k) - 1ull);
  const uint64_t mask = kmer_mask(k);
  for (uint32_t w = warp; w < a.n_words; w += n_warps) {
    const uint32_t lo = __ldg(a.inval + w), hi = __ldg(a.inval + w + 1);
    const uint64_t win = inval_window(lo, hi, lane);
    const bool start_ok = (win & kbits) == 0;
    if (!__any_sync(0xffffffffu, start_ok)) {
      continue;
    }
    const uint32_t p = (w << 5) + lane;
    // ---- stage 0
    uint64_t fwd = 0, rc = 0;
    bool V = false;
    if (start_ok) {
      fwd = kmer_at(a.packed, p, k);
      rc = revcomp(fwd, k);
      V = bloom_contains_all<NH>(a, fwd, rc);
    }
    // FORWARD half-step needs read[p+k]; BACKWARD needs read[p-1] (utils/ReadKmer.cpp:107-114)
    const bool has_next = V && !((win >> k) & 1ull);
    const bool has_prev = V && (lane ? !((lo >> (lane - 1)) & 1u) : (w && !(__ldg(a.inval + w - 1) >> 31)));
    const uint32_t real_f = has_next ? code_at(a.packed, p + k) : 0u;
    const uint32_t real_b = has_prev ? nt_comp(code_at(a.packed, p - 1)) : 0u;
    // ---- stage 1: first probe of the six alternates (t = 0..2 FORWARD, 3..5 BACKWARD, nucleotide order).
    // All six canonical forms, hashes and probe loads are issued before the first one is consumed.
    uint32_t my_idx = 0, my_idx_hi = 0;  // queue index of alternate t, one byte each
    uint32_t survived = 0;               // bit t: alternate t is queued
    int n1 = 0;
    uint64_t cn[6], hh[6];
    uint32_t wd[6];
    uint32_t cn_is_y = 0;
#pragma unroll
    for (int t = 0; t < 6; t++) {
      const bool fdir = t < 3;
      const uint32_t real = fdir ? real_f : real_b;
      const uint32_t tt = fdir ? t : t - 3;
      const uint32_t c = tt < real ? tt : tt + 1;  // the tt-th nucleotide that is not the real extension
      const uint64_t y = ext_fwd(fdir ? fwd : rc, c, mask), yr = ext_rc(fdir ? rc : fwd, c, k);
      cn_is_y |= (y < yr ? 1u : 0u) << t;
      cn[t] = y < yr ? y : yr;
      hh[t] = hash0(cn[t]);
      wd[t] = (fdir ? has_next : has_prev) ? bloom_word(a, hh[t]) : 0u;
    }
#pragma unroll
    for (int t = 0; t < 6; t++) {
      const bool hit = (wd[t] >> ((uint32_t)hh[t] & 31u)) & 1u;
      const uint32_t b = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int e = n1 + __popc(b & lt_mask);
        q.canon[e] = cn[t]; q.h0[e] = hh[t]; q.tag[e] = (cn_is_y >> t) & 1u;
        survived |= 1u << t;
        if (t < 4) my_idx |= (uint32_t)e << (8 * t); else my_idx_hi |= (uint32_t)e << (8 * (t - 4));
      }
      n1 += __popc(b);
    }
    __syncwarp();
    // ---- stage 2: remaining probes of the survivors
    int n2 = 0;
    for (int base = 0; base < n1; base += 32) {
      const int e = base + lane;
      bool full = false;
      if (e < n1) {
        full = true;
        if (nh > 1) {
          const uint64_t h1 = hash1(q.canon[e]);
          uint64_t h = q.h0[e];
          uint32_t ok = 1u;
#pragma unroll
          for (int i = 1; i < (NH ? NH : MAX_NHASH); i++) {  // all remaining probes in flight together
            if (i >= nh) break;
            h += h1;
            ok &= bloom_word(a, h) >> ((uint32_t)h & 31u);
          }
          full = ok & 1u;
        }
        q.res[e] = full ? (a.j == 0 ? 3 : 1) : 0;
      }
      const uint32_t b = __ballot_sync(0xffffffffu, full);
      if (full) q.cand[n2 + __popc(b & lt_mask)] = (uint8_t)e;
      n2 += __popc(b);
    }
    __syncwarp();
    // ---- stage 3: depth-j check of the full members (JChecker::jcheck)
    if (a.j == 1) {
      for (int base = 0; base < 4 * n2; base += 32) {
        const int task = base + lane;
        if (task < 4 * n2) {
          const int e = q.cand[task >> 2];
          const uint32_t c = task & 3;
          const uint64_t cn = q.canon[e], cr = revcomp(cn, k);
          const bool is_y = q.tag[e] & 1;
          const uint64_t y = is_y ? cn : cr, yr = is_y ? cr : cn;
          if (bloom_contains<NH>(a, ext_fwd(y, c, mask), ext_rc(yr, c, k))) q.res[e] = 3;
        }
      }
    } else if (a.j > 1) {
      for (int base = 0; base < n2; base += 32) {
        const int ci = base + lane;
        if (ci < n2) {
          const int e = q.cand[ci];
          const uint64_t cn = q.canon[e], cr = revcomp(cn, k);
          const bool is_y = q.tag[e] & 1;
          if (jcheck<NH>(a, is_y ? cn : cr, is_y ? cr : cn, mask)) q.res[e] = 3;
        }
      }
    }
    __syncwarp();
    // ---- stage 4: testForJunction's early exit, in nucleotide order (src/ReadScanner.cpp:41-53)
    uint32_t f = 0;
    if (start_ok) {
      f = V ? 1u : 0u;
#pragma unroll
      for (int d = 0; d < 2; d++) {
        uint32_t cnt = 0, junc = 0;
#pragma unroll
        for (int tt = 0; tt < 3; tt++) {
          const int t = 3 * d + tt;
          if (!junc && ((survived >> t) & 1u)) {
            const uint32_t e = t < 4 ? (my_idx >> (8 * t)) & 0xffu : (my_idx_hi >> (8 * (t - 4))) & 0xffu;
            const uint32_t r = q.res[e];
            cnt += r & 1u;
            junc = (r >> 1) & 1u;
          }
        }
        f |= d == 0 ? (junc << 1) | (cnt << 3) : (junc << 2) | (cnt << 5);
      }
      a.flags[p] = (uint8_t)f;
    }
    __syncwarp();  // the queues are reused by the next word
  }
}

// ---- the same flags, memoised per k-mer ---------------------------------------------------------------
// Everything ReadScanner asks about a k-mer X that is in the filter -- which of its four forward
// extensions are Bloom members and which of those pass the depth-j check, for X and for its reverse
// complement -- is a pure function of (bloo2, X): 16 bits.  A read set covers every genome k-mer
// `coverage` times, so those bits are computed once per DISTINCT k-mer and looked up afterwards:
// one 16-byte probe of a table in HBM (key = canonical k-mer) instead of ~14 Bloom probes and ~11
// evaluations of oldHash.  testForJunction for the FORWARD / BACKWARD half-step of a position is then
// derived from the masks and the read's real neighbour bases, in the reference's nucleotide order.
// The table is sized from the filter (it is a cache: when a neighbourhood is full, the k-mer is simply
// recomputed), holds only members (a k-mer that fails contains() has no flags), and must be cleared
// whenever bloo2 changes.  One 8-byte word per k-mer, so that the table of an E. coli-sized input
// (64 MB) stays in L2: the canonical k-mer goes through a BIJECTIVE 64-bit mix; the top log2(entries)
// bits pick the home slot and the rest (<= 44 bits) is stored next to the probe displacement, so the
// word identifies the k-mer exactly:  [0 | disp:3 | quotient:44 | masks:16], all ones = empty.
// masks: bits 0-3 member / 4-7 j-check of the canonical form's extensions, 8-11 / 12-15 of its reverse
// complement's.  One word = one atomicCAS to publish, no torn reads.
constexpr unsigned long long MEMO_EMPTY = ~0ull;
constexpr int MEMO_PROBES = 8;
constexpr int SCAN_QM = 256;  // 32 positions x 8 extensions

struct ScanQueueM {  // per warp
  unsigned long long canon[SCAN_QM];
  unsigned long long h0[SCAN_QM];
  uint8_t tag[SCAN_QM];
  uint8_t res[SCAN_QM];
  uint8_t cand[SCAN_QM];
};

// testForJunction (src/ReadScanner.cpp:36-56) from the masks of the half-step's base k-mer: bit 1 / 2 style result
// {junction, number of alternates that were j-checked}
__device__ __forceinline__ uint32_t junction_from_masks(uint32_t member, uint32_t jc, uint32_t real, bool have) {
  if (!have) return 0u;
  uint32_t cnt = 0, junc = 0;
#pragma unroll
  for (uint32_t nt = 0; nt < 4; nt++) {
    if (nt == real || junc) continue;
    if ((member >> nt) & 1u) { cnt++; junc = (jc >> nt) & 1u; }
  }
  return junc | (cnt << 1);
}

template <int NH>
__global__ void __launch_bounds__(SCAN_THREADS, SCAN_CTAS_PER_SM) scan_flags_memo_kernel(ScanArgs a) {
  __shared__ ScanQueueM queues[SCAN_WARPS];
  __shared__ uint8_t jlut[1024];  // junction_from_masks for every (member, j-check, real nucleotide)
  for (int i = threadIdx.x; i < 1024; i += SCAN_THREADS)
    jlut[i] = (uint8_t)junction_from_masks((uint32_t)i & 15u, ((uint32_t)i >> 4) & 15u, (uint32_t)i >> 8, true);
  __syncthreads();
  ScanQueueM& q = queues[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t warp = a.w_begin + ((blockIdx.x * SCAN_THREADS + threadIdx.x) >> 5);
  const uint32_t n_warps = (gridDim.x * SCAN_THREADS) >> 5;
  const int nh = NH ? NH : a.n_hash;
  const int k = a.k;
  const uint64_t kbits = k >= 32 ? 0xffffffffull : ((1ull << k) - 1ull);
  const uint64_t mask = kmer_mask(k);
  // The validity words of the NEXT word of text are fetched one iteration ahead, so a word without any k-mer start
  // (header and quality lines: ~60 % of a FASTQ) is dismissed without touching the code plane, and for the others the
  // k-mer's code words and the neighbour bases are requested together: planes -> memo probe -> flags.
  uint32_t lo_n = 0xffffffffu, hi_n = 0xffffffffu, prev_n = 0xffffffffu;
  if (warp < a.n_words) {
    lo_n = __ldg(a.inval + warp); hi_n = __ldg(a.inval + warp + 1);
    if (warp) prev_n = __ldg(a.inval + warp - 1);
  }
  for (uint32_t w = warp; w < a.n_words; w += n_warps) {
    const uint32_t p = (w << 5) + lane;
    const uint32_t lo = lo_n, hi = hi_n, prev_word = prev_n;
    if (w + n_warps < a.n_words) {
      lo_n = __ldg(a.inval + w + n_warps); hi_n = __ldg(a.inval + w + n_warps + 1); prev_n = __ldg(a.inval + w + n_warps - 1);
    }
    const uint64_t win = inval_window(lo, hi, lane);
    const bool start_ok = (win & kbits) == 0;
    if (!__any_sync(0xffffffffu, start_ok)) {
      continue;
    }
    const uint64_t fwd_raw = kmer_at(a.packed, p, k);
    const uint32_t next_code = code_at(a.packed, p + k);
    const uint32_t prev_code = p ? code_at(a.packed, p - 1) : 0u;
    uint64_t fwd = 0, rc = 0, cn = 0;
    bool is_c = true;       // the forward k-mer is the canonical form
    uint32_t masks = 0;     // as stored: canonical form's in the low byte
    bool have_masks = false, V = false;
    uint64_t home = 0, quot = 0;
    if (start_ok) {
      fwd = fwd_raw;
      rc = revcomp(fwd, k);
      is_c = fwd <= rc;
      cn = is_c ? fwd : rc;
      // ---- memo lookup
      uint64_t h = cn * 0x9E3779B97F4A7C15ull;  // odd multiplier, xor-shift: both bijections; the top bits of the
      h ^= h >> 29;                             // product (the home slot) depend on every bit of the k-mer
      home = h >> a.memo_qbits;
      quot = h & ((1ull << a.memo_qbits) - 1ull);
#pragma unroll 1
      for (uint64_t t = 0; t < MEMO_PROBES; t++) {
        const unsigned long long e = __ldcg(a.memo + ((home + t) & a.memo_mask));
        if (e == MEMO_EMPTY) break;
        if ((e >> 16) == ((t << 44) | quot)) { masks = (uint32_t)(e & 0xffffu); have_masks = true; V = true; break; }
      }
    }
    const bool miss = start_ok && !have_masks;
    if (__any_sync(0xffffffffu, miss)) {
      if (a.dbg) {
        const uint32_t mm = __ballot_sync(0xffffffffu, miss);
        if (lane == 0) { atomicAdd(a.dbg, (unsigned long long)__popc(mm)); atomicAdd(a.dbg + 1, 1ull); }
      }
      // ---- the long way for the lanes that missed: V, then all 8 one-base extensions (4 of fwd, 4 of rc)
      bool Vm = false;
      if (miss) Vm = bloom_contains_all<NH>(a, fwd, rc);
      uint32_t my_idx[2] = {0, 0};  // queue index of extension t, one byte each
      uint32_t survived = 0;
      int n1 = 0;
#pragma unroll
      for (int half = 0; half < 2; half++) {  // four extensions at a time keeps the register budget of the plain kernel
        uint64_t cq[4], hh[4];
        uint32_t wd[4];
        uint32_t cq_is_y = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const uint64_t y = ext_fwd(half ? rc : fwd, (uint32_t)c, mask), yr = ext_rc(half ? fwd : rc, (uint32_t)c, k);
          cq_is_y |= (y < yr ? 1u : 0u) << c;
          cq[c] = y < yr ? y : yr;
          hh[c] = hash0(cq[c]);
          wd[c] = (miss && Vm) ? bloom_word(a, hh[c]) : 0u;
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const bool hit = (wd[c] >> ((uint32_t)hh[c] & 31u)) & 1u;
          const uint32_t b = __ballot_sync(0xffffffffu, hit);
          if (hit) {
            const int e = n1 + __popc(b & lt_mask);
            q.canon[e] = cq[c]; q.h0[e] = hh[c]; q.tag[e] = (cq_is_y >> c) & 1u;
            survived |= 1u << (4 * half + c);
            my_idx[half] |= (uint32_t)e << (8 * c);
          }
          n1 += __popc(b);
        }
      }
      __syncwarp();
      int n2 = 0;
      for (int base = 0; base < n1; base += 32) {  // remaining probes of the survivors
        const int e = base + lane;
        bool full = false;
        if (e < n1) {
          full = true;
          if (nh > 1) {
            const uint64_t h1 = hash1(q.canon[e]);
            uint64_t h = q.h0[e];
            uint32_t ok = 1u;
#pragma unroll
            for (int i = 1; i < (NH ? NH : MAX_NHASH); i++) {
              if (i >= nh) break;
              h += h1;
              ok &= bloom_word(a, h) >> ((uint32_t)h & 31u);
            }
            full = ok & 1u;
          }
          q.res[e] = full ? (a.j == 0 ? 3 : 1) : 0;
        }
        const uint32_t b = __ballot_sync(0xffffffffu, full);
        if (full) q.cand[n2 + __popc(b & lt_mask)] = (uint8_t)e;
        n2 += __popc(b);
      }
      __syncwarp();
      if (a.j == 1) {  // depth-j check of the members (JChecker::jcheck)
        for (int base = 0; base < 4 * n2; base += 32) {
          const int task = base + lane;
          if (task < 4 * n2) {
            const int e = q.cand[task >> 2];
            const uint32_t c = task & 3;
            const uint64_t cc = q.canon[e], cr = revcomp(cc, k);
            const bool is_y = q.tag[e] & 1;
            const uint64_t y = is_y ? cc : cr, yr = is_y ? cr : cc;
            if (bloom_contains<NH>(a, ext_fwd(y, c, mask), ext_rc(yr, c, k))) q.res[e] = 3;
          }
        }
      } else if (a.j > 1) {
        for (int base = 0; base < n2; base += 32) {
          const int ci = base + lane;
          if (ci < n2) {
            const int e = q.cand[ci];
            const uint64_t cc = q.canon[e], cr = revcomp(cc, k);
            const bool is_y = q.tag[e] & 1;
            if (jcheck<NH>(a, is_y ? cc : cr, is_y ? cr : cc, mask)) q.res[e] = 3;
          }
        }
      }
      __syncwarp();
      if (miss) {
        V = Vm;
        if (Vm) {
          uint32_t mf = 0, mb = 0;  // {member nibble, j-check nibble} of fwd's and of rc's extensions
#pragma unroll
          for (int t = 0; t < 8; t++)
            if ((survived >> t) & 1u) {
              const uint32_t r = q.res[(my_idx[t >> 2] >> (8 * (t & 3))) & 0xffu];
              const uint32_t bits = (r & 1u) | ((r >> 1) & 1u) << 4;
              if (t < 4) mf |= bits << t; else mb |= bits << (t - 4);
            }
          masks = is_c ? (mf | (mb << 8)) : (mb | (mf << 8));
          have_masks = true;
          // publish (a cache: give up quietly when the neighbourhood is full or somebody else was first)
#pragma unroll 1
          for (uint64_t t = 0; t < MEMO_PROBES; t++) {
            const unsigned long long word = (((t << 44) | quot) << 16) | masks;
            const unsigned long long old = atomicCAS(a.memo + ((home + t) & a.memo_mask), MEMO_EMPTY, word);
            if (old == MEMO_EMPTY || (old >> 16) == (word >> 16)) break;
            if (a.dbg && t == MEMO_PROBES - 1) atomicAdd(a.dbg + 2, 1ull);
          }
        }
      }
      __syncwarp();  // the queues are reused by the next word
    }
    // ---- flags from the masks and the read's real neighbours
    uint32_t f = 0;
    if (start_ok && V) {
      const uint32_t mf = is_c ? (masks & 0xffu) : ((masks >> 8) & 0xffu), mb = is_c ? ((masks >> 8) & 0xffu) : (masks & 0xffu);
      // FORWARD half-step needs read[p+k]; BACKWARD needs read[p-1] (utils/ReadKmer.cpp:107-114)
      const bool has_next = !((win >> k) & 1ull);
      const bool has_prev = lane ? !((lo >> (lane - 1)) & 1u) : !(prev_word >> 31);
      const uint32_t real_f = has_next ? next_code : 0u;
      const uint32_t real_b = has_prev ? nt_comp(prev_code) : 0u;
      const uint32_t jf = has_next ? jlut[mf | (real_f << 8)] : 0u;
      const uint32_t jb = has_prev ? jlut[mb | (real_b << 8)] : 0u;
      f = 1u | ((jf & 1u) << 1) | ((jb & 1u) << 2) | ((jf >> 1) << 3) | ((jb >> 1) << 5);
    }
    if (start_ok) a.flags[p] = (uint8_t)f;
  }
}

// Batched form of JunctionMap::getValidJExtension (utils/JunctionMap.cpp:474-490), the Bloom query the contig build
// repeats at every step of findNeighbor (:231-462): for each oriented k-mer, which of its four forward extensions are
// Bloom members (low nibble) and which of those pass the depth-j check (high nibble).
template <int NH>
__global__ void ext_masks_kernel(ScanArgs a, const unsigned long long* __restrict__ kmers, unsigned long long n,
                                 uint8_t* __restrict__ out) {
  const uint64_t mask = kmer_mask(a.k);
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint64_t x = kmers[i] & mask, xrc = revcomp(x, a.k);
    uint32_t m = 0;
    for (uint32_t nt = 0; nt < 4; nt++) {
      const uint64_t y = ext_fwd(x, nt, mask), yrc = ext_rc(xrc, nt, a.k);
      if (bloom_contains<NH>(a, y, yrc)) {
        m |= 1u << nt;
        if (jcheck<NH>(a, y, yrc, mask)) m |= 16u << nt;
      }
    }
    out[i] = (uint8_t)m;
  }
}

}  // namespace faucet
