// Pass 1 ("Bloom load"): exact, order-free evaluation of load_two_filters (utils/Bloom.cpp:267-299).
//
// The reference is a sequential stream:   if bloo1.contains(x_t) bloo2.add(x_t) else bloo1.add(x_t).
// bloo1 only ever skips an add that would be a no-op, so bloo1(t) = OR of bits(x_u), u < t, and
//     contains1(t)  <=>  for every bit b of x_t :  F(b) < t,   F(b) = min{u : b in bits(x_u)}
// (SURVEY F3).  We keep F as a dense array of 32-bit stamps T[tai] and evaluate it per batch:
//
//   kernel A (all k-mers)   probe bloo1 as it stood before the batch.  All bits set => contained, OR
//                           the bits into bloo2 right away.  Otherwise atomicMin the stamp of every
//                           clear bit and flag the occurrence as pending.
//   kernel B (pending only) contained <=> every stamp < t.  OR into bloo2 if contained, into bloo1 if
//                           not.  B never READS bloo1, so there is no ordering hazard inside it.
//
// Stamps are monotone across batches (base + byte offset), so a stamp written by an earlier batch is
// automatically "< t".  The two filters live interleaved, one u64 = {bloo1 word, bloo2 word}: a probe
// of bloo1 and the follow-up update of bloo2 touch the same 32-byte sector.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer.cuh"

namespace faucet {

constexpr uint32_t STAMP_INF = 0xffffffffu;
constexpr int LOAD_THREADS = 256;
constexpr int LOAD_CTAS_PER_SM = 6;  // load_A is compiled for 6 resident CTAs per SM; its grid is exactly one wave
constexpr int MAX_NHASH = 10;  // NSEEDSBLOOM, utils/Bloom.h:40

struct LoadCounters {
  unsigned long long kmers;       // k-mer occurrences streamed
  unsigned long long segments;    // "Unambiguous reads" (utils/Bloom.cpp:287)
  unsigned long long pending;     // occurrences that needed kernel B
  unsigned long long fresh;       // occurrences NOT contained in bloo1 (went to bloo1)
  unsigned long long weight1, weight2;
};

struct LoadArgs {
  const uint32_t* inval;
  const uint32_t* packed;
  const uint32_t* skipA;
  uint32_t* pend;                 // 1 bit / byte offset
  uint32_t n_words;               // ceil(n / 32)
  uint32_t w_begin, w_end;        // words of this sub-batch (see faucet_session_load)
  unsigned long long* fused;      // tai/32 words: low half bloo1, high half bloo2
  uint32_t* stamps;               // T[tai]
  uint64_t tai_mask;
  uint32_t base;                  // stamp of byte offset 0 of this batch
  int k;
  int n_hash;
  LoadCounters* ctr;
  unsigned long long* memo;       // saturated k-mers (NULL: not used): see load_body_A
  uint64_t memo_mask;
  int memo_qbits;
  const uint8_t* text;            // complex-line kernels only
  const uint2* complex_list;
  uint32_t n_complex;
};

// k-mer starting at byte offset p (big-endian 2-bit stream, 16 bases per u32)
__device__ __forceinline__ uint64_t kmer_at(const uint32_t* __restrict__ packed, uint32_t p, int k) {
  uint32_t w = p >> 4, o = 2 * (p & 15);
  uint64_t hi = ((uint64_t)__ldg(packed + w) << 32) | __ldg(packed + w + 1);
  uint64_t lo = (uint64_t)__ldg(packed + w + 2) << 32;
  uint64_t x = o ? ((hi << o) | (lo >> (64 - o))) : hi;
  return x >> (64 - 2 * k);
}
__device__ __forceinline__ uint32_t code_at(const uint32_t* __restrict__ packed, uint32_t p) {
  return (__ldg(packed + (p >> 4)) >> (30 - 2 * (p & 15))) & 3u;
}
// window of validity bits starting at offset 32*w + lane (bit i <=> byte p+i is NOT a base)
__device__ __forceinline__ uint64_t inval_window(uint32_t lo, uint32_t hi, int lane) {
  return (((uint64_t)hi << 32) | lo) >> lane;
}

// A k-mer is SATURATED once every bit of it is set in bloo1 and in bloo2: whatever occurrence of it comes later (in
// any order: bits are only ever set) is contained in bloo1 and adds nothing to bloo2, i.e. it is a no-op of
// load_two_filters.  When the filters are too big for L2 (the session enables this from 2^29 bits on), saturated
// k-mers are remembered in a cache (the table layout of scan.cuh's memo: bijective mix, home slot + quotient, 8
// probes, one u64 per k-mer) and their later occurrences cost one 8-byte probe instead of two oldHash and n_hash
// probes of the filter in HBM.  (With an L2-resident filter the probes it saves are cheaper than the one it adds.)
constexpr int LMEMO_PROBES = 8;
__device__ __forceinline__ void lmemo_slot(const LoadArgs& a, uint64_t c, uint64_t* home, uint64_t* quot) {
  uint64_t h = c * 0x9E3779B97F4A7C15ull;
  h ^= h >> 29;
  *home = h >> a.memo_qbits;
  *quot = h & ((1ull << a.memo_qbits) - 1ull);
}

template <int NH, bool LM = false>  // LM: with the cache of saturated k-mers (a separate instantiation: the plain one keeps its registers)
__device__ __forceinline__ bool load_body_A(const LoadArgs& a, uint64_t fwd, uint32_t t) {
  const int nh = NH ? NH : a.n_hash;
  uint64_t rc = revcomp(fwd, a.k);
  uint64_t c = canon(fwd, rc);
  uint64_t mhome = 0, mquot = 0;
  if (LM) {
    lmemo_slot(a, c, &mhome, &mquot);
#pragma unroll 1
    for (uint64_t i = 0; i < LMEMO_PROBES; i++) {
      const unsigned long long e = __ldcg(a.memo + ((mhome + i) & a.memo_mask));
      if (e == ~0ull) break;
      if ((e >> 16) == ((i << 44) | mquot)) return false;  // saturated: nothing to do, not pending
    }
  }
  uint64_t h0 = hash0(c) & a.tai_mask, h1 = hash1(c) & a.tai_mask;
  uint64_t pos[NH ? NH : MAX_NHASH];
  unsigned long long wd[NH ? NH : MAX_NHASH];
#pragma unroll
  for (int i = 0; i < (NH ? NH : MAX_NHASH); i++)
    if (i < nh) {
      pos[i] = (h0 + (uint64_t)i * h1) & a.tai_mask;
      wd[i] = a.fused[(uint32_t)(pos[i] >> 5)];  // 32-bit word index (log2_tai <= 37): one IMAD.WIDE address
    }
  bool all1 = true;
#pragma unroll
  for (int i = 0; i < (NH ? NH : MAX_NHASH); i++)
    if (i < nh) all1 &= (bool)((wd[i] >> (pos[i] & 31)) & 1ull);
  if (all1) {
    // contained in bloo1 as of the batch start => Bloom::add on bloo2 (utils/Bloom.h:217-226)
    bool all2 = true;
#pragma unroll
    for (int i = 0; i < (NH ? NH : MAX_NHASH); i++)
      if (i < nh && !((wd[i] >> (32 + (pos[i] & 31))) & 1ull)) {
        all2 = false;
        atomicOr(reinterpret_cast<unsigned int*>(a.fused + (uint32_t)(pos[i] >> 5)) + 1, 1u << (pos[i] & 31));
      }
    if (LM && all2) {  // both filters already held every bit: remember the k-mer (a cache: give up when crowded)
#pragma unroll 1
      for (uint64_t i = 0; i < LMEMO_PROBES; i++) {
        const unsigned long long word = ((i << 44) | mquot) << 16;
        const unsigned long long old = atomicCAS(a.memo + ((mhome + i) & a.memo_mask), ~0ull, word);
        if (old == ~0ull || (old >> 16) == (word >> 16)) break;
      }
    }
    return false;
  }
#pragma unroll
  for (int i = 0; i < (NH ? NH : MAX_NHASH); i++)
    if (i < nh && !((wd[i] >> (pos[i] & 31)) & 1ull)) {
      if (a.stamps[pos[i]] > t) atomicMin(a.stamps + pos[i], t);
    }
  return true;
}

template <int NH>
__device__ __forceinline__ bool load_body_B(const LoadArgs& a, uint64_t fwd, uint32_t t) {
  const int nh = NH ? NH : a.n_hash;
  uint64_t rc = revcomp(fwd, a.k);
  uint64_t c = canon(fwd, rc);
  uint64_t h0 = hash0(c) & a.tai_mask, h1 = hash1(c) & a.tai_mask;
  uint64_t pos[NH ? NH : MAX_NHASH];
  uint32_t st[NH ? NH : MAX_NHASH];
#pragma unroll
  for (int i = 0; i < (NH ? NH : MAX_NHASH); i++)
    if (i < nh) {
      pos[i] = (h0 + (uint64_t)i * h1) & a.tai_mask;
      st[i] = a.stamps[pos[i]];
    }
  bool contained = true;
#pragma unroll
  for (int i = 0; i < (NH ? NH : MAX_NHASH); i++)
    if (i < nh) contained &= st[i] < t;
  const int half = contained ? 1 : 0;  // bloo2 if contained, bloo1 otherwise (utils/Bloom.cpp:293-298)
#pragma unroll
  for (int i = 0; i < (NH ? NH : MAX_NHASH); i++)
    if (i < nh) atomicOr(reinterpret_cast<unsigned int*>(a.fused + (uint32_t)(pos[i] >> 5)) + half, 1u << (pos[i] & 31));
  return contained;
}

template <int NH, bool LM = false>
__global__ void __launch_bounds__(LOAD_THREADS, 6) load_A_kernel(LoadArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * LOAD_THREADS + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * LOAD_THREADS) >> 5;
  const uint64_t kbits = a.k >= 32 ? 0xffffffffull : ((1ull << a.k) - 1ull);
  unsigned long long n_kmers = 0, n_segs = 0, n_pend = 0;
  for (uint32_t w = a.w_begin + warp; w < a.w_end; w += n_warps) {
    uint32_t lo = __ldg(a.inval + w), hi = __ldg(a.inval + w + 1);
    bool start_ok = (inval_window(lo, hi, lane) & kbits) == 0;
    if (!__any_sync(0xffffffffu, start_ok)) {
      if (lane == 0) a.pend[w] = 0;
      continue;
    }
    uint32_t sk = __ldg(a.skipA + w);
    bool valid = start_ok && !((sk >> lane) & 1u);
    uint32_t prev_inval = lane ? ((lo >> (lane - 1)) & 1u) : (w ? (__ldg(a.inval + w - 1) >> 31) : 1u);
    uint32_t vb = __ballot_sync(0xffffffffu, valid);
    uint32_t sb = __ballot_sync(0xffffffffu, valid && prev_inval);
    bool pending = false;
    if (valid) {
      uint32_t p = (w << 5) + lane;
      pending = load_body_A<NH, LM>(a, kmer_at(a.packed, p, a.k), a.base + p);
    }
    uint32_t pb = __ballot_sync(0xffffffffu, pending);
    if (lane == 0) {
      a.pend[w] = pb;
      n_kmers += __popc(vb);
      n_segs += __popc(sb);
      n_pend += __popc(pb);
    }
  }
  if (lane == 0 && n_kmers) {
    atomicAdd(&a.ctr->kmers, n_kmers);
    atomicAdd(&a.ctr->segments, n_segs);
    atomicAdd(&a.ctr->pending, n_pend);
  }
}

template <int NH>
__global__ void __launch_bounds__(LOAD_THREADS) load_B_kernel(LoadArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * LOAD_THREADS + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * LOAD_THREADS) >> 5;
  unsigned int n_fresh = 0;
  for (uint32_t w = a.w_begin + warp; w < a.w_end; w += n_warps) {
    uint32_t pb = __ldg(a.pend + w);
    if (!((pb >> lane) & 1u)) continue;
    uint32_t p = (w << 5) + lane;
    n_fresh += load_body_B<NH>(a, kmer_at(a.packed, p, a.k), a.base + p) ? 0u : 1u;
  }
  for (int o = 16; o; o >>= 1) n_fresh += __shfl_xor_sync(0xffffffffu, n_fresh, o);
  if (lane == 0 && n_fresh) atomicAdd(&a.ctr->fresh, (unsigned long long)n_fresh);
}

// Lines with several segments: getUnambiguousReads returns them LAST segment first
// (utils/Kmer.cpp:64-80), so segment [ss,ee) of line [s,e) is streamed at offset s + (e - ee).
// One warp per line; PHASE 0 = kernel A semantics, PHASE 1 = kernel B over every k-mer of the line.
template <int PHASE>
__global__ void __launch_bounds__(LOAD_THREADS) load_complex_kernel(LoadArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * LOAD_THREADS + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * LOAD_THREADS) >> 5;
  unsigned long long n_kmers = 0, n_segs = 0;
  for (uint32_t li = warp; li < a.n_complex; li += n_warps) {
    const uint32_t s = a.complex_list[li].x, e = a.complex_list[li].y;
    if ((s >> 5) < a.w_begin || (s >> 5) >= a.w_end) continue;  // a line belongs to the sub-batch of its first byte
    uint32_t q = s;
    while (q < e) {
      while (q < e && !nt_valid(a.text[q])) q++;
      uint32_t ss = q;
      while (q < e && nt_valid(a.text[q])) q++;
      uint32_t ee = q;
      if (ee - ss < (uint32_t)a.k) continue;
      n_segs++;
      n_kmers += ee - ss - a.k + 1;
      const uint32_t t0 = a.base + s + (e - ee);
      for (uint32_t p = ss + lane; p + a.k <= ee; p += 32) {
        uint64_t fwd = kmer_at(a.packed, p, a.k);
        if (PHASE == 0) load_body_A<0>(a, fwd, t0 + (p - ss));
        else if (!load_body_B<0>(a, fwd, t0 + (p - ss))) atomicAdd(&a.ctr->fresh, 1ull);
      }
    }
  }
  if (PHASE == 0 && lane == 0 && n_kmers) {
    atomicAdd(&a.ctr->kmers, n_kmers);
    atomicAdd(&a.ctr->segments, n_segs);
  }
}

// new stamp epoch: everything stamped so far is "before" anything to come
__global__ void stamps_epoch_kernel(uint32_t* __restrict__ stamps, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (stamps[i] != STAMP_INF) stamps[i] = 0;
}

// fused {bloo1,bloo2} words -> two plain bit arrays in the reference layout
// (bit h <-> byte h>>3, mask 1<<(h&7): utils/Bloom.h:44-53) + Bloom::weight() popcounts
__global__ void __launch_bounds__(256)
bloom_split_kernel(const unsigned long long* __restrict__ fused, uint64_t n_words, uint32_t* __restrict__ b1,
                   uint32_t* __restrict__ b2, LoadCounters* __restrict__ ctr) {
  unsigned long long w1 = 0, w2 = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long v = fused[i];
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    if (b1) b1[i] = lo;
    b2[i] = hi;
    w1 += __popc(lo);
    w2 += __popc(hi);
  }
  for (int o = 16; o; o >>= 1) {
    w1 += __shfl_xor_sync(0xffffffffu, w1, o);
    w2 += __shfl_xor_sync(0xffffffffu, w2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&ctr->weight1, w1);
    atomicAdd(&ctr->weight2, w2);
  }
}

}  // namespace faucet
