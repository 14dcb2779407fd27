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
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer.cuh"
#include "load.cuh"

namespace faucet {

constexpr int SCAN_THREADS = 256;
constexpr int MAX_J = 4;  // JChecker scratch arrays hold 1000 k-mers => j <= 4 (utils/JChecker.cpp:93-94)

struct ScanArgs {
  const uint32_t* inval;
  const uint32_t* packed;
  uint32_t n_words;
  const uint32_t* bloom;  // bloo2, plain reference layout viewed as little-endian u32 words
  uint64_t tai_mask;
  int k, j, n_hash;
  uint8_t* flags;
};

// Bloom::oldContains -> contains(h0,h1), utils/Bloom.h:162-173,242-258 (early exit on a clear bit)
template <int NH>
__device__ __forceinline__ bool bloom_contains(const ScanArgs& a, uint64_t x, uint64_t xrc) {
  const int nh = NH ? NH : a.n_hash;
  uint64_t c = canon(x, xrc);
  uint64_t h = hash0(c) & a.tai_mask;
  uint64_t h1 = hash1(c) & a.tai_mask;
#pragma unroll
  for (int i = 0; i < (NH ? NH : MAX_NHASH); i++) {
    if (i >= nh) break;
    if (!((__ldg(a.bloom + (h >> 5)) >> (h & 31)) & 1u)) return false;
    h = (h + h1) & a.tai_mask;
  }
  return true;
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

// testForJunction for the cursor whose oriented k-mer is `base` (revcomp `base_rc`) and whose real
// next nucleotide is `real`; returns bit0 = junction, bits1-2 = alternates that reached the j-check
template <int NH>
__device__ __forceinline__ uint32_t test_for_junction(const ScanArgs& a, uint64_t base, uint64_t base_rc,
                                                      uint32_t real, uint64_t mask) {
  uint32_t cnt = 0;
#pragma unroll
  for (uint32_t c = 0; c < 4; c++) {
    if (c == real) continue;
    uint64_t y = ext_fwd(base, c, mask), yr = ext_rc(base_rc, c, a.k);
    if (bloom_contains<NH>(a, y, yr)) {
      cnt++;
      if (jcheck<NH>(a, y, yr, mask)) return 1u | (cnt << 1);
    }
  }
  return cnt << 1;
}

template <int NH>
__global__ void __launch_bounds__(SCAN_THREADS) scan_flags_kernel(ScanArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * SCAN_THREADS + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * SCAN_THREADS) >> 5;
  const uint64_t kbits = a.k >= 32 ? 0xffffffffull : ((1ull << a.k) - 1ull);
  const uint64_t mask = kmer_mask(a.k);
  for (uint32_t w = warp; w < a.n_words; w += n_warps) {
    uint32_t lo = __ldg(a.inval + w), hi = __ldg(a.inval + w + 1);
    uint64_t win = inval_window(lo, hi, lane);
    bool start_ok = (win & kbits) == 0;
    if (!__any_sync(0xffffffffu, start_ok)) continue;
    if (!start_ok) continue;
    const uint32_t p = (w << 5) + lane;
    uint64_t fwd = kmer_at(a.packed, p, a.k);
    uint64_t rc = revcomp(fwd, a.k);
    uint32_t f = 0;
    if (bloom_contains<NH>(a, fwd, rc)) {
      f = 1;
      // FORWARD half-step needs read[p+k]; BACKWARD needs read[p-1] (utils/ReadKmer.cpp:107-114)
      bool has_next = !((win >> a.k) & 1ull);
      bool has_prev = lane ? !((lo >> (lane - 1)) & 1u) : (w && !(__ldg(a.inval + w - 1) >> 31));
      if (has_next) {
        uint32_t r = test_for_junction<NH>(a, fwd, rc, code_at(a.packed, p + a.k), mask);
        f |= (r & 1u) << 1 | (r >> 1) << 3;
      }
      if (has_prev) {
        uint32_t r = test_for_junction<NH>(a, rc, fwd, nt_comp(code_at(a.packed, p - 1)), mask);
        f |= (r & 1u) << 2 | (r >> 1) << 5;
      }
    }
    a.flags[p] = (uint8_t)f;
  }
}

}  // namespace faucet
