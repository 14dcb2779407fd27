// k-mer codec + Bloom hash shared by every kernel (and by the host-side stitch).
// Semantics follow the reference; the implementation is bit-parallel instead of table/loop driven.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FHD __host__ __device__ __forceinline__
#else
#define FHD inline
#endif

namespace faucet {

// NT2int, utils/Kmer.cpp:82-88: A=0 C=1 T=2 G=3
FHD uint32_t nt_code(uint8_t c) { return (c >> 1) & 3u; }
// isValidNuc, utils/Kmer.cpp:50-60 (upper-case ACGT only)
FHD bool nt_valid(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
// revcomp_int, utils/Kmer.cpp:90-93 == xor 2 in this code
FHD uint32_t nt_comp(uint32_t nt) { return nt ^ 2u; }

FHD uint64_t kmer_mask(int k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }

FHD uint64_t bitrev64(uint64_t x) {
#ifdef __CUDA_ARCH__
  return __brevll(x);
#else
  x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
  return __builtin_bswap64(x);
#endif
}

// revcomp(uint64,k), utils/Kmer.cpp:238-252 (byte LUT there): reverse the 2-bit groups, complement
FHD uint64_t revcomp(uint64_t x, int k) {
  uint64_t r = bitrev64(x);
  r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);  // un-swap bits inside groups
  r >>= (64 - 2 * k);
  return r ^ (0xAAAAAAAAAAAAAAAAull & kmer_mask(k));
}
FHD uint64_t canon(uint64_t fwd, uint64_t rc) { return fwd < rc ? fwd : rc; }  // get_canon, Kmer.cpp:531-533
// next_kmer(x, nt, FORWARD) / shift_kmer, utils/Kmer.cpp:410-425
FHD uint64_t ext_fwd(uint64_t x, uint32_t nt, uint64_t mask) { return ((x << 2) | nt) & mask; }
// revcomp of ext_fwd(x, nt): shift the revcomp right and put comp(nt) on top (DoubleKmer::forward, DoubleKmer.cpp:5-8)
FHD uint64_t ext_rc(uint64_t rc, uint32_t nt, int k) { return (rc >> 2) | ((uint64_t)nt_comp(nt) << (2 * k - 2)); }

// Bloom seeds: generate_hash_seed with user_seed 0, utils/Bloom.cpp:500-511 (rbase[0]*rbase[3], rbase[1]*rbase[4])
#define FAUCET_SEED0 0xffaa54ffe6e6e6e7ull
#define FAUCET_SEED1 0x1140aada557088a4ull

// Bloom::oldHash before masking, utils/Bloom.h:134-145.  Everything that depends only on the seed
// is folded at compile time.  The reference spells the mixing rounds as shift-adds; on the device they
// are written as what they are -- multiplications by constants (h + (h<<3) + (h<<8) = 265 h, ...) and
// xor-shifts whose high word is a multiply-high by a power of two -- so that ptxas emits IMAD /
// IMAD.WIDE / IMAD.HI on the fma pipe instead of LEA / SHF chains on the alu pipe: the kernels that hash
// are integer-issue bound and the alu pipe was the busy one (34 of 41 instructions -> 18 of 35).
#ifndef FAUCET_HASH_VARIANT
#define FAUCET_HASH_VARIANT 2
#endif
FHD uint64_t xorshr(uint64_t h, int s) {  // h ^ (h >> s), 0 < s < 32
#if defined(__CUDA_ARCH__) && FAUCET_HASH_VARIANT >= 2
  const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
  const uint32_t hs = __umulhi(hi, 1u << (32 - s));  // == hi >> s
  const uint32_t ls = __funnelshift_r(lo, hi, s);
  return ((uint64_t)(hi ^ hs) << 32) | (lo ^ ls);
#else
  return h ^ (h >> s);
#endif
}
template <uint64_t SEED>
FHD uint64_t old_hash(uint64_t key) {
  constexpr uint64_t s7 = SEED ^ (SEED << 7);
  constexpr uint64_t s3 = SEED >> 3, s11 = SEED << 11, s5 = SEED >> 5;
  uint64_t h = s7 ^ (key * s3) ^ (~(s11 + (key ^ s5)));
#if defined(__CUDA_ARCH__) && FAUCET_HASH_VARIANT >= 1
  h = h * 0x1fffffull - 1ull;  // (~h) + (h << 21)
  h = xorshr(h, 24);
  h = h * 265ull;              // (h + (h << 3)) + (h << 8)
  h = xorshr(h, 14);
  h = h * 21ull;               // (h + (h << 2)) + (h << 4)
  h = xorshr(h, 28);
  h = h * 0x80000001ull;       // h + (h << 31)
#else
  h = (~h) + (h << 21);
  h = h ^ (h >> 24);
  h = (h + (h << 3)) + (h << 8);
  h = h ^ (h >> 14);
  h = (h + (h << 2)) + (h << 4);
  h = h ^ (h >> 28);
  h = h + (h << 31);
#endif
  return h;
}
FHD uint64_t hash0(uint64_t key) { return old_hash<FAUCET_SEED0>(key); }
FHD uint64_t hash1(uint64_t key) { return old_hash<FAUCET_SEED1>(key); }

}  // namespace faucet
