// Parse stage: raw FASTA/FASTQ text resident in HBM -> bit planes over RAW BYTE OFFSETS.
//
//   inval  : 1 bit / byte.  0 <=> the byte is an upper-case A/C/G/T on a sequence line of a complete
//            record (what getline + getUnambiguousReads would hand to the k-mer loops:
//            utils/Bloom.cpp:280-286, utils/Kmer.cpp:50-80).
//   packed : 2 bits / byte, NT2int code (utils/Kmer.cpp:82-88), big-endian inside each u32 so that a
//            k-mer is a funnel shift of three consecutive words.
//   skipA  : 1 bit / byte, set over every line that holds >= 2 segments of length >= k.  Those lines
//            are the only place where getUnambiguousReads' REVERSED segment order (push_front,
//            utils/Kmer.cpp:77) changes the stream order of k-mers; the load pass handles them in a
//            dedicated kernel with explicit timestamps and the flat kernel skips them.
//
// No compaction of sequence lines is done: k-mer timestamps are raw byte offsets, which are
// monotone in stream order, and warps that land on header/quality bytes retire after one test.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer.cuh"

namespace faucet {

constexpr int PARSE_THREADS = 256;
constexpr int PARSE_BYTES_PER_THREAD = 32;
constexpr int PARSE_CHUNK = PARSE_THREADS * PARSE_BYTES_PER_THREAD;  // 8 KiB of text per CTA

struct ParseCounters {
  unsigned long long total_newlines;   // '\n' bytes in the batch
  unsigned long long cut;              // offset just past the last complete record (non-final batches)
  unsigned int n_complex;              // entries in the complex-line list
  unsigned int complex_overflow;
};

// bit i of the result <=> byte i of the 32-byte span satisfies the test.  A byte-wise compare leaves 0xff per
// matching byte; (x & 0x01010101) * 0x01020408 gathers the four flags of a word into its top nibble.
__device__ __forceinline__ uint32_t gather4(uint32_t eq) { return ((eq & 0x01010101u) * 0x01020408u) >> 24; }
__device__ __forceinline__ uint32_t newline_mask32(const uint4 a, const uint4 b) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) m |= gather4(__vcmpeq4(w[i], 0x0a0a0a0au)) << (4 * i);
  return m;
}
// upper-case A / C / G / T (isValidNuc, utils/Kmer.cpp:50-60)
__device__ __forceinline__ uint32_t base_mask32(const uint4 a, const uint4 b) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t eq = __vcmpeq4(w[i], 0x41414141u) | __vcmpeq4(w[i], 0x43434343u) | __vcmpeq4(w[i], 0x47474747u) | __vcmpeq4(w[i], 0x54545454u);
    m |= gather4(eq) << (4 * i);
  }
  return m;
}
// NT2int codes ((c >> 1) & 3) of four bytes, first byte in the top two bits of the returned byte
__device__ __forceinline__ uint32_t codes4(uint32_t w) { return (((w >> 1) & 0x03030303u) * 0x40100401u) >> 24; }

// text must be readable (padded) up to a multiple of PARSE_CHUNK; bytes >= n are ignored.
__global__ void __launch_bounds__(PARSE_THREADS)
parse_count_kernel(const uint8_t* __restrict__ text, size_t n, uint32_t* __restrict__ chunk_counts) {
  size_t off = ((size_t)blockIdx.x * PARSE_THREADS + threadIdx.x) * PARSE_BYTES_PER_THREAD;
  uint32_t cnt = 0;
  if (off < n) {
    const uint4* p = reinterpret_cast<const uint4*>(text + off);
    uint32_t m = newline_mask32(p[0], p[1]);
    size_t rem = n - off;
    if (rem < 32) m &= (1u << rem) - 1u;
    cnt = __popc(m);
  }
  __shared__ uint32_t wsum[PARSE_THREADS / 32];
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < PARSE_THREADS / 32; i++) t += wsum[i];
    chunk_counts[blockIdx.x] = t;
  }
}

// single-CTA exclusive scan of the per-chunk newline counts (n_chunks ~ text/8KiB)
__global__ void __launch_bounds__(1024)
parse_scan_kernel(uint32_t* __restrict__ chunk_counts, uint32_t n_chunks, ParseCounters* __restrict__ ctr) {
  __shared__ unsigned long long wtot[32];
  __shared__ unsigned long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_chunks; base += 1024) {
    uint32_t i = base + threadIdx.x;
    unsigned long long v = i < n_chunks ? chunk_counts[i] : 0, x = v;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned long long t = wtot[threadIdx.x], s = t;
      for (int o = 1; o < 32; o <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, s, o);
        if (threadIdx.x >= o) s += y;
      }
      wtot[threadIdx.x] = s - t;  // exclusive warp offsets
    }
    __syncthreads();
    unsigned long long excl = carry_s + wtot[threadIdx.x >> 5] + x - v;
    // line indices only matter modulo the record period and for comparisons below 2^32 lines/batch
    if (i < n_chunks) chunk_counts[i] = (uint32_t)excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) ctr->total_newlines = carry_s;
}

struct ParseArgs {
  const uint8_t* text;
  size_t n;
  uint32_t* inval;
  uint32_t* packed;
  uint32_t* skipA;
  const uint32_t* chunk_prefix;
  ParseCounters* ctr;
  uint2* complex_list;
  uint32_t complex_cap;
  int period_mask;  // 3 for FASTQ (4-line records), 1 for FASTA
  int final_batch;
  int k;
  // record table for the stitch: sequence line of record r = text[seq_start[r], seq_end[r])
  // (both pre-zeroed by the caller: a record whose sequence line is missing reads as empty)
  uint32_t* seq_start;
  uint32_t* seq_end;
  int rec_shift;    // 2 for FASTQ, 1 for FASTA
  uint32_t rec_cap; // entries in seq_start / seq_end (records beyond it are dropped; the host checks the count)
};

// a thread that owns the FIRST non-ACGT byte of a sequence line checks whether the line holds two
// or more segments of length >= k; if so the line is published as "complex".
__device__ void parse_check_complex(const ParseArgs& a, size_t p) {
  const uint8_t* t = a.text;
  size_t s = p;
  while (s > 0 && t[s - 1] != '\n') {
    --s;
    if (!nt_valid(t[s])) return;  // an earlier invalid byte owns this line
  }
  size_t e = s;
  int nseg = 0;
  size_t run = 0;
  while (e < a.n && t[e] != '\n') {
    if (nt_valid(t[e])) {
      run++;
    } else {
      if (run >= (size_t)a.k) nseg++;
      run = 0;
    }
    e++;
  }
  if (run >= (size_t)a.k) nseg++;
  if (nseg < 2) return;
  uint32_t slot = atomicAdd(&a.ctr->n_complex, 1u);
  if (slot >= a.complex_cap) { a.ctr->complex_overflow = 1; return; }
  a.complex_list[slot] = make_uint2((uint32_t)s, (uint32_t)e);
  for (size_t w = s >> 5; w <= (e - 1) >> 5; w++) {
    uint32_t lo = w == (s >> 5) ? (uint32_t)(s & 31) : 0u;
    uint32_t hi = w == ((e - 1) >> 5) ? (uint32_t)((e - 1) & 31) : 31u;
    uint32_t m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
    atomicOr(&a.skipA[w], m);
  }
}

__global__ void __launch_bounds__(PARSE_THREADS)
parse_planes_kernel(ParseArgs a) {
  size_t off = ((size_t)blockIdx.x * PARSE_THREADS + threadIdx.x) * PARSE_BYTES_PER_THREAD;
  uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
  uint32_t nlm = 0;
  if (off < a.n) {
    const uint4* p = reinterpret_cast<const uint4*>(a.text + off);
    v0 = p[0];
    v1 = p[1];
    nlm = newline_mask32(v0, v1);
    size_t rem = a.n - off;
    if (rem < 32) nlm &= (1u << rem) - 1u;
  }
  // CTA-wide exclusive scan of per-thread newline counts
  uint32_t cnt = __popc(nlm), x = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  __shared__ uint32_t wtot[PARSE_THREADS / 32];
  if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
  __syncthreads();
  uint32_t woff = 0;
  for (int i = 0; i < (int)(threadIdx.x >> 5); i++) woff += wtot[i];
  uint32_t line = a.chunk_prefix[blockIdx.x] + woff + x - cnt;

  const unsigned long long total_nl = a.ctr->total_newlines;
  const uint32_t pm = (uint32_t)a.period_mask;
  // non-final batch: lines that do not belong to a complete record are left for the next batch
  const unsigned long long complete = a.final_batch ? ~0ull : (total_nl & ~(unsigned long long)pm);
  // std::getline quirk: an unterminated trailing HEADER line is re-used as its own sequence line
  // (failed sentry leaves the string untouched: utils/Bloom.cpp:280-282, src/ReadScanner.cpp:306-308)
  unsigned long long quirk = ~0ull;
  if (a.final_batch && a.n > 0 && a.text[a.n - 1] != '\n' && (total_nl & pm) == 0) quirk = total_nl;

  // bytes at or past n (the CTA grid covers n rounded up to PARSE_CHUNK) come out as invalid
  const size_t lim = off >= a.n ? 0 : (a.n - off < 32 ? a.n - off : 32);
  const uint32_t limmask = lim >= 32 ? 0xffffffffu : ((1u << lim) - 1u);
  auto is_seq = [&](uint32_t ln) {
    return (((ln & pm) == 1u) && (unsigned long long)ln < complete) || (unsigned long long)ln == quirk;
  };
  // which bytes of the span lie on a sequence line: walk the (usually 0 or 1) newlines of the span
  uint32_t seqm = 0;
  {
    uint32_t m = nlm, pos = 0;
    while (true) {
      const uint32_t nxt = m ? (uint32_t)(__ffs(m) - 1) : 32u;
      const bool sq = is_seq(line);
      if (sq && nxt > pos) seqm |= (nxt >= 32 ? 0xffffffffu : ((1u << nxt) - 1u)) & ~((1u << pos) - 1u);
      if (nxt >= 32) break;
      // the newline at nxt ends `line` (nlm only holds bytes below lim)
      if (sq && (line >> a.rec_shift) < a.rec_cap) a.seq_end[line >> a.rec_shift] = (uint32_t)(off + nxt);
      line++;
      if (!a.final_batch && (unsigned long long)line == complete && complete > 0) a.ctr->cut = off + nxt + 1;
      if (is_seq(line) && (line >> a.rec_shift) < a.rec_cap) a.seq_start[line >> a.rec_shift] = (uint32_t)(off + nxt + 1);
      pos = nxt + 1;
      m &= m - 1;
      if (pos >= 32) break;
    }
  }
  const uint32_t vm = base_mask32(v0, v1);
  const uint32_t inval = ~(seqm & vm & limmask);
  // a non-base on a sequence line (N, lower case, '\r'): the thread that owns it checks the line for several segments
  for (uint32_t cm = seqm & ~vm & ~nlm & limmask; cm; cm &= cm - 1) parse_check_complex(a, off + (__ffs(cm) - 1));
  const uint32_t pk0 = (codes4(v0.x) << 24) | (codes4(v0.y) << 16) | (codes4(v0.z) << 8) | codes4(v0.w);
  const uint32_t pk1 = (codes4(v1.x) << 24) | (codes4(v1.y) << 16) | (codes4(v1.z) << 8) | codes4(v1.w);
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.final_batch && a.n > 0 && a.text[a.n - 1] != '\n') {
    // unterminated last line: it is either a sequence line or the re-used header (quirk); a quirk line
    // that is the only line of the batch starts at offset 0, which the zero-fill already says
    if ((quirk != ~0ull || (total_nl & pm) == 1u) && (total_nl >> a.rec_shift) < a.rec_cap)
      a.seq_end[total_nl >> a.rec_shift] = (uint32_t)a.n;
  }
  a.inval[off >> 5] = inval;
  a.packed[off >> 4] = pk0;
  a.packed[(off >> 4) + 1] = pk1;
}

}  // namespace faucet
