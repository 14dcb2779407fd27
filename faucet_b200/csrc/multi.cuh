// Multi-GPU pieces (one process per GPU on one NVSwitch box, peer buffers mapped through CUDA IPC).
//
// Pass 1 stays EXACT when the stream is cut into contiguous shards, one per GPU:
//   bloo1 at the start of shard g  =  OR of the bits of every k-mer of shards 0..g-1
// (the reference only skips bloo1 adds that are no-ops, utils/Bloom.cpp:293-298), so
//   1. every GPU ORs all k-mers of its shard into a plain array           (bloom_add_all_kernel)
//   2. exclusive prefix-OR over the GPUs, read straight from peer HBM      (bloom_prefix_or_kernel)
//   3. the exact two-filter load of load.cuh runs on the shard, starting from that bloo1
//   4. the per-shard bloo2 arrays are OR-all-reduced in place over NVLink  (bloom_or_allreduce_kernel;
//      NCCL has no bitwise-OR reduction)
// Pass 2: scan_flags shards freely (pure); the stitch is one sequential state -- shard.cuh shards what of it commutes
// (the sharded epoch); the serial form (GPU 0 pulls the other GPUs' planes over NVLink shard by shard) stays as the
// path of scans that feed the pair filters.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer.cuh"
#include "load.cuh"

namespace faucet {

constexpr int MAX_PEERS = 16;

// every valid k-mer of the parsed batch -> all its bits into `bits` (plain reference layout)
template <int NH>
__global__ void __launch_bounds__(LOAD_THREADS) bloom_add_all_kernel(LoadArgs a, uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * LOAD_THREADS + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * LOAD_THREADS) >> 5;
  const int nh = NH ? NH : a.n_hash;
  const uint64_t kbits = a.k >= 32 ? 0xffffffffull : ((1ull << a.k) - 1ull);
  for (uint32_t w = warp; w < a.n_words; w += n_warps) {
    uint32_t lo = __ldg(a.inval + w), hi = __ldg(a.inval + w + 1);
    if ((inval_window(lo, hi, lane) & kbits) != 0) continue;
    const uint64_t fwd = kmer_at(a.packed, (w << 5) + lane, a.k);
    const uint64_t c = canon(fwd, revcomp(fwd, a.k));
    uint64_t h = hash0(c) & a.tai_mask;
    const uint64_t h1 = hash1(c) & a.tai_mask;
#pragma unroll
    for (int i = 0; i < (NH ? NH : MAX_NHASH); i++) {
      if (i >= nh) break;
      const uint32_t bit = 1u << (h & 31);
      if (!(__ldg(bits + (h >> 5)) & bit)) atomicOr(bits + (h >> 5), bit);
      h = (h + h1) & a.tai_mask;
    }
  }
}

struct PeerPtrs {
  const uint32_t* in[MAX_PEERS];
  uint32_t* out[MAX_PEERS];
};

// fused[i] = { OR of peers[0..n_before) bloo1 word i , 0 }.  Every bit that is set this way was "first
// touched before this shard": its stamp becomes 0, which is what kernel B of the load compares against.
__global__ void bloom_prefix_or_kernel(PeerPtrs p, int n_before, unsigned long long* __restrict__ fused,
                                       uint32_t* __restrict__ stamps, uint64_t n_words) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t v = 0;
    for (int r = 0; r < n_before; r++) v |= p.in[r][i];
    fused[i] = (unsigned long long)v;
    for (uint32_t m = v; m; m &= m - 1) stamps[i * 32 + (__ffs(m) - 1)] = 0u;
  }
}

// in-place OR all-reduce: this GPU owns the words [w0, w1); it ORs that range over all peers and writes
// the result back to every peer (peer word i is only ever touched by its owner => no race)
__global__ void bloom_or_allreduce_kernel(PeerPtrs p, int n_ranks, uint64_t w0, uint64_t w1) {
  // 16 bytes per thread per step
  const uint64_t v0 = w0 / 4, v1 = w1 / 4;
  for (uint64_t i = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += (uint64_t)gridDim.x * blockDim.x) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int r = 0; r < n_ranks; r++) {
      const uint4 x = reinterpret_cast<const uint4*>(p.out[r])[i];
      acc.x |= x.x; acc.y |= x.y; acc.z |= x.z; acc.w |= x.w;
    }
    for (int r = 0; r < n_ranks; r++) reinterpret_cast<uint4*>(p.out[r])[i] = acc;
  }
}

}  // namespace faucet
