// Pass 2, stream-order part, ACROSS GPUs: the sharded epoch.
//
// The stitch (stitch.cuh) is one sequential state, but after the first few x coverage almost every record is QUIET: it
// creates no junction, raises no stored distance, sets no new link -- it only counts coverage, and coverage counts
// commute and are never read by a walk (src/ReadScanner.cpp:134-192, utils/Junction.cpp:59-71).  The single-GPU epoch
// scheme of stitch.cuh (classify / execute / verify / apply) exploits that inside one GPU; its data-parallel steps cost
// about as much as executing the records in order, so it does not pay there.  It does pay when those steps run on the
// GPU that owns the records, all GPUs at once:
//
//   prefix     GPU 0 runs the first records of shard 0 through the ordered executor: the table T0.
//   replicate  every other GPU copies T0's keys / distances / links -- packed by GPU 0 into 16 bytes per junction -- out of
//              GPU 0's HBM (NVLink) and rebuilds the table locally.
//   classify   every GPU walks ITS OWN records read-only against its replica of T0, all records in parallel.  A quiet
//              record commits at once: coverage counts into the GPU's own count array (4 u32 per table slot), scan
//              counters into the GPU's own counter block.  The others form the exact set E (per-shard ascending lists).
//   execute    EVERY GPU runs the WHOLE exact set -- the lists of all shards, in stream order -- through the ordered
//              executor on its own replica; the lines of foreign shards are gathered from their owner's planes over
//              NVLink into one small local batch first.  The executor is sequential-equivalent, hence deterministic in everything a walk reads
//              (keys, distances, links, creation stamps): all replicas stay identical in those fields without a single
//              message, and so do the per-slot "written by" marks (global record indices).
//   verify     every GPU checks its own quiet records against the marks: a record with an earlier write under one of its
//              reservation slots may have seen a stale T0.  Earlier writes only -> it is walked again on the replica as
//              it stands now (= what it would have seen); earlier and later writes -> it takes its commit back and joins
//              E.  The GPUs exchange the sizes of their lists; while any list grew, all replicas are restored to T0 and
//              E runs again.
//   merge      GPU 0 adds the count arrays of all GPUs to its table (looked up by key: the slot of a junction created
//              during the epoch may differ between replicas), and the scan counters.  Its table is the JunctionMap.
//
// Exactness is the induction of stitch.cuh's epochs over the global stream order: a record outside E never has an earlier
// write under its slots that it did not see, so it behaves as in the sequential run and changes nothing a walk reads;
// therefore E, executed in order from T0, sees the sequential state.  The per-GPU work is the classify walk of 1/N of the
// records plus the (small, replicated) exact set; what crosses NVLink is T0 once, the lines of E, and 24 bytes per table
// slot for the merge.  The protocol is sequenced by faucet_b200/multi.py (ShardedJob.scan); every step is a
// faucet_session_shard_* call of the C ABI.
#pragma once
#include "multi.cuh"
#include "stitch.cuh"

namespace faucet {

// ---- the exact set, gathered: every GPU copies the lines of ALL members (its own and the other shards', read from
// their owner's planes over NVLink, thousands of lines in flight at once) into one small local batch in stream order,
// and runs the ordered executor over that batch -- one dependency sort and one launch per iteration, and the chains of
// dependent records (members cluster under the written slots) pay local instead of NVLink latency per link.
// A line keeps its offset modulo 32, so plane words are copied verbatim.
struct GatherArgs {
  const uint32_t* inval[MAX_PEERS];
  const uint32_t* packed[MAX_PEERS];
  const uint8_t* flags[MAX_PEERS];
  const uint32_t* seq_start[MAX_PEERS];
  const uint32_t* seq_end[MAX_PEERS];
  const uint32_t* list[MAX_PEERS];   // ascending record indices of shard g's members
  uint32_t first[MAX_PEERS + 1];     // entries [first[g], first[g+1]) of the gathered batch come from shard g
  uint32_t rec_base[MAX_PEERS];      // global index of record 0 of shard g
  int n_ranks;
  uint32_t n;
  uint32_t* span;                    // per entry: positions reserved for it (a multiple of 32); then their exclusive prefix sum
  uint32_t *o_inval, *o_packed;
  uint8_t* o_flags;
  uint32_t *o_seq_start, *o_seq_end, *o_gid;
};
constexpr uint32_t GATHER_MIN_SPAN = 384;  // >= what the walk stages of a short line (PK_WORDS x 16, INV_WORDS x 32 positions)

__device__ __forceinline__ int gather_shard_of(const GatherArgs& g, uint32_t i) {
  int r = 0;
  while (r + 1 < g.n_ranks && i >= g.first[r + 1]) r++;
  return r;
}
__global__ void shard_gather_spans_kernel(GatherArgs g) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += gridDim.x * blockDim.x) {
    const int r = gather_shard_of(g, i);
    const uint32_t rec = g.list[r][i - g.first[r]];
    const uint32_t ls = g.seq_start[r][rec], le = g.seq_end[r][rec];
    const uint32_t len = le > ls ? le - ls : 0u;
    const uint32_t need = (((ls & 31u) + len + 31u) & ~31u) + 96u;
    g.span[i] = need > GATHER_MIN_SPAN ? need : GATHER_MIN_SPAN;
  }
}
// one warp per entry; span[] holds the exclusive prefix sums
__global__ void __launch_bounds__(256) shard_gather_copy_kernel(GatherArgs g, uint32_t total) {
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * 256u + threadIdx.x) >> 5, n_warps = (gridDim.x * 256u) >> 5;
  for (uint32_t i = gw; i < g.n; i += n_warps) {
    const int r = gather_shard_of(g, i);
    const uint32_t rec = g.list[r][i - g.first[r]];
    const uint32_t ls = g.seq_start[r][rec], le = g.seq_end[r][rec];
    const uint32_t len = le > ls ? le - ls : 0u;
    const uint32_t off = g.span[i], span = (i + 1 < g.n ? g.span[i + 1] : total) - off;
    const uint32_t s0 = ls & ~31u;
    for (uint32_t w = lane; w < span / 16; w += 32) g.o_packed[(off >> 4) + w] = __ldg(g.packed[r] + (s0 >> 4) + w);
    for (uint32_t w = lane; w < span / 32; w += 32) g.o_inval[(off >> 5) + w] = __ldg(g.inval[r] + (s0 >> 5) + w);
    const uint32_t nls = off + (ls & 31u);
    for (uint32_t t = lane; t < len; t += 32) g.o_flags[nls + t] = g.flags[r][ls + t];
    if (lane == 0) { g.o_seq_start[i] = nls; g.o_seq_end[i] = nls + len; g.o_gid[i] = g.rec_base[r] + rec; }
  }
}

// ---- T0 for the other GPUs: what a walk reads of a junction, 16 bytes per OCCUPIED slot instead of 72 per slot
struct PackedJunction {
  unsigned long long key;
  uint8_t dist[5];
  uint8_t link;
  uint8_t pad[2];
};
static_assert(sizeof(PackedJunction) == 16, "PackedJunction layout");

__global__ void shard_pack_kernel(StitchArgs a, PackedJunction* __restrict__ out, unsigned int* __restrict__ n_out) {
  const int lane = threadIdx.x & 31;
  for (unsigned long long base = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; base <= a.cap;
       base += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long i = base + lane;
    bool occ = false;
    if (i < a.cap) occ = a.keys[i] != KEY_EMPTY;
    else if (i == a.cap) occ = *a.special != 0;
    const uint32_t m = __ballot_sync(0xffffffffu, occ);
    if (!m) continue;
    unsigned int at = 0;
    if (lane == 0) at = atomicAdd(n_out, (unsigned int)__popc(m));
    at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
    if (occ) {
      const uint32_t* r = a.recs + i * REC_WORDS;
      PackedJunction p;
      p.key = i == a.cap ? KEY_EMPTY : a.keys[i];
      for (int f = 0; f < 5; f++) p.dist[f] = (uint8_t)r[REC_DIST + f];
      p.link = (uint8_t)r[REC_LINK];
      p.pad[0] = p.pad[1] = 0;
      out[at] = p;
    }
  }
}
// into a cleared table (keys all KEY_EMPTY, records zero); the entry counter and the KEY_EMPTY flag come with the state block
__global__ void shard_unpack_kernel(StitchArgs a, const PackedJunction* __restrict__ in, unsigned int n) {
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const PackedJunction p = in[i];
    unsigned long long h;
    if (p.key == KEY_EMPTY) {
      h = a.cap;
    } else {
      h = mix64(p.key) & (a.cap - 1);
      while (atomicCAS(a.keys + h, KEY_EMPTY, p.key) != KEY_EMPTY) h = (h + 1) & (a.cap - 1);
    }
    uint32_t* r = a.recs + h * REC_WORDS;
    for (int f = 0; f < 5; f++) r[REC_DIST + f] = p.dist[f];
    r[REC_LINK] = p.link;
  }
}

// counts[slot][nt] of a shard (peer HBM, or this GPU's own) -> this GPU's junction records.  same_table: the counts
// are this GPU's own, slot for slot.
__global__ void shard_merge_kernel(const unsigned long long* __restrict__ peer_keys, const uint4* __restrict__ peer_cov,
                                   unsigned long long cap, StitchArgs mine, int same_table) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= cap;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint4 c = peer_cov[i];
    if (!(c.x | c.y | c.z | c.w)) continue;
    int slot = (int)i;
    if (!same_table) {
      const unsigned long long key = i == cap ? KEY_EMPTY : peer_keys[i];
      slot = i < cap && key == KEY_EMPTY ? -1 : tbl_find(mine, key);
    }
    if (slot < 0) { atomicAdd(&mine.st->stats[SS_DRY_ERROR], 1ull); continue; }  // cannot happen: the replicas hold the same keys
    uint32_t* r = rec_field(mine, slot, REC_COV);
    if (c.x) atomicAdd(r + 0, c.x);
    if (c.y) atomicAdd(r + 1, c.y);
    if (c.z) atomicAdd(r + 2, c.z);
    if (c.w) atomicAdd(r + 3, c.w);
  }
}

}  // namespace faucet
