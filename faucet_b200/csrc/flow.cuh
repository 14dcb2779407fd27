// Pass 2, stream-order part WITHOUT rounds: the dataflow form of the ordered kernel (stitch.cuh).
//
// A record may only run after every earlier record that shares one of its reservation slots (minimizers of its
// k-mers) -- that is the whole ordering constraint of the stitch; the rounds of stitch_kernel enforce it with two
// grid barriers per round and re-derive it every round.  Here it is computed ONCE per list of records:
//   rows     slots of every record of the list                                    (flow_rows_kernel)
//   pairs    (slot, list index, column) for every slot of every record            (flow_pairs_kernel)
//   sort     stable LSD radix sort of the pairs by slot: inside a slot the pairs stay in list order
//            (radix_hist_kernel / radix_scatter_kernel, 8 bits per pass)
//   preds    the predecessor of (record, column) = the previous pair of the same slot  (flow_preds_kernel)
// and then one persistent kernel runs the list: warps take records in list order from an atomic counter, wait
// until the predecessors of their record are flagged done, run the SAME walk as the ordered kernel
// (scan_line<.., false>), fence, flag the record done.  No barrier; a chain of records that share a slot costs
// one record latency per link instead of one round.  Progress: records are taken in increasing order and only
// wait for smaller ones, and every taken record is held by a resident warp.
#pragma once
#include "stitch.cuh"

namespace faucet {

#ifndef FAUCET_FLOW_BLOCKS
#define FAUCET_FLOW_BLOCKS 3
#endif
#ifndef FAUCET_FLOW_SLEEP
#define FAUCET_FLOW_SLEEP 100
#endif
constexpr uint32_t PRED_NONE = 0xffffffffu;
constexpr int RADIX_SUB = 512;      // pairs per warp of the radix passes (16 steps of 32, in order: stable)
constexpr int RADIX_THREADS = 256;
constexpr int FLOW_COL_BITS = 5, FLOW_IDX_BITS = 32;  // pair = slot << 37 | list index << 5 | column

struct FlowArgs {
  uint32_t n;                 // entries of the list
  uint32_t begin;             // list == NULL: entry i is record begin + i
  uint32_t* rows;             // ROW_WORDS per entry
  uint32_t* preds;            // ROW_WORDS per entry: preds[i][1 + c] = list index that must be done first, or PRED_NONE
  uint32_t* done;             // per entry
  uint32_t* counts;           // per entry: slots (then their exclusive prefix sum)
  unsigned long long* pairs;  // in
  unsigned long long* pairs2; // out
  uint32_t* hist;             // 256 x n_sub counters, digit-major
  uint32_t n_sub;
  uint32_t n_pairs;
  int shift;
  unsigned int* big;          // records with more slots than a row holds (they need the round-based kernel)
  unsigned int* next;         // the ticket counter, alone in its cache line (every warp of the grid hammers it)
};
// Tickets.  One global counter hit by every warp for every record is a same-address atomic per record (~6 ns each,
// serialised at one L2 slice: it alone capped the executor at ~160 M records/s); a warp that takes SEVERAL tickets at
// once sits on records it has not started, and everything that depends on them waits (4 per warp: 1.8x slower, 16:
// 17x).  So each CTA draws FAUCET_FLOW_POOL tickets at a time into a shared-memory pool and its warps take them one by
// one: the global atomics drop by that factor and a drawn ticket is picked up within a fraction of a record time.
// The jslot run filter (stitch.cuh, prefetch_line) halves the executor's L2 sectors and skips the hashing of two thirds of
// the positions, for one more dependent load in front of the key probes (measured: 20.3 vs 21.1 ms at configs[1]).
#ifndef FAUCET_FLOW_JSLOT
#define FAUCET_FLOW_JSLOT 1
#endif
#ifndef FAUCET_FLOW_POOL
#define FAUCET_FLOW_POOL 8
#endif

__global__ void __launch_bounds__(DRY_THREADS) flow_rows_kernel(StitchArgs a, FlowArgs f) {
  __shared__ uint32_t keep_s[DRY_WARPS][ROW_WORDS];
  uint32_t* keep = keep_s[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * DRY_THREADS + threadIdx.x) >> 5, n_warps = (gridDim.x * DRY_THREADS) >> 5;
  for (uint32_t i = gw; i < f.n; i += n_warps) {
    const unsigned long long t_take = gtime_ns();
    const uint32_t rec = a.list ? __ldg(a.list + i) : f.begin + i;
    const uint32_t ls = __ldg(a.seq_start + rec), le = __ldg(a.seq_end + rec);
    const uint32_t len = le > ls ? le - ls : 0u;
    int n = 0;
    line_reservations<3, false>(a, a.packed, ls, len, rec, lane, keep, &n, ROW_WORDS - 1);
    __syncwarp();
    uint32_t* row = f.rows + (size_t)i * ROW_WORDS;
    const bool big = n >= ROW_WORDS;
    if (lane == 0) {
      row[0] = big ? 0u : (uint32_t)n;
      f.counts[i] = big ? 0u : (uint32_t)n;
      if (big) atomicAdd(f.big, 1u);
    }
    if (!big && lane < n) row[1 + lane] = keep[lane];
    __syncwarp();
  }
}

// counts[] holds the exclusive prefix sum of the slot counts
__global__ void flow_pairs_kernel(FlowArgs f) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < f.n; i += gridDim.x * blockDim.x) {
    const uint32_t* row = f.rows + (size_t)i * ROW_WORDS;
    uint32_t* pr = f.preds + (size_t)i * ROW_WORDS;
    const uint32_t c = row[0], off = f.counts[i];
    for (uint32_t q = 0; q < c; q++) {
      f.pairs[off + q] = ((unsigned long long)(row[1 + q] & ROW_SLOT_MASK) << (FLOW_IDX_BITS + FLOW_COL_BITS)) | ((unsigned long long)i << FLOW_COL_BITS) | q;
      pr[1 + q] = PRED_NONE;
    }
  }
}

__device__ __forceinline__ uint32_t flow_digit(unsigned long long p, int shift) {
  return (uint32_t)(p >> (FLOW_IDX_BITS + FLOW_COL_BITS + shift)) & 255u;
}

// one warp per RADIX_SUB consecutive pairs: digit counts of its sub-tile -> hist[digit * n_sub + sub]
__global__ void __launch_bounds__(RADIX_THREADS) radix_hist_kernel(FlowArgs f) {
  __shared__ uint32_t cnt_s[RADIX_THREADS / 32][256];
  uint32_t* cnt = cnt_s[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const uint32_t sub = (blockIdx.x * RADIX_THREADS + threadIdx.x) >> 5;
  if (sub >= f.n_sub) return;
  for (int d = lane; d < 256; d += 32) cnt[d] = 0;
  __syncwarp();
  const uint32_t base = sub * RADIX_SUB;
  for (int step = 0; step < RADIX_SUB / 32; step++) {
    const uint32_t q = base + step * 32 + lane;
    const bool act = q < f.n_pairs;
    const uint32_t d = act ? flow_digit(f.pairs[q], f.shift) : 0xffffffffu;
    const uint32_t m = __match_any_sync(0xffffffffu, d);
    if (act && lane == __ffs(m) - 1) cnt[d] += __popc(m);
    __syncwarp();
  }
  for (int d = lane; d < 256; d += 32) f.hist[(size_t)d * f.n_sub + sub] = cnt[d];
}
// hist[] holds its exclusive prefix sum: where the first pair of (digit, sub) goes
__global__ void __launch_bounds__(RADIX_THREADS) radix_scatter_kernel(FlowArgs f) {
  __shared__ uint32_t pos_s[RADIX_THREADS / 32][256];
  uint32_t* pos = pos_s[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const uint32_t sub = (blockIdx.x * RADIX_THREADS + threadIdx.x) >> 5;
  if (sub >= f.n_sub) return;
  for (int d = lane; d < 256; d += 32) pos[d] = f.hist[(size_t)d * f.n_sub + sub];
  __syncwarp();
  const uint32_t base = sub * RADIX_SUB;
  for (int step = 0; step < RADIX_SUB / 32; step++) {
    const uint32_t q = base + step * 32 + lane;
    const bool act = q < f.n_pairs;
    const unsigned long long p = act ? f.pairs[q] : 0ull;
    const uint32_t d = act ? flow_digit(p, f.shift) : 0xffffffffu;
    const uint32_t m = __match_any_sync(0xffffffffu, d);
    uint32_t at = 0;
    if (act) at = pos[d] + __popc(m & ((1u << lane) - 1u));
    __syncwarp();
    if (act && lane == __ffs(m) - 1) pos[d] += __popc(m);
    __syncwarp();
    if (act) f.pairs2[at] = p;
  }
}

// sorted by slot, list order inside a slot: the predecessor of a pair is the closest earlier pair of the same slot
// that belongs to ANOTHER record (two runs of one line may hash to the same slot)
__global__ void flow_preds_kernel(FlowArgs f) {
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < f.n_pairs; q += gridDim.x * blockDim.x) {
    const unsigned long long p = f.pairs[q];
    const unsigned long long slot = p >> (FLOW_IDX_BITS + FLOW_COL_BITS);
    const uint32_t idx = (uint32_t)(p >> FLOW_COL_BITS), col = (uint32_t)p & ((1u << FLOW_COL_BITS) - 1u);
    uint32_t pred = PRED_NONE;
    for (uint32_t b = q; b > 0;) {
      b--;
      const unsigned long long o = f.pairs[b];
      if ((o >> (FLOW_IDX_BITS + FLOW_COL_BITS)) != slot) break;
      const uint32_t oi = (uint32_t)(o >> FLOW_COL_BITS);
      if (oi != idx) { pred = oi; break; }
    }
    f.preds[(size_t)idx * ROW_WORDS + 1 + col] = pred;
  }
}

// The done flags are polled with an L2 (ld.cg) load.  An acquire -- or even relaxed -- gpu-scope load is followed by an invalidation of the
// SM's whole L1 (CCTL.IVALL in SASS), per poll, for every warp of the SM -- and nothing here needs it: what a record reads
// after the wait and another record may have written (keys, records, stamps, jslot) is read with ld.cg, i.e. from L2,
// where the writer's atomics and its release store were performed in order; the loads are issued after the branch on
// the flag value (no speculation), and what IS cached in L1 (text planes, flags, rows) is never written by this kernel.
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// The persistent executor.  *f.next = next list entry to hand out; entries already flagged done (a relaunch after
// the table grew or the extension buffer was drained) are skipped.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(STITCH_THREADS, MIN_BLOCKS) stitch_flow_kernel(StitchArgs a, FlowArgs f) {
  extern __shared__ __align__(16) unsigned char stitch_smem[];
  WarpScratch* S = reinterpret_cast<WarpScratch*>(stitch_smem) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const uint32_t n_warps = (gridDim.x * STITCH_THREADS) >> 5;
  StitchState* st = a.st;
  WarpCtx c;
  if (lane < SS_COUNT) S->st[lane] = 0;
  __syncwarp();
  c.S = S; c.stage = S->stage; c.cnt = S->st; c.n_stage = 0; c.part = 0; c.stamp = 0; c.n_vis = 0; c.wrote = false; c.grec = 0;
  c.land_slot = nullptr; c.land_nt = nullptr; c.n_land = 0; c.emit = true;
  __shared__ unsigned long long pool;  // {end : 32 | next : 32} of the CTA's drawn tickets
  if (threadIdx.x == 0) pool = 0ull;
  __syncthreads();
  while (true) {
    uint32_t i = 0;
    if (lane == 0) {
      while (true) {
        const unsigned long long v = atomicAdd(&pool, 1ull);  // optimistic: take the next ticket of the pool
        const uint32_t nx = (uint32_t)v, en = (uint32_t)(v >> 32);
        if (nx < en) { i = nx; break; }
        if (nx == en) {  // this warp found the pool just empty: it draws the next batch and keeps its first ticket
          i = atomicAdd(f.next, (unsigned int)FAUCET_FLOW_POOL);
          atomicExch(&pool, ((unsigned long long)(i + FAUCET_FLOW_POOL) << 32) | (unsigned long long)(i + 1u));
          break;
        }
        while ((uint32_t)(*(volatile unsigned long long*)&pool >> 32) == en) {}  // somebody is drawing: retry after
      }
    }
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= f.n) break;
    if (ld_acquire_u32(f.done + i)) continue;
    const unsigned long long t_take = gtime_ns();
    const uint32_t rec = a.list ? __ldg(a.list + i) : f.begin + i;
    const uint32_t ls = __ldg(a.seq_start + rec), le = __ldg(a.seq_end + rec);
    const uint32_t len = le > ls ? le - ls : 0u;
    const int n_pos = len >= (uint32_t)a.k ? (int)(len - a.k + 1) : 0;
    const bool fast = n_pos > 0 && n_pos <= POS_CAP;
    // what does not depend on the junction table is fetched before the wait
    if (fast) {
      if (lane < PK_WORDS) S->pk[lane] = __ldg(a.packed + (ls >> 4) + lane);
      if (lane < INV_WORDS) S->inv[lane] = __ldg(a.inval + (ls >> 5) + lane);
      for (int pos = lane; pos < n_pos; pos += 32) S->flag[pos] = a.flags[ls + pos];
    }
    const uint32_t* pr = f.preds + (size_t)i * ROW_WORDS;
    const uint32_t n_row = __ldg(f.rows + (size_t)i * ROW_WORDS);
    const uint32_t pred = lane < (int)n_row ? __ldg(pr + 1 + lane) : PRED_NONE;
    // ---- room for what this record and the other resident warps may create?  (Read before the wait: the bound covers
    // every record in flight, so what the predecessors create meanwhile is already counted.)
    unsigned int stop = ST_DONE;
    {
      unsigned long long need = __ldcg(&st->max_need);
      if (2ull * len + 2 > need) { need = 2ull * len + 2; if (lane == 0) atomicMax(&st->max_need, need); }
      const unsigned long long bound = need * n_warps;
      if (__ldcg(&st->n_entries) + bound > a.cap / 2) stop = ST_GROW_TABLE;
      else if (a.ext && __ldcg(&st->ext_used) + 2 * bound + n_warps > a.ext_cap) stop = ST_DRAIN_EXT;
    }
    // ---- wait for the earlier records that share a slot with this one (or for the run to be called off)
    bool off = false;
    const unsigned long long t_wait = gtime_ns();
    uint32_t polls = 0;
    while (true) {
      const bool ready = pred == PRED_NONE || ld_acquire_u32(f.done + pred) != 0u;
      if (__all_sync(0xffffffffu, ready)) break;
      if (__ldcg(&st->status) != ST_DONE) { off = true; break; }
      polls++;
      __nanosleep(FAUCET_FLOW_SLEEP);
    }
    const unsigned long long t_go = gtime_ns();
    if (lane == 0) {  // executor statistics (reported as faucet_timings.stitch_phase_ns[0..5])
      S->st[SS_T_PHASE1] += polls; S->st[SS_T_SYNC1] += polls ? 1 : 0; S->st[SS_T_PHASE2] += t_go - t_wait;
      S->st[SS_T_P1A] += t_wait - t_take;
    }
    if (off) continue;  // (keeps taking entries: they all see the status and fall through)
    if (stop != ST_DONE) { if (lane == 0) atomicCAS(&st->status, (unsigned int)ST_DONE, stop); continue; }
    if (__ldcg(&st->status) != ST_DONE) continue;
    __syncwarp();
    c.rec = rec; c.part = 0; c.n_stage = 0; c.n_vis = 0; c.ls = ls; c.wrote = false;
    c.grec = global_rec(a, rec);
    c.stamp = global_stamp(a, rec);
    if (fast) {
      if (lane < (int)n_row) S->reskey[lane] = __ldg(f.rows + (size_t)i * ROW_WORDS + 1 + lane);
      __syncwarp();
      // (looking every position up at once instead of asking the run filter first was tried for the gathered exact set of
      // a sharded epoch, which is bound by its chains of dependent records: one dependent load less per link, but 3x the
      // probes -- 28.4 instead of 25.9 ms at 8 GPUs)
      prefetch_line<2>(a, S, ls, n_pos, lane, FAUCET_FLOW_JSLOT ? (int)n_row : -1);
      if (lane == 0) S->st[SS_T_P1C] += gtime_ns() - t_go;
      c.n_pos = n_pos; c.pk = S->pk; c.pk_base = ls & ~15u; c.inv = S->inv; c.inv_base = ls & ~31u;
      scan_line<true, false>(a, c, ls, ls + len, lane);
    } else if (len) {
      c.n_pos = 0; c.pk = a.packed; c.pk_base = 0; c.inv = a.inval; c.inv_base = 0;
      scan_line<false, false>(a, c, ls, ls + len, lane);
    }
    if (a.ext && c.n_stage) ext_flush(a, c, lane);
    if (c.wrote && lane == 0) S->st[SS_WRITERS]++;
    if (lane == 0) S->st[SS_T_P2A] += gtime_ns() - t_go;
    // every write a later record may read (keys, stamps, records, jslot / dirty marks) was issued by lane 0: its release
    // store orders them before the flag
    __syncwarp();
    if (lane == 0) {
      st_release_u32(f.done + i, 1u);
      S->st[SS_T_SYNC2] += gtime_ns() - t_go; S->st[SS_T_P1B] += 1;
    }
  }
  __syncwarp();
  if (lane < SS_COUNT && S->st[lane]) atomicAdd(&st->stats[lane], S->st[lane]);
}

}  // namespace faucet
