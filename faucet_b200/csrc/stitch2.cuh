// Pass 2, stream-order part, second kernel: ONE THREAD walks a record (stitch.cuh holds the default
// warp-per-record kernel; this one is selected with faucet_gpu_set_tuning("stitch_impl", 2)).
//
// Same schedule as stitch.cuh -- windowed deterministic reservations on minimizers, two grid barriers
// per round, junction table in HBM.  The per-round parallelism of that kernel is capped by "one record
// per warp" (3.5 k records per round on a B200, while the E. coli-sized workload supports ~8 k
// conflict-free records per round at a window of ~14 k).  Here a warp holds 1..16 records per round:
//   * the minimizer reservation slots of every record are a pure function of the text and are listed
//     once per batch by stitch2_rows_kernel (one 128-byte row per record): reserving is <= 31 atomics
//     and checking + releasing is one trip, with no minimizer computed inside the round loop;
//   * phase 1: the G = 32, 16, .. 2 lanes that share a record stage the line's plane words in shared
//     memory and look up the FORWARD and BACKWARD key of every k-mer position; the answers are two
//     128-bit K planes + up to 8 parked (slot, skip distance) pairs in the record's shared-memory slot;
//   * phase 2: one thread per record runs the walk of stitch2_walk.cuh -- find-first-set over
//     scan_flags' bit planes and the K planes instead of a per-half-step loop.
// Lines with more than S2_POS_CAP k-mer positions or more than 31 reservation slots are rare; the
// warp that owns such a record processes it cooperatively with the code of stitch.cuh (direct path).
//
// Measured (B200, 4.6 Mbp x 100x x 150 bp, profiles/r1_stitch2_sweep.txt): bit-exact like stitch.cuh, but
// 53-59 ms against 35.5 ms.  The single-thread walk is ~2.6 k dependent instructions (8.5 us, with a 19 us
// tail behind the second barrier: s2_publish after a creation is a serial scan of the line), and with
// G < 16 lanes per record the lookups of phase 1 stretch faster than the window adds parallelism
// (16 us at G = 16, 40 us at G = 4).  Kept as a tested alternative and as the device half of the
// host-compiled walk that tests/test_stitch2_host.py holds to the oracle.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "stitch.cuh"
#include "stitch2_walk.cuh"

namespace faucet {

constexpr int S2_THREADS = 512;
constexpr int S2_WARPS = S2_THREADS / 32;
static_assert(S2_KEY_EMPTY == KEY_EMPTY, "one empty marker");

// reservation rows: rows[32 r] = number of slots of record row_base + r (> 31: did not fit), then the slots
__global__ void __launch_bounds__(256) stitch2_rows_kernel(StitchArgs a, uint32_t* __restrict__ rows) {
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * 256 + threadIdx.x) >> 5, n_warps = (gridDim.x * 256) >> 5;
  for (uint32_t rec = a.row_base + gw; rec < a.row_end; rec += n_warps) {
    const uint32_t ls = __ldg(a.seq_start + rec), le = __ldg(a.seq_end + rec);
    const uint32_t len = le > ls ? le - ls : 0u;
    uint32_t* row = rows + (size_t)(rec - a.row_base) * S2_ROW;
    int n = 0;
    line_reservations<3, false>(a, a.packed, ls, len, rec, lane, row + 1, &n, S2_ROW - 1);
    if (lane == 0) row[0] = (uint32_t)n;
  }
}

// the line's words of every plane, staged in shared memory by the lanes that share the record
struct LineStage {
  uint32_t fp[6 * FP_STRIDE];  // flag planes, from word ls >> 5 (a line of <= 159 bytes spans <= 6 words)
  uint32_t inv[6];             // validity plane, from word ls >> 5
  uint32_t pk[13];             // 2-bit plane, from word ls >> 4 (11 words + the 2 a k-mer fetch may run over)
  uint32_t pad;
};
struct RecSlot {
  LineStage G;
  LineState L;
};
constexpr int S2_RPW = 32;  // records one warp may hold per round (one lane each); fewer records => more lanes per record
// Shared memory of one warp: its record slots, then the tail of a WarpScratch (reskey, visited, stage, st: all the
// direct path of stitch.cuh touches).  The WarpScratch pointer is placed so that its tail lands there; its head
// (the parking arrays of the warp-per-record kernel) overlays the record slots and is never touched here.
constexpr size_t S2_TAIL = sizeof(WarpScratch) - offsetof(WarpScratch, reskey);
constexpr size_t S2_WARP_BYTES = S2_RPW * sizeof(RecSlot) + S2_TAIL;
static_assert(S2_WARP_BYTES >= sizeof(WarpScratch) && S2_WARP_BYTES % 16 == 0 && (S2_WARP_BYTES - sizeof(WarpScratch)) % 8 == 0, "smem layout");
constexpr size_t S2_SMEM = (size_t)S2_WARPS * S2_WARP_BYTES;

struct DevEnv {
  const StitchArgs& a;
  int k, j, spacer;
  bool pairs, want_ext;
  unsigned st[S2_COUNTERS];
  unsigned n_created;
  unsigned long long stamp_next;
  uint32_t rec, part, n_ext;
  const LineStage* G;
  uint32_t w5, w4;  // ls >> 5, ls >> 4: word 0 of the staged planes
  unsigned long long ext_buf[S2_EXT];

  __device__ __forceinline__ uint32_t inval_word(uint32_t w) const { return G->inv[w - w5]; }
  __device__ __forceinline__ uint32_t fp_word(int p, uint32_t w) const { return G->fp[(w - w5) * FP_STRIDE + p]; }
  __device__ __forceinline__ uint32_t packed_word(uint32_t w) const { return G->pk[w - w4]; }
  __device__ __forceinline__ uint64_t tbl_home(uint64_t key) const { return mix64(key) & (a.cap - 1); }
  __device__ __forceinline__ uint64_t tbl_next(uint64_t h) const { return (h + 1) & (a.cap - 1); }
  __device__ __forceinline__ uint64_t tbl_key(uint64_t h) const { return __ldcg(a.keys + h); }
  __device__ __forceinline__ int find(uint64_t key) const { return tbl_find(a, key); }
  __device__ __forceinline__ int insert(uint64_t key, bool* created) {
    const int s = tbl_insert_nc(a, key, created);
    if (*created) n_created++;
    return s;
  }
  __device__ __forceinline__ void stamp(int slot) { a.stamps[slot] = stamp_next++; }
  __device__ __forceinline__ uint32_t dist_peek(int slot, int idx) const { return __ldcg(rec_field(a, slot, REC_DIST + idx)); }
  __device__ __forceinline__ void add_cov(int slot, int nt) const { rec_add_cov(a, slot, nt); }
  __device__ __forceinline__ void update(int slot, int idx, int length) const { rec_update(a, slot, idx, length); }
  __device__ __forceinline__ void link(int slot, int idx) const { rec_link(a, slot, idx); }
  __device__ __forceinline__ uint32_t dist_now(int slot, int idx) const { return rec_dist_now(a, slot, idx); }
  __device__ __forceinline__ void spf_pair(uint64_t k1, uint64_t k2) const { spf_add_pair(a, k1, k2); }
  __device__ __forceinline__ bool aborted() const { return false; }
  __device__ void ext_flush() {
    // chunk = header {record:32 | part:16 | count:16} + count real-extension k-mers (pair_filter_host.hpp)
    const unsigned long long off = atomicAdd(&a.st->ext_used, (unsigned long long)n_ext + 1);
    a.ext[off] = ((unsigned long long)rec << 32) | ((unsigned long long)(part & 0xffffu) << 16) | n_ext;
    for (uint32_t i = 0; i < n_ext; i++) a.ext[off + 1 + i] = ext_buf[i];
    n_ext = 0;
    part++;
  }
  __device__ __forceinline__ void ext_push(uint64_t kmer) {
    ext_buf[n_ext++] = kmer;
    if (n_ext == S2_EXT) ext_flush();
  }
};

// phase 1, cooperative: the G lanes that share a record look up both keys of every k-mer position
// (positions strided over the lanes, two positions = four probes in flight per lane) and leave the
// answers in the record's shared-memory slot: K planes by atomicOr, the first S2_PARK hits parked.
__device__ __forceinline__ void s2_lookup_group(const StitchArgs& a, RecSlot& RS, uint32_t ls, int n_pos, int lgi, int G) {
  const int k = a.k;
  const uint32_t o4 = ls & 15u;
  for (int p0 = lgi; p0 < n_pos; p0 += 2 * G) {
    uint64_t kk[4], hh[4], got[4];
    int pos[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      pos[u] = p0 + u * G;
      const uint64_t f = kmer_at_t<true>(RS.G.pk, o4 + (pos[u] < n_pos ? pos[u] : 0), k);
      kk[2 * u] = f; kk[2 * u + 1] = revcomp(f, k);
    }
#pragma unroll
    for (int v = 0; v < 4; v++) {
      hh[v] = mix64(kk[v]) & (a.cap - 1);
      got[v] = __ldcg(a.keys + hh[v]);
    }
#pragma unroll
    for (int v = 0; v < 4; v++) {
      const int ps = pos[v >> 1];
      if (ps >= n_pos) continue;
      const int dir = (v & 1) ? 0 : 1;  // even entries hold the forward k-mer = the FORWARD key
      int slot = -1;
      if (kk[v] == KEY_EMPTY) slot = tbl_find(a, kk[v]);
      else {
        while (got[v] != kk[v] && got[v] != KEY_EMPTY) { hh[v] = (hh[v] + 1) & (a.cap - 1); got[v] = __ldcg(a.keys + hh[v]); }
        if (got[v] == kk[v]) slot = (int)hh[v];
      }
      if (slot < 0) continue;
      atomicOr((dir ? RS.L.kf : RS.L.kb) + (ps >> 5), 1u << (ps & 31));
      const uint32_t at = atomicAdd(&RS.L.n_park, 1u);
      if (at < (uint32_t)S2_PARK) {
        // dist[fwdIdx]: facing forward fwdIdx = the read's next base, facing backward fwdIdx = 4
        const int idx = dir ? (int)code_at_t<true>(RS.G.pk, o4 + ps + k) : 4;
        RS.L.pslot[at] = (uint32_t)slot;
        RS.L.pinfo[at] = ((uint32_t)(2 * ps + dir) << 8) | (__ldcg(rec_field(a, slot, REC_DIST + idx)) & 0xffu);
      }
    }
  }
}

__global__ void __launch_bounds__(S2_THREADS, 1) stitch2_kernel(StitchArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char stitch_smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wblock = stitch_smem + (size_t)wib * S2_WARP_BYTES;
  RecSlot* slots = reinterpret_cast<RecSlot*>(wblock);
  WarpScratch* S = reinterpret_cast<WarpScratch*>(wblock + S2_WARP_BYTES - sizeof(WarpScratch));  // direct path + per-warp counters
  // consecutive window slots go to different SMs: slot i belongs to warp i mod (warps of the grid)
  const uint32_t total_warps = gridDim.x * S2_WARPS;
  const uint32_t gw = (uint32_t)wib * gridDim.x + blockIdx.x;
  const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
  StitchState* st = a.st;
  uint32_t next = __ldcg(&st->next), W = __ldcg(&st->W), round = __ldcg(&st->round);
  if (W > a.w_max) W = a.w_max;
  if (lane < SS_COUNT) S->st[lane] = 0;
  if (timer) { st->nb[0] = 0; st->nb[1] = 0; st->nb[2] = 0; }  // first used after the first grid barrier
  __syncwarp();
  WarpCtx c;
  c.S = S; c.pk = a.packed; c.pk_base = 0; c.inv = a.inval; c.inv_base = 0; c.n_stage = 0; c.part = 0; c.rec = 0; c.stamp = 0; c.ls = 0; c.n_pos = 0; c.n_vis = 0;
  StitchArgs aw = a;  // the direct path reserves through `res`: this copy makes it reserve in the writers' table
  aw.res = a.resw;
  DevEnv e{a, a.k, a.j, a.spacer, !a.no_cleaning && a.spf != nullptr, a.ext != nullptr};
  for (int i = 0; i < S2_COUNTERS; i++) e.st[i] = 0;
  e.n_created = 0; e.n_ext = 0; e.part = 0; e.rec = 0; e.stamp_next = 0; e.G = nullptr; e.w5 = 0; e.w4 = 0;
  uint32_t status = ST_DONE;
  unsigned long long need_seen = 0;  // the largest 2 len + 2 this thread has reported
  unsigned fix_it = 0;               // fix-point iterations so far (all threads agree): counter nb[fix_it % 3]
  unsigned n_quiet = 0;

  while (true) {
    const int cur = round & 1, nxt = cur ^ 1;
    const uint32_t nd = __ldcg(&st->nd[cur]);
    const uint32_t room = a.row_end - next;
    const uint32_t n_new = W > nd ? (W - nd < room ? W - nd : room) : 0u;
    const uint32_t n_win = nd + n_new;  // <= w_max <= S2_RPW * warps of the grid
    if (n_win == 0) {
      if (a.row_end < a.n_recs) status = ST_MORE_ROWS;
      break;
    }
    // n_entries / ext_used only move in the execution phase, so this snapshot is the same in every thread
    const unsigned long long entries0 = __ldcg(&st->n_entries), ext0 = __ldcg(&st->ext_used);
    if (timer) { __stcg(&st->nd[nxt], 0u); __stcg(&st->min_w[nxt], RES_FREE); }
    // records per warp this round (a power of two) and lanes per record
    int lg2 = 0;
    while (((uint64_t)total_warps << lg2) < n_win) lg2++;
    const int G = 32 >> lg2;
    const int gid = lane >> (5 - lg2), lgi = lane & (G - 1);
    const uint32_t gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u) << (gid * G));
    const uint32_t my = gw + total_warps * (uint32_t)gid;
    RecSlot& RS = slots[gid];
    // ---- phase 1: reservations, lookups, dry run
    const unsigned long long t0 = gtime_ns();
    const bool have = my < n_win;
    uint32_t rec = 0, ls = 0, len = 0, n_res = 0;
    int n_pos = 0;
    bool simple = false;
    const uint32_t* row = nullptr;
    if (have) {
      rec = my < nd ? __ldcg(a.deferred[cur] + my) : next + (my - nd);
      ls = __ldg(a.seq_start + rec);
      const uint32_t le = __ldg(a.seq_end + rec);
      len = le > ls ? le - ls : 0u;
      n_pos = len >= (uint32_t)a.k ? (int)(len - a.k + 1) : 0;
      row = a.rows + (size_t)(rec - a.row_base) * S2_ROW;
      n_res = __ldg(row);
      simple = n_pos <= S2_POS_CAP && n_res < (uint32_t)S2_ROW;
      if (simple) {
        for (uint32_t i = lgi; i < n_res; i += G) atomicMin(a.res + __ldg(row + 1 + i), rec);
        const uint32_t* fp = a.fplanes + (size_t)(ls >> 5) * FP_STRIDE;
        for (int i = lgi; i < 6 * FP_STRIDE; i += G) RS.G.fp[i] = __ldg(fp + i);
        for (int i = lgi; i < 6; i += G) RS.G.inv[i] = __ldg(a.inval + (ls >> 5) + i);
        for (int i = lgi; i < 13; i += G) RS.G.pk[i] = __ldg(a.packed + (ls >> 4) + i);
        for (int i = lgi; i < (int)(sizeof(LineState) / 4); i += G) reinterpret_cast<uint32_t*>(&RS.L)[i] = 0u;
      }
      if (lgi == 0 && 2ull * len + 2 > need_seen) {
        need_seen = 2ull * len + 2;
        if (need_seen > __ldcg(&st->max_need)) atomicMax(&st->max_need, need_seen);
      }
    }
    __syncwarp();
    if (have && simple) s2_lookup_group(a, RS, ls, n_pos, lgi, G);
    __syncwarp();
    // the record's first lane walks the line read-only: does it change anything a later record can see?
    bool writer = false;
    if (have && simple && lgi == 0) {
      e.G = &RS.G; e.w5 = ls >> 5; e.w4 = ls >> 4;
      writer = !s2_is_quiet(e, RS.L, ls, ls + len);
    }
    writer = __shfl_sync(0xffffffffu, (int)writer, gid * G) != 0;
    if (have && simple && writer)
      for (uint32_t i = lgi; i < n_res; i += G) atomicMin(a.resw + __ldg(row + 1 + i), rec);
    const uint32_t hard_mask = __ballot_sync(0xffffffffu, have && !simple && lgi == 0);
    for (uint32_t hm = hard_mask; hm; hm &= hm - 1) {  // rare: the warp reserves for these records together, as writers
      const int src = __ffs(hm) - 1;
      const uint32_t r_ = __shfl_sync(0xffffffffu, rec, src), ls_ = __shfl_sync(0xffffffffu, ls, src), len_ = __shfl_sync(0xffffffffu, len, src);
      int n_keep = 0;
      line_reservations<0, false>(a, a.packed, ls_, len_, r_, lane, S->reskey, &n_keep);
      line_reservations<0, false>(aw, a.packed, ls_, len_, r_, lane, S->reskey, &n_keep);
    }
    {  // the smallest writer of the round: readers before it can never be blocked
      const uint32_t wmin = __reduce_min_sync(0xffffffffu, (have && lgi == 0 && (writer || !simple)) ? rec : RES_FREE);
      if (lane == 0 && wmin != RES_FREE) atomicMin(&st->min_w[cur], wmin);
    }
    const unsigned long long t1 = gtime_ns();
    grid.sync();
    const unsigned long long t2 = gtime_ns();
    {
      const unsigned long long bound = __ldcg(&st->max_need) * n_win;
      if (entries0 + bound > a.cap / 2) { status = ST_GROW_TABLE; break; }
      if (a.ext && ext0 + 2 * bound + n_win > a.ext_cap) { status = ST_DRAIN_EXT; break; }
    }
    // ---- fix point: a reader that shares a slot with an earlier writer is blocked, and from then on counts as
    //      a writer for the readers after it (it will be looked at again next round, against the new state)
    bool blocked = false;
    for (int it = 0;; it++, fix_it++) {
      bool hit = false;
      if (have && simple && !writer && !blocked)
        for (uint32_t i = lgi; i < n_res; i += G)
          if (__ldcg(a.resw + __ldg(row + 1 + i)) < rec) hit = true;
      const uint32_t hm = __ballot_sync(0xffffffffu, hit);
      if (hm & gmask) {
        blocked = true;
        for (uint32_t i = lgi; i < n_res; i += G) atomicMin(a.resw + __ldg(row + 1 + i), rec);
      }
      if (lane == 0 && hm) atomicAdd(&st->nb[fix_it % 3], 1u);
      if (timer) __stcg(&st->nb[(fix_it + 1) % 3], 0u);
      grid.sync();
      const bool more = __ldcg(&st->nb[fix_it % 3]) != 0;
      if (!more) { fix_it++; break; }
      if (it >= 24) {  // give up: only the readers before the first writer still run this round (always a valid schedule)
        if (have && simple && !writer && rec > __ldcg(&st->min_w[cur])) blocked = true;
        fix_it++;
        break;
      }
    }
    const unsigned long long t2a = gtime_ns();
    // ---- execution: writers need every slot of theirs (no earlier unexecuted record of any kind shares a key);
    //      readers that are not blocked just run.  Everybody drops the reservations it holds.
    bool bad = false;
    if (have && simple)
      for (uint32_t i = lgi; i < n_res; i += G)
        if (__ldcg(a.res + __ldg(row + 1 + i)) != rec) bad = true;
    const uint32_t badm = __ballot_sync(0xffffffffu, bad);
    // release AFTER the whole check: two runs of a line may hash to the same slot, and a slot released by the first
    // one must not look foreign to the second
    if (have && simple)
      for (uint32_t i = lgi; i < n_res; i += G) {
        const uint32_t sl = __ldg(row + 1 + i);
        if (__ldcg(a.res + sl) == rec) __stcg(a.res + sl, RES_FREE);
        if ((writer || blocked) && __ldcg(a.resw + sl) == rec) __stcg(a.resw + sl, RES_FREE);
      }
    bool mine = have && simple && (writer ? !(badm & gmask) : !blocked);  // meaningful in the group's first lane
    for (uint32_t hm = hard_mask; hm; hm &= hm - 1) {
      const int src = __ffs(hm) - 1;
      const uint32_t r_ = __shfl_sync(0xffffffffu, rec, src), ls_ = __shfl_sync(0xffffffffu, ls, src), len_ = __shfl_sync(0xffffffffu, len, src);
      const bool m = line_reservations<1, false>(a, a.packed, ls_, len_, r_, lane, nullptr, nullptr);
      line_reservations<2, false>(a, a.packed, ls_, len_, r_, lane, nullptr, nullptr);
      line_reservations<2, false>(aw, a.packed, ls_, len_, r_, lane, nullptr, nullptr);
      if (lane == src) mine = m;
    }
    const unsigned long long t2b = gtime_ns();
    if (have && simple && mine && lgi == 0) {  // one thread walks the line out of shared memory
      e.rec = rec; e.part = 0; e.n_ext = 0;
      e.stamp_next = (a.rec_base + rec) << 20;
      e.G = &RS.G; e.w5 = ls >> 5; e.w4 = ls >> 4;
      s2_line(e, RS.L, ls, ls + len);
      if (e.want_ext && e.n_ext) e.ext_flush();
      if (!writer) n_quiet++;
    }
    for (uint32_t xm = __ballot_sync(0xffffffffu, have && !simple && mine && lgi == 0); xm; xm &= xm - 1) {
      const int src = __ffs(xm) - 1;
      const uint32_t r_ = __shfl_sync(0xffffffffu, rec, src), ls_ = __shfl_sync(0xffffffffu, ls, src), len_ = __shfl_sync(0xffffffffu, len, src);
      c.rec = r_; c.part = 0; c.n_stage = 0; c.n_vis = 0; c.ls = ls_; c.n_pos = 0;
      c.stamp = (a.rec_base + r_) << 20;
      if (len_) scan_line<false>(a, c, ls_, ls_ + len_, lane);
      if (a.ext && c.n_stage) ext_flush(a, c, lane);
    }
    {  // deferred records go to the next round's list (one atomic per warp)
      const bool defer = have && !mine && lgi == 0;
      const uint32_t dm = __ballot_sync(0xffffffffu, defer);
      if (dm) {
        uint32_t base = 0;
        if (lane == 0) { base = atomicAdd(&st->nd[nxt], (uint32_t)__popc(dm)); S->st[SS_DEFERRED] += (unsigned)__popc(dm); }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (defer) a.deferred[nxt][base + __popc(dm & ((1u << lane) - 1u))] = rec;
      }
      unsigned nc = __reduce_add_sync(0xffffffffu, e.n_created);
      e.n_created = 0;
      if (lane == 0 && nc) atomicAdd(&st->n_entries, (unsigned long long)nc);
    }
    if (timer) { S->st[SS_T_P1A] += t2a - t2; S->st[SS_T_P1B] += t2b - t2a; }
    const unsigned long long t3 = gtime_ns();
    grid.sync();
    if (timer) {
      S->st[SS_T_PHASE1] += t1 - t0; S->st[SS_T_SYNC1] += t2 - t1; S->st[SS_T_PHASE2] += t3 - t2; S->st[SS_T_SYNC2] += gtime_ns() - t3;
      S->st[SS_ROUNDS]++;
    }
    const uint32_t nd_next = __ldcg(&st->nd[nxt]);
    if (nd_next >= n_win) {  // cannot happen: the earliest record of a window always executes
      status = ST_STUCK;
      if (have && lgi == 0) {  // post-mortem for the error message
        atomicAdd(&st->stats[SS_T_P1C], 1ull);
        if (writer || !simple) atomicAdd(&st->stats[SS_T_P2A], 1ull);
        if (blocked) atomicAdd(&st->stats[SS_T_P1B], 1ull);
        atomicMin(&st->max_need, ((unsigned long long)rec << 8) | (writer ? 1u : 0u) | (blocked ? 2u : 0u) | (simple ? 4u : 0u) | ((badm & gmask) ? 8u : 0u));
      }
      break;
    }
    if (nd_next * a.shrink_den > n_win) W = W / 2 > a.w_min ? W / 2 : a.w_min;
    else if (nd_next * a.grow_den < n_win && n_win >= W) W = W * 2 < a.w_max ? W * 2 : a.w_max;
    next += n_new;
    round++;
  }
  // counters: thread path (registers) + direct path / scheduling (shared) -> device state
  {
    const int map[S2_COUNTERS] = {SS_JCHECK, SS_NOJUNC, SS_PROCESSED, SS_SKIPPED, SS_NOERR, SS_UNAMBIG};
#pragma unroll
    for (int i = 0; i < S2_COUNTERS; i++) {
      // 32-bit per-thread counters, 64-bit sum
      unsigned long long v = e.st[i];
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) S->st[map[i]] += v;
    }
  }
  __syncwarp();
  if (lane < SS_COUNT && S->st[lane]) atomicAdd(&st->stats[lane], S->st[lane]);
  {
    const unsigned nq = __reduce_add_sync(0xffffffffu, n_quiet);
    if (lane == 0 && nq) atomicAdd(&st->quiet_runs, (unsigned long long)nq);
  }
  if (timer) {
    __stcg(&st->next, next); __stcg(&st->W, W); __stcg(&st->round, round); __stcg(&st->status, status);
  }
}

}  // namespace faucet
