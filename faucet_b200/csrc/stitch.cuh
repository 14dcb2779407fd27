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
//   phase 1  every record of the window (the deferred ones + the next new ones, so every unexecuted
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
// Latency.  The junction table does not change during phase 1, and during phase 2 it only changes on
// keys of executing records, which are private to them.  So a warp can look up EVERY half-step of its
// line (table slot, stored skip distance, flag byte) in phase 1, with all loads in flight at once,
// park the answers in shared memory across the grid barrier, and walk the line in phase 2 without
// waiting on memory: the walk only re-reads what the line itself changed (tracked exactly).  Record
// updates are commutative (dist = max, cov = count, linked = or) and are issued as fire-and-forget
// atomics; what later records READ is ordered by the rounds.
//
// Epochs.  After the first few x coverage almost every record is QUIET: it creates no
// junction, raises no stored distance, sets no new link -- it only counts coverage, which nothing reads.
// The stream is therefore cut into epochs.  An epoch either runs entirely through the ordered kernel above
// (dense phase), or:
//   classify  every record walks READ-ONLY against the table as it stood when the epoch began (T0), fully
//             parallel, no ordering (stitch_dry_kernel); the ones that would write form the exact set E;
//   execute   E runs through the ordered kernel, in stream order, on the live table; every real write
//             (creation / raised distance) marks dirty[minimizer slot of that k-mer] = min(record);
//   verify    a record outside E whose line touches a slot written by an EARLIER record may have seen a
//             stale T0: it joins E, the table is restored to T0 and E runs again -- until nothing joins;
//   apply     the records outside E (quiet under T0, and T0 is what they would have seen) add their
//             coverage counts, again fully parallel, reading T0 (stitch_dry_kernel).  When the scan has no
//             pair filters to feed, classify applies at once and the few records that verify adds to E are
//             retracted (the same walk, counts subtracted) before E runs again.
// By induction over the stream every record outside E behaves exactly as in the sequential run (its keys are
// untouched before its turn), so E -- executed in order from T0 -- sees exactly the sequential state.
//
// The junction map is an open-addressing table in HBM: key = oriented k-mer (ReadKmer::getKmer),
// 64-byte record of u32 fields (dist[5], linked mask, cov[4] counts), 8-byte creation stamp =
// (record index << STAMP_SHIFT | n-th creation in that record).  Sorting by stamp gives the reference's creation
// order; cov saturates at 255 when the map is collected (utils/Junction.cpp:59-67).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer.cuh"
#include "load.cuh"

namespace faucet {

namespace cg = cooperative_groups;

#ifndef FAUCET_STITCH_THREADS
#define FAUCET_STITCH_THREADS 256
#endif
constexpr int STITCH_THREADS = FAUCET_STITCH_THREADS;
constexpr int STITCH_WARPS = STITCH_THREADS / 32;
constexpr unsigned long long KEY_EMPTY = ~0ull;
constexpr uint32_t RES_FREE = 0xffffffffu;
constexpr int EXT_STAGE = 32;   // staged real-extension k-mers per warp before a chunk is flushed
#ifndef FAUCET_POS_CAP
#define FAUCET_POS_CAP 256
#endif
constexpr int POS_CAP = FAUCET_POS_CAP;  // k-mer positions per line served from shared memory (longer lines: direct path)
constexpr int PK_WORDS = (15 + POS_CAP + 32 + 15) / 16 + 1, INV_WORDS = (31 + POS_CAP + 32 + 31) / 32 + 1;  // plane words of such a line
constexpr int RES_CAP = 96;     // reservation slots per line kept in shared memory
constexpr int VIS_CAP = 32;     // junction slots a line touched (beyond it every skip distance is re-read)
constexpr int REC_WORDS = 16;   // u32 per junction record
enum { REC_DIST = 0, REC_LINK = 5, REC_COV = 6 };
constexpr int STAMP_SHIFT = 26; // creation stamp = record index << 26 | n-th creation of that record (a line is < 2^24 bytes)
constexpr unsigned long long STAMP_LOW = (1ull << STAMP_SHIFT) - 1ull;
constexpr int LAND_CAP = 160;   // junctions one line may land on in the read-only walk (more: the ordered kernel takes it)
constexpr int DRY_THREADS = 256;
constexpr int DRY_WARPS = DRY_THREADS / 32;
enum { DRY_CLASSIFY = 0, DRY_APPLY = 1, DRY_CLASSIFY_APPLY = 2, DRY_RETRACT = 3, DRY_RECHECK = 4 };
// in_exact[] of a classify epoch.  Members of the exact set: EX_*.
enum { EX_QUIET = 0,      // quiet under T0 and no earlier write under its slots: T0 is what it would have seen
       EX_MEMBER = 1,     // exact set; never committed
       EX_COMMITTED = 2,  // exact set; had committed under T0 at classify time: to be retracted
       EX_RETRACTED = 3,  // exact set; retracted
       EX_SETTLED = 4,    // earlier writes under its slots but none later, and quiet on the LIVE table (= the state it would have seen)
       EX_RECHECK = 5 };  // earlier writes, none later: to be walked on the live table
constexpr uint32_t ROW_SLOT_MASK = 0xffffffu;  // a listed reservation slot = slot | first position of its run << 24 (res_log2 <= 24)
constexpr int ROW_WORDS = 32;   // reservation row of a record: [0] = count (255: recompute), [1..31] = slots


enum { ST_DONE = 0, ST_GROW_TABLE = 1, ST_DRAIN_EXT = 2, ST_MORE_ROWS = 3, ST_STUCK = 4 };
enum { SS_JCHECK = 0, SS_NOJUNC, SS_PROCESSED, SS_SKIPPED, SS_NOERR, SS_UNAMBIG, SS_ROUNDS, SS_DEFERRED,
       SS_T_PHASE1, SS_T_SYNC1, SS_T_PHASE2, SS_T_SYNC2, SS_T_P1A, SS_T_P1B, SS_T_P1C, SS_T_P2A,  // SS_T_*: ns seen by warp 0 of the grid
       SS_WRITERS,    // records the ordered kernel ran that wrote something a later record can see (or a new link)
       SS_NONQUIET,   // records the read-only walk handed to the ordered kernel
       SS_TAINTED,    // records the verify step added to the exact set
       SS_DRY_ERROR,  // apply found a record that is not quiet (must stay 0)
       SS_COUNT };
constexpr int SS_WALK = 6;      // the first six are the reference's scan counters; a read-only walk holds them per record

struct StitchState {             // device-resident; survives kernel launches and batches
  unsigned long long n_entries;  // occupied slots of the junction table
  unsigned long long stats[SS_COUNT];
  unsigned long long ext_used;   // u64 words used in the ext buffer
  unsigned long long max_need;   // upper bound of the junction events ONE record can cause (2 x longest line + 2)
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
  const uint8_t* flags;           // scan_flags output, one byte per k-mer start
  const uint32_t* seq_start;
  const uint32_t* seq_end;
  uint32_t n_recs;               // ordered kernel: it runs the entries [st->next, n_recs) ...
  const uint32_t* list;          // ... of this ascending list of record indices (NULL: the record indices themselves)
  unsigned long long rec_base;   // global index of record 0 of this batch ...
  const uint32_t* gid;           // ... or, per record of the batch, its global index (the gathered exact set of a sharded epoch)
  int k, j, spacer;
  int no_cleaning, paired;
  unsigned long long* keys;      // cap + 1 entries (the last one is the home of the KEY_EMPTY k-mer)
  uint32_t* recs;                // REC_WORDS u32 per slot
  unsigned long long* stamps;
  unsigned long long cap;        // power of two, < 2^31
  uint32_t* res;                 // reservation table
  uint32_t res_mask;
  uint32_t* jslot;               // one bit per reservation slot: some junction's k-mer has that minimizer (never cleared during
                                 // a scan, so it may err towards set): runs of a line with a clear bit need no table lookup
  uint32_t* dirty;               // same slots: smallest record index that wrote a junction whose k-mer has that minimizer
                                 // (ordered kernel marks, verify reads; NULL outside classify epochs) ...
  uint32_t* dirty_max;           // ... and 1 + the largest such index
  // read-only walk (stitch_dry_kernel / stitch_verify_kernel): keys/recs above are the epoch's T0 snapshot
  uint32_t* cov_out;             // records of the LIVE table: apply adds the coverage counts here (same slots as T0)
  uint32_t cov_stride, cov_off;  // ... laid out as cov_out[slot * cov_stride + cov_off + nt]: the records themselves
                                 // (REC_WORDS, REC_COV) or a shard's own count array (4, 0; sharded epochs, multi.cuh)
  uint32_t* cov_out2;            // retract: the snapshot's records too (a later restore must not bring the counts back)
  StitchState* st2;              // retract: the snapshot's counters too
  uint8_t* in_exact;             // per record of the batch: 0 = quiet (applied, if classify applies), 1 = exact set (never
                                 // applied), 2 = joined the exact set after it was applied, 3 = the same, retracted
  uint8_t taint_mark;            // what joins the exact set after classify gets: EX_MEMBER, or EX_COMMITTED when classify applied
  uint8_t want_flag, flag_after; // apply / retract / recheck take the records with in_exact == want_flag; retract leaves flag_after
  uint8_t recheck;               // verify: records with earlier but no later writes are rechecked on the live table (else they join)
  uint8_t lazy;                  // read-only walk: look keys up as the walk reaches them instead of parking the whole line first
  uint8_t rows_ready;            // classify: a.rows already holds the reservation rows (stitch_rows_kernel ran ahead)
  uint32_t* rows;                // classify writes, verify and the later walks read: ROW_WORDS u32 per record of the epoch
  uint32_t rows_base;            // record of row 0
  uint32_t r_begin, r_end;       // the epoch
  int dry_mode;
  uint32_t* deferred[2];         // w_max entries each
  StitchState* st;
  const unsigned int* special;   // "the KEY_EMPTY k-mer is present" of the table behind keys (the live state's or the snapshot's)
  uint32_t* spf;                 // short pair filter on the device (NULL: none)
  unsigned long long spf_mask;
  int spf_nh;
  unsigned long long* ext;       // real-extension chunks for the host-side long pair filter (NULL: none)
  unsigned long long ext_cap;
  uint32_t w_min, w_max;         // w_max <= warps of the grid: one record per warp per round
  uint32_t shrink_den, grow_den; // window halves when deferred > win/shrink_den, doubles when < win/grow_den
};

struct WarpScratch {             // shared memory of one warp; filled in phase 1, consumed in phase 2
  int slot[2 * POS_CAP];              // table slot of the key at half-step 2*pos+dir, -1 = not a junction
  uint8_t hop[2 * POS_CAP];           // stored dist[fwdIdx] of that junction when the round started ...
  uint8_t dback[2 * POS_CAP];         // ... its dist[backIdx] ...
  uint8_t lnk[2 * POS_CAP];           // ... and its link mask (the read-only walk checks them instead of writing)
  uint8_t flag[POS_CAP];              // scan_flags byte of the position
  uint8_t plist[POS_CAP];             // the positions whose keys have to be looked up (prefetch_line)
  uint32_t pmask[POS_CAP / 32];       // the same as a bit mask
  uint32_t pk[PK_WORDS];              // the line's words of the 2-bit plane, from word ls>>4
  uint32_t inv[INV_WORDS];            // the line's words of the validity plane, from word ls>>5
  uint32_t reskey[RES_CAP];           // reservation slots of the line
  int visited[VIS_CAP];               // slots of the junctions this line has touched
  unsigned long long stage[EXT_STAGE];
  unsigned long long st[SS_COUNT];    // counters of this warp (lane 0), flushed when the kernel ends
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// plane words either through the read-only path (global planes) or from the warp's shared copy
template <bool SH>
__device__ __forceinline__ uint32_t ldw(const uint32_t* p) { return SH ? *p : __ldg(p); }
template <bool SH>
__device__ __forceinline__ uint64_t kmer_at_t(const uint32_t* packed, uint32_t p, int k) {
  const uint32_t w = p >> 4, o = 2 * (p & 15);
  const uint64_t hi = ((uint64_t)ldw<SH>(packed + w) << 32) | ldw<SH>(packed + w + 1);
  const uint64_t lo = (uint64_t)ldw<SH>(packed + w + 2) << 32;
  const uint64_t x = o ? ((hi << o) | (lo >> (64 - o))) : hi;
  return x >> (64 - 2 * k);
}
template <bool SH>
__device__ __forceinline__ uint32_t code_at_t(const uint32_t* packed, uint32_t p) {
  return (ldw<SH>(packed + (p >> 4)) >> (30 - 2 * (p & 15))) & 3u;
}

__device__ __forceinline__ uint32_t flag_byte(const StitchArgs& a, uint32_t p) { return a.flags[p]; }

// h(canonical s-mer), s <= 16; x = the s-mer, first base in the high bits of its 2s-bit value
__device__ __forceinline__ uint32_t smer_canon_hash(uint32_t x, int s) {
  uint32_t r = __brev(x << (32 - 2 * s));                       // reversed bit order, low-aligned
  r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);      // un-swap inside the 2-bit groups
  r ^= (s == 16 ? 0xaaaaaaaau : (0xaaaaaaaau & ((1u << (2 * s)) - 1u)));
  return mix32(x < r ? x : r);
}
// the s-mer starting at byte offset q
template <bool SH>
__device__ __forceinline__ uint32_t smer_hash(const uint32_t* packed, uint32_t q, int s) {
  uint32_t w0 = ldw<SH>(packed + (q >> 4)), w1 = ldw<SH>(packed + (q >> 4) + 1);
  return smer_canon_hash(__funnelshift_l(w1, w0, 2 * (q & 15)) >> (32 - 2 * s), s);
}
// reservation slot of ONE k-mer given by value (either orientation gives the same slot): the minimum over
// its k-s+1 <= 17 s-mers, one per lane.  Equals what line_reservations finds for that k-mer inside a line.
__device__ __forceinline__ uint32_t kmer_res_slot(uint64_t key, int k, uint32_t res_mask, int lane) {
  const int s = k < 16 ? k : 16, w = k - s + 1;
  uint32_t h = 0xffffffffu;
  if (lane < w) {
    const uint32_t x = (uint32_t)(key >> (2 * (k - s - lane))) & (s == 16 ? 0xffffffffu : ((1u << (2 * s)) - 1u));
    h = smer_canon_hash(x, s);
  }
  return __reduce_min_sync(0xffffffffu, h) & res_mask;
}

// value at position lane+d of the 64-entry sequence (x0 = entries 0..31, x1 = entries 32..63)
__device__ __forceinline__ uint32_t shift64(uint32_t x0, uint32_t x1, int d, int lane) {
  const int src = (lane + d) & 31;
  const uint32_t a = __shfl_sync(0xffffffffu, x0, src), b = __shfl_sync(0xffffffffu, x1, src);
  return lane + d < 32 ? a : b;
}

// One warp walks the minimizers of line [ls, ls+len): one reservation slot per run of equal values.
// MODE 0: reserve with atomicMin and remember the slots in `keep` (*n_keep > keep_cap: did not fit);
// MODE 1: check (false on a foreign reservation); MODE 2: release own reservations;
// MODE 3: only list the slots in `keep` (the reservation rows of stitch2.cuh).
template <int MODE, bool SH>
__device__ bool line_reservations(const StitchArgs& a, const uint32_t* packed, uint32_t ls, uint32_t len, uint32_t rec,
                                  int lane, uint32_t* keep, int* n_keep, int keep_cap = RES_CAP) {
  const int k = a.k, s = k < 16 ? k : 16, w = k - s + 1;
  if (MODE == 0 || MODE == 3) *n_keep = 0;
  if (len < (uint32_t)k) return true;
  const uint32_t nk = len - k + 1, ns = len - s + 1;
  int lg = 0;
  while ((2 << lg) <= w) lg++;  // 2^lg <= w < 2^(lg+1)
  bool ok = true;
  int nkeep = 0;
  uint32_t g0 = lane < (int)ns ? smer_hash<SH>(packed, ls + lane, s) : 0xffffffffu;
  uint32_t prev_last = 0;  // minimizer of the last position of the previous chunk
  for (uint32_t base = 0; base < nk; base += 32) {
    const uint32_t q1 = base + 32 + lane;
    const uint32_t g1 = q1 < ns ? smer_hash<SH>(packed, ls + q1, s) : 0xffffffffu;
    // sliding minimum of width w by doubling: m_(2d)[p] = min(m_d[p], m_d[p+d])
    uint32_t x0 = g0, x1 = g1;
    for (int d = 1; d < (1 << lg); d <<= 1) {
      const uint32_t y0 = shift64(x0, x1, d, lane);
      uint32_t y1 = __shfl_sync(0xffffffffu, x1, (lane + d) & 31);
      if (lane + d >= 32) y1 = 0xffffffffu;
      x0 = x0 < y0 ? x0 : y0;
      x1 = x1 < y1 ? x1 : y1;
    }
    uint32_t m = x0;
    if (w > (1 << lg)) {
      const uint32_t y = shift64(x0, x1, w - (1 << lg), lane);
      m = m < y ? m : y;
    }
    uint32_t left = __shfl_up_sync(0xffffffffu, m, 1);
    if (lane == 0) left = prev_last;
    const bool active = base + lane < nk && (base + lane == 0 || m != left);
    uint32_t* slot = a.res + (m & a.res_mask);
    if (MODE == 0 || MODE == 3) {
      if (MODE == 0 && active) atomicMin(slot, rec);
      const uint32_t b = __ballot_sync(0xffffffffu, active);
      const int at = nkeep + __popc(b & ((1u << lane) - 1u));
      if (active && at < keep_cap) keep[at] = (m & a.res_mask) | ((base + lane < 255u ? base + lane : 255u) << 24);
      nkeep += __popc(b);
    }
    if (MODE == 1 && active && __ldcg(slot) != rec) ok = false;
    if (MODE == 2 && active && __ldcg(slot) == rec) __stcg(slot, RES_FREE);
    prev_last = __shfl_sync(0xffffffffu, m, 31);
    g0 = g1;
  }
  if (MODE == 0 || MODE == 3) *n_keep = nkeep;
  return MODE == 1 ? __all_sync(0xffffffffu, ok) : true;
}

// ---- junction table -----------------------------------------------------------------------------
__device__ __forceinline__ int tbl_find(const StitchArgs& a, uint64_t key) {
  if (key == KEY_EMPTY) return __ldcg(a.special) ? (int)a.cap : -1;
  uint64_t h = mix64(key) & (a.cap - 1);
  while (true) {
    unsigned long long kk = __ldcg(a.keys + h);
    if (kk == key) return (int)h;
    if (kk == KEY_EMPTY) return -1;
    h = (h + 1) & (a.cap - 1);
  }
}
// one thread; returns the slot and whether the key was created (JunctionMap::createJunction, zeroed record)
__device__ __forceinline__ int tbl_insert(const StitchArgs& a, uint64_t key, bool* created) {
  if (key == KEY_EMPTY) {
    *created = atomicExch(&a.st->special, 1u) == 0u;
    if (*created) { a.keys[a.cap] = key; atomicAdd(&a.st->n_entries, 1ull); }
    return (int)a.cap;
  }
  uint64_t h = mix64(key) & (a.cap - 1);
  while (true) {
    unsigned long long old = atomicCAS(a.keys + h, KEY_EMPTY, (unsigned long long)key);
    if (old == KEY_EMPTY) { *created = true; atomicAdd(&a.st->n_entries, 1ull); return (int)h; }
    if (old == key) { *created = false; return (int)h; }
    h = (h + 1) & (a.cap - 1);
  }
}
// the same without touching the shared entry counter (the caller adds its creations up once per round)
__device__ __forceinline__ int tbl_insert_nc(const StitchArgs& a, uint64_t key, bool* created) {
  if (key == KEY_EMPTY) {
    *created = atomicExch(&a.st->special, 1u) == 0u;
    if (*created) a.keys[a.cap] = key;
    return (int)a.cap;
  }
  uint64_t h = mix64(key) & (a.cap - 1);
  while (true) {
    unsigned long long old = atomicCAS(a.keys + h, KEY_EMPTY, (unsigned long long)key);
    if (old == KEY_EMPTY) { *created = true; return (int)h; }
    if (old == key) { *created = false; return (int)h; }
    h = (h + 1) & (a.cap - 1);
  }
}
__device__ __forceinline__ uint32_t* rec_field(const StitchArgs& a, int slot, int f) {
  return a.recs + (size_t)slot * REC_WORDS + f;
}
// Junction::update: dist = max(dist, (unsigned char)length)  (utils/Junction.cpp:69-71; u8 narrowing at the call)
__device__ __forceinline__ void rec_update(const StitchArgs& a, int slot, int idx, int length) {
  atomicMax(rec_field(a, slot, REC_DIST + idx), (uint32_t)length & 0xffu);
}
__device__ __forceinline__ void rec_link(const StitchArgs& a, int slot, int idx) {
  atomicOr(rec_field(a, slot, REC_LINK), 1u << idx);
}
__device__ __forceinline__ void rec_add_cov(const StitchArgs& a, int slot, int nt) {
  atomicAdd(rec_field(a, slot, REC_COV + nt), 1u);
}
// the current dist[idx], ordered after this thread's own atomics on the same word
__device__ __forceinline__ uint32_t rec_dist_now(const StitchArgs& a, int slot, int idx) {
  return atomicMax(rec_field(a, slot, REC_DIST + idx), 0u);
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

__device__ __forceinline__ uint32_t global_rec(const StitchArgs& a, uint32_t rec) {
  return a.gid ? __ldg(a.gid + rec) : (uint32_t)a.rec_base + rec;
}
__device__ __forceinline__ unsigned long long global_stamp(const StitchArgs& a, uint32_t rec) {
  return (a.gid ? (unsigned long long)__ldg(a.gid + rec) : a.rec_base + rec) << STAMP_SHIFT;
}

struct WarpCtx {
  unsigned long long stamp;         // next creation stamp of the current record
  WarpScratch* S;                   // lookups parked by phase 1 (ordered kernel, lines that fit POS_CAP)
  unsigned long long* stage;        // EXT_STAGE staged real-extension k-mers
  unsigned long long* cnt;          // where the walk's scan counters go (lane 0): the warp's totals, or -- in the
                                    // read-only walk -- the record's own, added to the totals only if it commits
  uint32_t ls;                      // byte offset of the current line
  const uint32_t* pk;               // 2-bit plane as the walk sees it (shared copy or global) ...
  uint32_t pk_base;                 // ... and the byte offset of its word 0
  const uint32_t* inv;
  uint32_t inv_base;
  int n_pos;                        // k-mer positions of the line held in S (0: direct path)
  int n_vis;                        // entries of S->visited (> VIS_CAP: overflowed)
  uint32_t n_stage, part, rec;
  uint32_t grec;                    // global index of the record (creation stamps, epoch marks)
  bool wrote;                       // ordered kernel: the record changed something a later record can see, or set a link
  // read-only walk
  uint32_t* land_slot;              // junction slots the line landed on ...
  uint8_t* land_nt;                 // ... and the real next nucleotide there (coverage index)
  int n_land;
  bool emit;                        // side effects on (pair filter adds, extension lists): apply, never classify
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

// the result list of one valid sub-read (scan_forward's `result`, src/ReadScanner.cpp:146) as far as the pair filters need it
struct OutList {
  uint64_t v_prev1 = 0, v_prev2 = 0, fb_ext = 0, lf_ext = 0;  // v[n-1], v[n-2]; first backward / last forward extension
  uint32_t n_out = 0;
  int rev_pos = 0, for_pos = 0;
  bool have_fb = false, have_lf = false;
};
// dir: 0 / 1 = the junction faces backward / forward, -1 = the fake mid-read junction (:198-201)
__device__ __forceinline__ void out_push(const StitchArgs& a, WarpCtx& c, OutList& o, uint64_t real_ext, int dir, int pos,
                                         bool pairs, bool want_ext, int lane) {
  if (dir == 0) { if (!o.have_fb) { o.have_fb = true; o.fb_ext = real_ext; o.rev_pos = pos; } }
  else if (dir == 1) { if (!o.have_lf) { o.have_lf = true; o.for_pos = pos; } o.lf_ext = real_ext; }
  // pairs (v[i], v[i+2]) once the list has more than two entries; a list that ends with exactly two
  // entries is handled by out_finish (:208-225)
  if (pairs && o.n_out >= 2 && lane == 0) spf_add_pair(a, o.v_prev2, real_ext);
  o.v_prev2 = o.v_prev1; o.v_prev1 = real_ext; o.n_out++;
  if (want_ext) {
    if (lane == 0) c.stage[c.n_stage] = real_ext;
    c.n_stage++;
    __syncwarp();
    if (c.n_stage == EXT_STAGE - 1) ext_flush(a, c, lane);
  }
}
__device__ __forceinline__ void out_finish(const StitchArgs& a, const OutList& o, bool pairs, int lane) {
  if (pairs && o.n_out == 2 && lane == 0) {  // :208-218
    if (o.have_fb && o.have_lf && !(o.rev_pos > o.for_pos)) spf_add_pair(a, o.fb_ext, o.lf_ext);
    if (o.have_fb != o.have_lf) spf_add_pair(a, o.v_prev2, o.v_prev1);
  }
}

// forward k-mer at position `pos` of the line staged in c.S
__device__ __forceinline__ uint64_t line_kmer(const StitchArgs& a, const WarpCtx& c, int pos) {
  return kmer_at_t<true>(c.S->pk, (c.ls & 15u) + pos, a.k);
}
// has this line already touched `slot`?  (then the skip distance parked in phase 1 may be stale)
__device__ __forceinline__ bool line_visited(const WarpCtx& c, int slot, int lane) {
  if (c.n_vis > VIS_CAP) return true;
  const bool hit = lane < c.n_vis && c.S->visited[lane] == slot;
  return __any_sync(0xffffffffu, hit);
}
__device__ __forceinline__ void line_visit(WarpCtx& c, int slot, int lane) {
  if (c.n_vis < VIS_CAP) { if (lane == 0) c.S->visited[c.n_vis] = slot; }
  c.n_vis++;
  __syncwarp();
}
// a key this line just created: every other half-step of the line with that key must now see it
__device__ __forceinline__ void line_publish(const StitchArgs& a, WarpCtx& c, uint64_t key, int slot, int lane) {
  for (int pos = lane; pos < c.n_pos; pos += 32) {
    const uint64_t f = line_kmer(a, c, pos);
    if (f == key) c.S->slot[2 * pos + 1] = slot;
    if (revcomp(f, a.k) == key) c.S->slot[2 * pos] = slot;
  }
  __syncwarp();
}
// epoch bookkeeping of the ordered kernel: a junction with this key was created or had a distance raised by `rec`
// (and `created`: the slot now holds a junction).  `rec` is the GLOBAL record index (u32: a scan of < 2^32 - 2 records):
// the exact set of a sharded epoch spans the shards of several GPUs.
__device__ __forceinline__ void mark_written(const StitchArgs& a, uint64_t key, uint32_t rec, bool created, bool visible, int lane) {
  if (!created && !(visible && a.dirty)) return;
  const uint32_t slot = kmer_res_slot(key, a.k, a.res_mask, lane);
  if (lane == 0) {
    if (created) atomicOr(a.jslot + (slot >> 5), 1u << (slot & 31u));
    if (visible && a.dirty) { atomicMin(a.dirty + slot, rec); atomicMax(a.dirty_max + slot, rec + 1u); }
  }
}

// one lane: the junction updates of a landing.  Returns bit 0: this junction changed visibly (created / distance
// raised), bit 1: the previous junction's distance was raised, bit 2: anything changed (links included).
__device__ __forceinline__ uint32_t landing_updates(const StitchArgs& a, int slot, int real, int back_idx, bool created,
                                                    bool have_last, int last_slot, int last_fwd_idx, int d, int first_len) {
  uint32_t wr = created ? 5u : 0u;
  rec_add_cov(a, slot, real);
  if (have_last) {  // directLinkJunctions (utils/JunctionMap.cpp:551-561)
    const uint32_t dv = (uint32_t)d & 0xffu;
    const uint32_t o1 = atomicMax(rec_field(a, last_slot, REC_DIST + last_fwd_idx), dv);
    const uint32_t o2 = atomicOr(rec_field(a, last_slot, REC_LINK), 1u << last_fwd_idx);
    const uint32_t o3 = atomicMax(rec_field(a, slot, REC_DIST + back_idx), dv);
    const uint32_t o4 = atomicOr(rec_field(a, slot, REC_LINK), 1u << back_idx);
    if (o1 < dv) wr |= 6u;
    if (o3 < dv) wr |= 5u;
    if (!((o2 >> last_fwd_idx) & 1u) || !((o4 >> back_idx) & 1u)) wr |= 4u;
  } else {
    const uint32_t dv = (uint32_t)first_len & 0xffu;
    if (atomicMax(rec_field(a, slot, REC_DIST + back_idx), dv) < dv) wr |= 5u;
  }
  return wr;
}

// What the walk knows about a junction's record without asking memory again: dist[fwdIdx], dist[backIdx] and the link
// mask as parked by prefetch_line (or zeros for a junction this line just created).  The keys of a running record are
// private to it, so these stay true until the line itself changes them -- which it tracks -- and the updates of a
// landing can be issued as fire-and-forget reductions, only where they change something.
struct KnownRec {
  uint32_t d_fwd, d_back, link;
  bool ok;  // false: the line touched this junction before (the parked values may be stale): ask memory
};

// scan_forward (src/ReadScanner.cpp:112-231) on the valid sub-read at byte offset s0, `len` bases.
// FAST: the line's lookups were parked in shared memory in phase 1 (c.S, index = s0 - c.ls + pos).
template <bool FAST>
__device__ void scan_forward(const StitchArgs& a, WarpCtx& c, uint32_t s0, int len, int lane) {
  const int k = a.k, j = a.j;
  const uint64_t mask = kmer_mask(k);
  const int rel = (int)(s0 - c.ls);
  const int tested_end = 2 * len - 2 * k + 1 - 2 * j;  // distToEnd > 2j  <=>  tp < tested_end
  int tp = 2 * j + 1, last_junc_pos = 0;
  bool have_last = false;
  int last_tp = 0, last_fwd_idx = 0, last_slot = -1;
  uint64_t last_key = 0;
  KnownRec last_rec = {0, 0, 0, false};
  OutList o;
  const bool pairs = !a.no_cleaning && a.spf != nullptr;
  const bool want_ext = a.ext != nullptr;
  uint32_t n_jcheck = 0, n_processed = 0, n_skipped = 0;  // warp-uniform; added to the record's counters at the end

  while (true) {
    // ---- find_next_junction (:61-86): 32 half-steps per warp iteration
    bool found = false;
    int slot = -1;
    if (FAST && have_last && tp < tested_end) {  // after a skip the cursor normally stands on the next known junction
      slot = c.S->slot[2 * rel + tp];
      found = slot >= 0;
    }
    while (!found && tp < tested_end) {
      const int t = tp + lane;
      const bool active = t < tested_end;
      bool known = false, spc = false, tst = false;
      uint32_t cnt = 0;
      int sl = -1;
      if (active) {
        const int pos = t >> 1, dir = t & 1;
        uint32_t f;
        if (FAST) {
          sl = c.S->slot[2 * (rel + pos) + dir];
          f = c.S->flag[rel + pos];
        } else {
          const uint64_t fwd = kmer_at_t<false>(a.packed, s0 + pos, k);
          sl = tbl_find(a, dir ? fwd : revcomp(fwd, k));
          f = flag_byte(a, s0 + pos);
        }
        known = sl >= 0;
        spc = t - last_junc_pos >= 2 * a.spacer - 1;
        cnt = dir ? (f >> 3) & 3u : (f >> 5) & 3u;
        tst = dir ? (f & 2u) != 0 : (f & 4u) != 0;
      }
      const uint32_t am = __ballot_sync(0xffffffffu, active);
      const uint32_t hb = __ballot_sync(0xffffffffu, active && (known || spc || tst));
      const int hit = hb ? __ffs(hb) - 1 : 32;
      const uint32_t upto = hit < 32 ? (hit == 31 ? 0xffffffffu : ((2u << hit) - 1u)) : am;
      // NbJCheckKmer (:46): every half-step that reached testForJunction, the hit one included
      uint32_t jc = (active && ((upto >> lane) & 1u) && !known && !spc) ? cnt : 0u;
      n_jcheck += __reduce_add_sync(0xffffffffu, jc);
      if (hb) {
        n_processed += hit;
        tp += hit;
        slot = __shfl_sync(0xffffffffu, sl, hit);
        found = true;
        break;
      }
      n_processed += __popc(am);
      tp += 32;
    }
    if (!found) break;
    // ---- the junction at half-step tp (:134-192)
    const int pos = tp >> 1, dir = tp & 1;
    const bool known = slot >= 0;
    uint64_t key = 0;  // needed to create the junction, to mark a write (epochs) and for the pair filters
    if (!known || a.dirty || pairs || want_ext) {
      const uint64_t fwd = FAST ? line_kmer(a, c, rel + pos) : kmer_at_t<false>(a.packed, s0 + pos, k);
      key = dir ? fwd : revcomp(fwd, k);
    }
    const int real = dir ? (int)code_at_t<FAST>(c.pk, s0 + pos + k - c.pk_base)
                         : (int)nt_comp(code_at_t<FAST>(c.pk, s0 + pos - 1 - c.pk_base));
    const int fwd_idx = dir ? real : 4, back_idx = dir ? 4 : real;  // getExtensionIndex (utils/ReadKmer.cpp:95-100)
    bool created = false;
    if (!known) {
      if (lane == 0) {
        slot = tbl_insert(a, key, &created);
        if (created) a.stamps[slot] = c.stamp++;
      }
      slot = __shfl_sync(0xffffffffu, slot, 0);
      created = __shfl_sync(0xffffffffu, (int)created, 0) != 0;
      c.stamp = __shfl_sync(0xffffffffu, c.stamp, 0);
      if (FAST) line_publish(a, c, key, slot, lane);
    }
    const bool seen = line_visited(c, slot, lane);
    KnownRec cur = {0, 0, 0, false};
    if (created) cur.ok = true;  // a zeroed record
    else if (FAST && known && !seen) {
      const int hs = 2 * (rel + pos) + dir;
      cur.d_fwd = c.S->hop[hs]; cur.d_back = c.S->dback[hs]; cur.link = c.S->lnk[hs]; cur.ok = true;
    }
    uint32_t wr = 0;
    int dist;
    if (cur.ok && (!have_last || last_rec.ok)) {
      // every value is known: reductions only where something changes (warp-uniform decisions, lane 0 issues)
      wr = created ? 5u : 0u;
      if (lane == 0) rec_add_cov(a, slot, real);
      if (have_last) {  // directLinkJunctions (utils/JunctionMap.cpp:551-561)
        const uint32_t dv = (uint32_t)(tp - last_tp) & 0xffu;
        if (dv > last_rec.d_fwd) { wr |= 6u; if (lane == 0) rec_update(a, last_slot, last_fwd_idx, (int)dv); }
        if (!((last_rec.link >> last_fwd_idx) & 1u)) { wr |= 4u; if (lane == 0) rec_link(a, last_slot, last_fwd_idx); }
        if (dv > cur.d_back) { wr |= 5u; cur.d_back = dv; if (lane == 0) rec_update(a, slot, back_idx, (int)dv); }
        if (!((cur.link >> back_idx) & 1u)) { wr |= 4u; cur.link |= 1u << back_idx; if (lane == 0) rec_link(a, slot, back_idx); }
      } else {
        const uint32_t dv = (uint32_t)(tp - 2 * j) & 0xffu;
        if (dv > cur.d_back) { wr |= 5u; cur.d_back = dv; if (lane == 0) rec_update(a, slot, back_idx, (int)dv); }
      }
      dist = (int)cur.d_fwd;
    } else {
      if (lane == 0) wr = landing_updates(a, slot, real, back_idx, created, have_last, last_slot, last_fwd_idx, tp - last_tp, tp - 2 * j);
      wr = __shfl_sync(0xffffffffu, wr, 0);
      if (created) dist = 0;  // a zeroed record; back_idx != fwd_idx, so nothing written above shows here
      else {
        dist = lane == 0 ? (int)rec_dist_now(a, slot, fwd_idx) : 0;
        dist = __shfl_sync(0xffffffffu, dist, 0);
      }
      cur.ok = false;
    }
    if (dist < 1) dist = 1;
    n_processed += 1; n_skipped += (uint32_t)(dist - 1);
    line_visit(c, slot, lane);
    if (wr) {
      c.wrote = true;
      mark_written(a, key, c.grec, created, (wr & 1u) != 0, lane);
      mark_written(a, last_key, c.grec, false, (wr & 2u) != 0, lane);
    }
    if (pairs || want_ext) out_push(a, c, o, ext_fwd(key, (uint32_t)real, mask), dir, pos, pairs, want_ext, lane);
    have_last = true;
    last_junc_pos = tp; last_tp = tp; last_slot = slot; last_fwd_idx = fwd_idx; last_key = key; last_rec = cur;
    tp += dist;
  }
  if (!have_last) {  // add_fake_junction (:92-104): mid-read, facing forward
    const int pos = len / 2 - k / 2;
    const uint64_t key = FAST ? line_kmer(a, c, rel + pos) : kmer_at_t<false>(a.packed, s0 + pos, k);
    const int real = (int)code_at_t<FAST>(c.pk, s0 + pos + k - c.pk_base);
    const int mtp = 2 * pos + 1;
    const uint32_t d4 = (uint32_t)(mtp - 2 * j) & 0xffu, dr = (uint32_t)((2 * len - mtp - 2 * k + 1) - 2 * j) & 0xffu;
    int slot = FAST ? c.S->slot[2 * (rel + pos) + 1] : -1;
    bool created = false;
    uint32_t wr = 0;
    if (lane == 0) c.cnt[SS_NOJUNC]++;
    if (FAST && slot >= 0 && !line_visited(c, slot, lane)) {  // it is there already: its parked distances say what changes
      const int hs = 2 * (rel + pos) + 1;
      if (lane == 0) rec_add_cov(a, slot, real);
      if (d4 > c.S->dback[hs]) { wr = 1u; if (lane == 0) rec_update(a, slot, 4, (int)d4); }
      if (dr > c.S->hop[hs]) { wr = 1u; if (lane == 0) rec_update(a, slot, real, (int)dr); }
    } else {
      if (lane == 0) {
        slot = tbl_insert(a, key, &created);
        if (created) { a.stamps[slot] = c.stamp++; wr = 1u; }
        rec_add_cov(a, slot, real);
        if (atomicMax(rec_field(a, slot, REC_DIST + 4), d4) < d4) wr = 1u;
        if (atomicMax(rec_field(a, slot, REC_DIST + real), dr) < dr) wr = 1u;
      }
      slot = __shfl_sync(0xffffffffu, slot, 0);
      created = __shfl_sync(0xffffffffu, (int)created, 0) != 0;
      c.stamp = __shfl_sync(0xffffffffu, c.stamp, 0);
      wr = __shfl_sync(0xffffffffu, wr, 0);
    }
    if (wr) { c.wrote = true; mark_written(a, key, c.grec, created, true, lane); }
    if (FAST && created) line_publish(a, c, key, slot, lane);
    line_visit(c, slot, lane);
    if (pairs || want_ext) out_push(a, c, o, ext_fwd(key, (uint32_t)real, mask), -1, pos, pairs, want_ext, lane);
  } else {  // :205
    const uint32_t dv = (uint32_t)((2 * len - last_tp - 2 * k + 1) - 2 * j) & 0xffu;
    uint32_t wr = 0;
    if (last_rec.ok) {
      if (dv > last_rec.d_fwd) { wr = 1u; if (lane == 0) rec_update(a, last_slot, last_fwd_idx, (int)dv); }
    } else {
      if (lane == 0) wr = atomicMax(rec_field(a, last_slot, REC_DIST + last_fwd_idx), dv) < dv;
      wr = __shfl_sync(0xffffffffu, wr, 0);
    }
    if (wr) { c.wrote = true; mark_written(a, last_key, c.grec, false, true, lane); }
  }
  if (lane == 0) { c.cnt[SS_JCHECK] += n_jcheck; c.cnt[SS_PROCESSED] += n_processed; c.cnt[SS_SKIPPED] += n_skipped; }
  __syncwarp();
  out_finish(a, o, pairs, lane);
}

// The same walk READ-ONLY against the table in a.keys / a.recs: returns false as soon as the sub-read would create a
// junction, raise a stored distance or set a new link (then the record belongs to the ordered kernel).  A quiet
// sub-read leaves its landings in c.land_* (coverage counts, added when the record commits), its counters in c.cnt
// and, when c.emit, its pair-filter / extension-list side effects.
// FAST: slots, distances and link masks of every half-step of the line were parked in c.S by prefetch_line.
// LAZY (with FAST planes): junction keys and records are looked up as the walk reaches them.
template <bool FAST, bool LAZY = false>
__device__ bool dry_forward(const StitchArgs& a, WarpCtx& c, uint32_t s0, int len, int lane) {
  const int k = a.k, j = a.j;
  const uint64_t mask = kmer_mask(k);
  const int rel = (int)(s0 - c.ls);
  const int tested_end = 2 * len - 2 * k + 1 - 2 * j;
  int tp = 2 * j + 1, last_junc_pos = 0;
  bool have_last = false;
  int last_tp = 0, last_fwd_idx = 0;
  uint32_t last_dist_fwd = 0, last_link = 0;
  OutList o;
  const bool pairs = c.emit && !a.no_cleaning && a.spf != nullptr;
  const bool want_ext = c.emit && a.ext != nullptr;

  auto land = [&](int slot, int real) -> bool {
    if (c.n_land >= LAND_CAP) return false;
    if (lane == 0) { c.land_slot[c.n_land] = (uint32_t)slot; c.land_nt[c.n_land] = (uint8_t)real; }
    c.n_land++;
    return true;
  };
  // dist[0..4] and the link mask of a junction, one word per lane 0..5 (one 32-byte sector of the record)
  auto head = [&](int slot) -> uint32_t { return lane < 6 ? __ldcg(a.recs + (size_t)slot * REC_WORDS + lane) : 0u; };

  while (true) {
    bool found = false;
    int slot = -1;
    // (a single probe of the half-step the cursor lands on after a skip -- normally the next known junction -- before the
    // 32-wide look-up was measured slower with the key array in L2: 18.7 instead of 16.7 ms per 3.07 M records)
    while (tp < tested_end) {
      const int t = tp + lane;
      const bool active = t < tested_end;
      bool known = false, spc = false, tst = false;
      uint32_t cnt = 0;
      int sl = -1;
      if (active) {
        const int pos = t >> 1, dir = t & 1;
        uint32_t f;
        if (FAST && !LAZY) {
          sl = c.S->slot[2 * (rel + pos) + dir];
          f = c.S->flag[rel + pos];
        } else if (FAST) {
          const uint64_t fwd = line_kmer(a, c, rel + pos);
          sl = tbl_find(a, dir ? fwd : revcomp(fwd, k));
          f = c.S->flag[rel + pos];
        } else {
          const uint64_t fwd = kmer_at_t<false>(a.packed, s0 + pos, k);
          sl = tbl_find(a, dir ? fwd : revcomp(fwd, k));
          f = flag_byte(a, s0 + pos);
        }
        known = sl >= 0;
        spc = t - last_junc_pos >= 2 * a.spacer - 1;
        cnt = dir ? (f >> 3) & 3u : (f >> 5) & 3u;
        tst = dir ? (f & 2u) != 0 : (f & 4u) != 0;
      }
      const uint32_t am = __ballot_sync(0xffffffffu, active);
      const uint32_t hb = __ballot_sync(0xffffffffu, active && (known || spc || tst));
      const int hit = hb ? __ffs(hb) - 1 : 32;
      const uint32_t upto = hit < 32 ? (hit == 31 ? 0xffffffffu : ((2u << hit) - 1u)) : am;
      uint32_t jc = (active && ((upto >> lane) & 1u) && !known && !spc) ? cnt : 0u;
      jc = __reduce_add_sync(0xffffffffu, jc);
      if (lane == 0) c.cnt[SS_JCHECK] += jc;
      if (hb) {
        if (lane == 0) c.cnt[SS_PROCESSED] += hit;
        tp += hit;
        slot = __shfl_sync(0xffffffffu, sl, hit);
        found = true;
        break;
      }
      if (lane == 0) c.cnt[SS_PROCESSED] += __popc(am);
      tp += 32;
    }
    if (!found) break;
    if (slot < 0) return false;  // a junction would be created here
    const int pos = tp >> 1, dir = tp & 1;
    const int real = dir ? (int)code_at_t<FAST>(c.pk, s0 + pos + k - c.pk_base)
                         : (int)nt_comp(code_at_t<FAST>(c.pk, s0 + pos - 1 - c.pk_base));
    const int fwd_idx = dir ? real : 4, back_idx = dir ? 4 : real;
    uint32_t d_fwd, d_back, link;
    if (FAST && !LAZY) {
      const int hs = 2 * (rel + pos) + dir;
      d_fwd = c.S->hop[hs]; d_back = c.S->dback[hs]; link = c.S->lnk[hs];
    } else {
      const uint32_t hd = head(slot);
      d_fwd = __shfl_sync(0xffffffffu, hd, fwd_idx); d_back = __shfl_sync(0xffffffffu, hd, back_idx);
      link = __shfl_sync(0xffffffffu, hd, REC_LINK);
    }
    if (have_last) {
      const uint32_t dv = (uint32_t)(tp - last_tp) & 0xffu;
      if (dv > last_dist_fwd || dv > d_back || !((last_link >> last_fwd_idx) & 1u) || !((link >> back_idx) & 1u)) return false;
    } else if (((uint32_t)(tp - 2 * j) & 0xffu) > d_back) {
      return false;
    }
    if (!land(slot, real)) return false;
    const int dist = d_fwd < 1u ? 1 : (int)d_fwd;
    if (lane == 0) { c.cnt[SS_PROCESSED] += 1; c.cnt[SS_SKIPPED] += (unsigned long long)(dist - 1); }
    if (pairs || want_ext) {
      const uint64_t fwd = FAST ? line_kmer(a, c, rel + pos) : kmer_at_t<false>(a.packed, s0 + pos, k);
      out_push(a, c, o, ext_fwd(dir ? fwd : revcomp(fwd, k), (uint32_t)real, mask), dir, pos, pairs, want_ext, lane);
    }
    have_last = true;
    last_junc_pos = tp; last_tp = tp; last_fwd_idx = fwd_idx; last_dist_fwd = d_fwd; last_link = link;
    tp += dist;
  }
  if (!have_last) {  // the fake mid-read junction must already be there, with distances at least as long
    const int pos = len / 2 - k / 2;
    const uint64_t key = FAST ? line_kmer(a, c, rel + pos) : kmer_at_t<false>(a.packed, s0 + pos, k);
    const int real = (int)code_at_t<FAST>(c.pk, s0 + pos + k - c.pk_base);
    int slot;
    uint32_t have4, have_real;  // its dist[4] and dist[real]
    if (FAST && !LAZY) {
      const int hs = 2 * (rel + pos) + 1;
      slot = c.S->slot[hs]; have4 = c.S->dback[hs]; have_real = c.S->hop[hs];
    } else {
      slot = tbl_find(a, key);
      const uint32_t hd = slot >= 0 ? head(slot) : 0u;
      have4 = __shfl_sync(0xffffffffu, hd, 4); have_real = __shfl_sync(0xffffffffu, hd, real);
    }
    if (slot < 0) return false;
    const int mtp = 2 * pos + 1;
    const uint32_t d4 = (uint32_t)(mtp - 2 * j) & 0xffu, dr = (uint32_t)((2 * len - mtp - 2 * k + 1) - 2 * j) & 0xffu;
    if (d4 > have4 || dr > have_real) return false;
    if (!land(slot, real)) return false;
    if (lane == 0) c.cnt[SS_NOJUNC]++;
    if (pairs || want_ext) out_push(a, c, o, ext_fwd(key, (uint32_t)real, mask), -1, pos, pairs, want_ext, lane);
  } else if (((uint32_t)((2 * len - last_tp - 2 * k + 1) - 2 * j) & 0xffu) > last_dist_fwd) {
    return false;
  }
  out_finish(a, o, pairs, lane);
  return true;
}

// highest position in [s, pos) whose plane bit equals `want`, or -1; uniform across the warp
template <bool SH>
__device__ long long find_prev_bit(const uint32_t* plane, uint32_t s, uint32_t pos, bool want) {
  if (pos <= s) return -1;
  const uint32_t p = pos - 1, ws = s >> 5;
  uint32_t w = p >> 5;
  uint32_t word = ldw<SH>(plane + w);
  if (!want) word = ~word;
  if ((p & 31) != 31) word &= (2u << (p & 31)) - 1u;
  while (true) {
    if (w == ws) word &= ~((1u << (s & 31)) - 1u);
    if (word) return ((long long)w << 5) + 31 - __clz(word);
    if (w == ws) return -1;
    w--;
    word = ldw<SH>(plane + w);
    if (!want) word = ~word;
  }
}

// scanInputRead (:260-282) + getValidReads (:233-257) for the sequence line [ls, le).
// DRY: the read-only walk; returns false when the line is not quiet (always true otherwise).
template <bool FAST, bool DRY, bool LAZY = false>
__device__ bool scan_line(const StitchArgs& a, WarpCtx& c, uint32_t ls, uint32_t le, int lane) {
  const int k = a.k, j = a.j;
  uint32_t pos = le;
  while (pos > ls) {  // getUnambiguousReads hands the segments over LAST first (utils/Kmer.cpp:64-80)
    // (positions are taken relative to inv_base while searching the validity plane)
    long long hi = find_prev_bit<FAST>(c.inv, ls - c.inv_base, pos - c.inv_base, false);
    if (hi < 0) break;
    const uint32_t ee = (uint32_t)hi + 1 + c.inv_base;
    long long lo = find_prev_bit<FAST>(c.inv, ls - c.inv_base, ee - c.inv_base, true);
    const uint32_t ss = lo < 0 ? ls : (uint32_t)lo + 1 + c.inv_base;
    pos = ss;
    const int L = (int)(ee - ss);
    if (L < k || L < k + 2 * j + 1) continue;
    if (lane == 0) c.cnt[SS_UNAMBIG]++;
    // getValidReads: maximal runs of >= k Bloom-positive k-mers; npos is a virtual negative position
    const int npos = L - k + 1;
    int run_start = -1;
    for (int base = 0; base <= npos; base += 32) {
      const int i = base + lane;
      const bool v = i < npos && ((FAST ? c.S->flag[ss - ls + i] : flag_byte(a, ss + i)) & 1u);
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
          const int run_len = base + bit - run_start;
          if (run_len >= k) {
            if (DRY) { if (!dry_forward<FAST, LAZY>(a, c, ss + run_start, run_len + k - 1, lane)) return false; }
            else scan_forward<FAST>(a, c, ss + run_start, run_len + k - 1, lane);
            if (lane == 0) c.cnt[SS_NOERR]++;
          }
          run_start = -1;
        }
      }
    }
  }
  return true;
}

// phase 1: everything the walk will want to know about the line, all loads in flight together.
// n_runs >= 0: S->reskey[0 .. n_runs) lists the line's minimizer runs (slot | first position << 24, in position order);
// a run whose slot has no junction at all (a.jslot) cannot hold a junction key, so its positions are not looked up.
// n_runs < 0: every position is looked up.
template <int PF>  // positions per lane per pass
__device__ void prefetch_line(const StitchArgs& a, WarpScratch* S, uint32_t ls, int n_pos, int lane, int n_runs) {
  const int k = a.k;
  for (int i = lane; i < 2 * n_pos; i += 32) S->slot[i] = -1;
  int n_list = n_pos;
  if (n_runs >= 0) {
    if (lane < POS_CAP / 32) S->pmask[lane] = 0u;
    __syncwarp();
    for (int c = lane; c < n_runs; c += 32) {
      const uint32_t e = S->reskey[c], sl = e & ROW_SLOT_MASK;
      if ((__ldcg(a.jslot + (sl >> 5)) >> (sl & 31u)) & 1u) {
        const int st = (int)(e >> 24), en = c + 1 < n_runs ? (int)(S->reskey[c + 1] >> 24) : n_pos;
        for (int w = st >> 5; w <= (en - 1) >> 5 && en > st; w++) {
          const int lo = st > 32 * w ? st - 32 * w : 0, hi = en < 32 * w + 32 ? en - 32 * w : 32;  // bits [lo, hi) of word w
          atomicOr(&S->pmask[w], (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u));
        }
      }
    }
    __syncwarp();
    n_list = 0;
    for (int base = 0; base < n_pos; base += 32) {
      const int pos = base + lane;
      const bool on = pos < n_pos && ((S->pmask[base >> 5] >> lane) & 1u);
      const uint32_t b = __ballot_sync(0xffffffffu, on);
      if (on) S->plist[n_list + __popc(b & ((1u << lane) - 1u))] = (uint8_t)pos;
      n_list += __popc(b);
    }
    __syncwarp();
  }
  for (int base = 0; base < n_list; base += 32 * PF) {
    uint64_t fwd[PF], rcv[PF];
    unsigned long long kf[PF], kb[PF];
    uint64_t hf[PF], hb[PF];
    int ps[PF];
#pragma unroll
    for (int i = 0; i < PF; i++) {
      const int idx = base + 32 * i + lane;
      ps[i] = idx < n_list ? (n_runs >= 0 ? (int)S->plist[idx] : idx) : -1;
      if (ps[i] >= 0) fwd[i] = kmer_at_t<true>(S->pk, (ls & 15u) + ps[i], k);
    }
#pragma unroll
    for (int i = 0; i < PF; i++) {
      if (ps[i] >= 0) {
        rcv[i] = revcomp(fwd[i], k);
        hf[i] = mix64(fwd[i]) & (a.cap - 1); hb[i] = mix64(rcv[i]) & (a.cap - 1);
        kf[i] = __ldcg(a.keys + hf[i]); kb[i] = __ldcg(a.keys + hb[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < PF; i++) {
      const int pos = ps[i];
      if (pos >= 0) {
        int sf, sb;
        if (fwd[i] == KEY_EMPTY) sf = tbl_find(a, fwd[i]);
        else {
          while (kf[i] != fwd[i] && kf[i] != KEY_EMPTY) { hf[i] = (hf[i] + 1) & (a.cap - 1); kf[i] = __ldcg(a.keys + hf[i]); }
          sf = kf[i] == fwd[i] ? (int)hf[i] : -1;
        }
        if (rcv[i] == KEY_EMPTY) sb = tbl_find(a, rcv[i]);
        else {
          while (kb[i] != rcv[i] && kb[i] != KEY_EMPTY) { hb[i] = (hb[i] + 1) & (a.cap - 1); kb[i] = __ldcg(a.keys + hb[i]); }
          sb = kb[i] == rcv[i] ? (int)hb[i] : -1;
        }
        S->slot[2 * pos + 1] = sf;
        S->slot[2 * pos] = sb;
        // facing forward: fwdIdx = the read's next base, backIdx = 4; facing backward: fwdIdx = 4, backIdx = complement
        // of the base before the k-mer (utils/ReadKmer.cpp:95-114).  dist[0..4] and the link mask share one 32-byte sector.
        if (sf >= 0) {
          const uint32_t* r = rec_field(a, sf, 0);
          const uint4 d03 = __ldcg(reinterpret_cast<const uint4*>(r));
          const uint2 d45 = __ldcg(reinterpret_cast<const uint2*>(r + 4));
          const uint32_t real = code_at_t<true>(S->pk, (ls & 15u) + pos + k);
          S->hop[2 * pos + 1] = (uint8_t)(real == 0 ? d03.x : real == 1 ? d03.y : real == 2 ? d03.z : d03.w);
          S->dback[2 * pos + 1] = (uint8_t)d45.x;
          S->lnk[2 * pos + 1] = (uint8_t)d45.y;
        }
        if (sb >= 0) {
          const uint32_t* r = rec_field(a, sb, 0);
          const uint4 d03 = __ldcg(reinterpret_cast<const uint4*>(r));
          const uint2 d45 = __ldcg(reinterpret_cast<const uint2*>(r + 4));
          const uint32_t back = pos ? nt_comp(code_at_t<true>(S->pk, (ls & 15u) + pos - 1)) : 0u;
          S->hop[2 * pos] = (uint8_t)d45.x;
          S->dback[2 * pos] = (uint8_t)(back == 0 ? d03.x : back == 1 ? d03.y : back == 2 ? d03.z : d03.w);
          S->lnk[2 * pos] = (uint8_t)d45.y;
        }
      }
    }
  }
  __syncwarp();
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(STITCH_THREADS, MIN_BLOCKS) stitch_kernel(StitchArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char stitch_smem[];
  WarpScratch* S = reinterpret_cast<WarpScratch*>(stitch_smem) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * STITCH_THREADS + threadIdx.x) >> 5;
  StitchState* st = a.st;
  uint32_t next = __ldcg(&st->next), W = __ldcg(&st->W), round = __ldcg(&st->round);
  if (W > a.w_max) W = a.w_max;
  WarpCtx c;
  if (lane < SS_COUNT) S->st[lane] = 0;
  __syncwarp();
  c.S = S; c.stage = S->stage; c.cnt = S->st; c.pk = a.packed; c.pk_base = 0; c.inv = a.inval; c.inv_base = 0; c.n_stage = 0; c.part = 0;
  c.rec = 0; c.grec = 0; c.stamp = 0; c.ls = 0; c.n_pos = 0; c.n_vis = 0; c.wrote = false;
  c.land_slot = nullptr; c.land_nt = nullptr; c.n_land = 0; c.emit = true;
  uint32_t status = ST_DONE;

  while (true) {
    const int cur = round & 1, nxt = cur ^ 1;
    const uint32_t nd = __ldcg(&st->nd[cur]);
    const uint32_t room = a.n_recs - next;
    const uint32_t n_new = W > nd ? (W - nd < room ? W - nd : room) : 0u;
    const uint32_t n_win = nd + n_new;  // <= w_max <= warps of the grid
    if (n_win == 0) break;
    // n_entries / ext_used only move in phase 2, so this snapshot is the same in every thread
    const unsigned long long entries0 = __ldcg(&st->n_entries), ext0 = __ldcg(&st->ext_used);
    if (gw == 0 && lane == 0) __stcg(&st->nd[nxt], 0u);
    // ---- phase 1: reservations + lookups of the warp's record
    const unsigned long long t0 = gtime_ns();
    const bool have = gw < n_win;
    uint32_t rec = 0, ls = 0, len = 0;
    int n_res = 0;
    bool fast = false;
    if (have) {
      if (gw < nd) rec = __ldcg(a.deferred[cur] + gw);
      else { const uint32_t idx = next + (gw - nd); rec = a.list ? __ldg(a.list + idx) : idx; }
      ls = __ldg(a.seq_start + rec);
      const uint32_t le = __ldg(a.seq_end + rec);
      len = le > ls ? le - ls : 0u;
      const int n_pos = len >= (uint32_t)a.k ? (int)(len - a.k + 1) : 0;
      fast = n_pos > 0 && n_pos <= POS_CAP;
      if (fast) {
        // one coalesced sweep fetches everything the line needs from the planes
        if (lane < PK_WORDS) S->pk[lane] = __ldg(a.packed + (ls >> 4) + lane);
        if (lane < INV_WORDS) S->inv[lane] = __ldg(a.inval + (ls >> 5) + lane);
        for (int pos = lane; pos < n_pos; pos += 32) S->flag[pos] = a.flags[ls + pos];
        __syncwarp();
        const unsigned long long ta = gtime_ns();
        line_reservations<0, true>(a, S->pk, ls & 15u, len, rec, lane, S->reskey, &n_res);
        const unsigned long long tb = gtime_ns();
        prefetch_line<(MIN_BLOCKS >= 4 ? 1 : 2)>(a, S, ls, n_pos, lane, n_res <= RES_CAP ? n_res : -1);
        if (gw == 0 && lane == 0) { S->st[SS_T_P1A] += ta - t0; S->st[SS_T_P1B] += tb - ta; S->st[SS_T_P1C] += gtime_ns() - tb; }
      } else {
        line_reservations<0, false>(a, a.packed, ls, len, rec, lane, S->reskey, &n_res);
      }
      if (lane == 0 && 2ull * len + 2 > __ldcg(&st->max_need)) atomicMax(&st->max_need, 2ull * len + 2);
    }
    const unsigned long long t1 = gtime_ns();
    grid.sync();
    const unsigned long long t2 = gtime_ns();
    {
      const unsigned long long bound = __ldcg(&st->max_need) * n_win;
      if (entries0 + bound > a.cap / 2) { status = ST_GROW_TABLE; break; }
      if (a.ext && ext0 + 2 * bound + n_win > a.ext_cap) { status = ST_DRAIN_EXT; break; }
    }
    // ---- phase 2: execute or defer
    if (have) {
      bool mine;
      if (n_res <= RES_CAP) {
        bool ok = true;
        for (int i = lane; i < n_res; i += 32)
          if (__ldcg(a.res + (S->reskey[i] & ROW_SLOT_MASK)) != rec) ok = false;
        mine = __all_sync(0xffffffffu, ok);
        // (release after the whole check: two runs of a line may hash to the same slot)
        for (int i = lane; i < n_res; i += 32)
          if (__ldcg(a.res + (S->reskey[i] & ROW_SLOT_MASK)) == rec) __stcg(a.res + (S->reskey[i] & ROW_SLOT_MASK), RES_FREE);
      } else {
        mine = line_reservations<1, false>(a, a.packed, ls, len, rec, lane, nullptr, nullptr);
        line_reservations<2, false>(a, a.packed, ls, len, rec, lane, nullptr, nullptr);
      }
      if (gw == 0 && lane == 0) S->st[SS_T_P2A] += gtime_ns() - t2;
      if (mine) {
        c.rec = rec; c.part = 0; c.n_stage = 0; c.n_vis = 0; c.ls = ls; c.wrote = false;
        c.grec = global_rec(a, rec);
        c.stamp = global_stamp(a, rec);
        if (fast) {
          c.n_pos = (int)(len - a.k + 1);
          c.pk = S->pk; c.pk_base = ls & ~15u; c.inv = S->inv; c.inv_base = ls & ~31u;
          scan_line<true, false>(a, c, ls, ls + len, lane);
        } else if (len) {
          c.n_pos = 0;
          c.pk = a.packed; c.pk_base = 0; c.inv = a.inval; c.inv_base = 0;
          scan_line<false, false>(a, c, ls, ls + len, lane);
        }
        if (a.ext && c.n_stage) ext_flush(a, c, lane);
        if (c.wrote && lane == 0) S->st[SS_WRITERS]++;
      } else if (lane == 0) {
        a.deferred[nxt][atomicAdd(&st->nd[nxt], 1u)] = rec;
        S->st[SS_DEFERRED]++;
      }
    }
    const unsigned long long t3 = gtime_ns();
    grid.sync();
    if (gw == 0 && lane == 0) {
      S->st[SS_T_PHASE1] += t1 - t0; S->st[SS_T_SYNC1] += t2 - t1; S->st[SS_T_PHASE2] += t3 - t2; S->st[SS_T_SYNC2] += gtime_ns() - t3;
    }
    const uint32_t nd_next = __ldcg(&st->nd[nxt]);
    if (nd_next * a.shrink_den > n_win) W = W / 2 > a.w_min ? W / 2 : a.w_min;
    else if (nd_next * a.grow_den < n_win && n_win >= W) W = W * 2 < a.w_max ? W * 2 : a.w_max;
    next += n_new;
    round++;
    if (gw == 0 && lane == 0) S->st[SS_ROUNDS]++;
  }
  __syncwarp();
  if (lane < SS_COUNT && S->st[lane]) atomicAdd(&st->stats[lane], S->st[lane]);
  if (gw == 0 && lane == 0) {
    __stcg(&st->next, next); __stcg(&st->W, W); __stcg(&st->round, round); __stcg(&st->status, status);
  }
}

// ---- epochs: the read-only walk, the verify step, the exact-set list ------------------------------
struct DryScratch {  // shared memory of one warp of the read-only kernel
  WarpScratch W;                     // the line's planes, flags and parked lookups; W.st = the warp's totals
  unsigned long long cnt[SS_WALK];   // the current record's counters
  uint32_t land_slot[LAND_CAP];
  uint8_t land_nt[LAND_CAP];
};

// DRY_CLASSIFY        every record of [r_begin, r_end) walks against the table in a.keys / a.recs (nothing of it that a
//                     walk reads changes meanwhile); in_exact[rec] := EX_MEMBER when it is not quiet; its reservation
//                     slots go to a.rows for the verify step.
// DRY_CLASSIFY_APPLY  the same, and a quiet record commits at once: coverage counts into cov_out, scan counters.
// DRY_APPLY           the records with in_exact == want_flag walk again and commit, pair-filter adds and extension
//                     lists included.
// DRY_RETRACT         the records with in_exact == want_flag walk again and take their commit back (cov_out, and
//                     cov_out2 / st2 when given); in_exact := flag_after.
// DRY_RECHECK         the records with in_exact == want_flag walk (the caller points a.keys / a.recs at the live table):
//                     quiet -> EX_SETTLED, else -> taint_mark (they join the exact set).
__global__ void __launch_bounds__(DRY_THREADS, 3) stitch_dry_kernel(StitchArgs a) {
  extern __shared__ __align__(16) unsigned char stitch_smem[];
  DryScratch* S = reinterpret_cast<DryScratch*>(stitch_smem) + (threadIdx.x >> 5);
  WarpScratch* W = &S->W;
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * DRY_THREADS + threadIdx.x) >> 5, n_warps = (gridDim.x * DRY_THREADS) >> 5;
  if (lane < SS_COUNT) W->st[lane] = 0;
  __syncwarp();
  WarpCtx c;
  c.S = W; c.stage = W->stage; c.cnt = S->cnt; c.n_vis = 0; c.stamp = 0; c.wrote = false;
  c.land_slot = S->land_slot; c.land_nt = S->land_nt;
  c.emit = a.dry_mode == DRY_APPLY;
  const int mode = a.dry_mode;
  const bool classify = mode == DRY_CLASSIFY || mode == DRY_CLASSIFY_APPLY;
  // 32 consecutive records per warp and step: the lanes read the flags, the warp then takes the wanted records in turn
  for (uint32_t base = a.r_begin + gw * 32u; base < a.r_end; base += n_warps * 32u) {
   uint32_t todo = __ballot_sync(0xffffffffu, base + lane < a.r_end && (classify || a.in_exact[base + lane] == a.want_flag));
   while (todo) {
    const uint32_t rec = base + (uint32_t)(__ffs(todo) - 1);
    todo &= todo - 1u;
    const uint32_t ls = __ldg(a.seq_start + rec), le = __ldg(a.seq_end + rec);
    const uint32_t len = le > ls ? le - ls : 0u;
    const int n_pos = len >= (uint32_t)a.k ? (int)(len - a.k + 1) : 0;
    const bool fast = n_pos > 0 && n_pos <= POS_CAP;
    uint32_t* row = classify && !a.rows_ready ? a.rows + (size_t)(rec - a.rows_base) * ROW_WORDS : nullptr;
    c.rec = rec; c.grec = 0; c.part = 0; c.n_stage = 0; c.n_land = 0; c.ls = ls;
    if (lane < SS_WALK) S->cnt[lane] = 0;
    bool quiet = true;
    if (fast) {
      if (lane < PK_WORDS) W->pk[lane] = __ldg(a.packed + (ls >> 4) + lane);
      if (lane < INV_WORDS) W->inv[lane] = __ldg(a.inval + (ls >> 5) + lane);
      for (int pos = lane; pos < n_pos; pos += 32) W->flag[pos] = a.flags[ls + pos];
      __syncwarp();
      int n_res = 0;
      if (row) {
        line_reservations<3, true>(a, W->pk, ls & 15u, len, rec, lane, W->reskey, &n_res, ROW_WORDS - 1);
        __syncwarp();
        if (lane == 0) row[0] = n_res < ROW_WORDS ? (uint32_t)n_res : 255u;
        if (lane + 1 < ROW_WORDS && lane < n_res) row[1 + lane] = W->reskey[lane];
      } else if (!a.lazy) {  // the row classify (or stitch_rows_kernel) left
        const uint32_t* rr = a.rows + (size_t)(rec - a.rows_base) * ROW_WORDS;
        n_res = (int)__ldg(rr);
        if (lane + 1 < ROW_WORDS && lane < n_res) W->reskey[lane] = __ldg(rr + 1 + lane);
        __syncwarp();
      }
      c.n_pos = n_pos; c.pk = W->pk; c.pk_base = ls & ~15u; c.inv = W->inv; c.inv_base = ls & ~31u;
      if (a.lazy) {
        quiet = scan_line<true, true, true>(a, c, ls, le, lane);
      } else {
        prefetch_line<2>(a, W, ls, n_pos, lane, n_res < ROW_WORDS ? n_res : -1);
        quiet = scan_line<true, true>(a, c, ls, le, lane);
      }
    } else {
      if (row && lane == 0) row[0] = len ? 255u : 0u;
      if (len) {
        __syncwarp();
        c.n_pos = 0; c.pk = a.packed; c.pk_base = 0; c.inv = a.inval; c.inv_base = 0;
        quiet = scan_line<false, true>(a, c, ls, le, lane);
      }
    }
    __syncwarp();
    if (mode == DRY_RECHECK) {
      if (lane == 0) a.in_exact[rec] = quiet ? (uint8_t)EX_SETTLED : a.taint_mark;
      continue;
    }
    if (!quiet) {
      if (classify) { if (lane == 0) { a.in_exact[rec] = EX_MEMBER; W->st[SS_NONQUIET]++; } }
      else if (lane == 0) W->st[SS_DRY_ERROR]++;  // cannot happen: the walk that was found quiet is repeated
      c.n_stage = 0;
      continue;
    }
    if (mode == DRY_CLASSIFY) continue;
    const uint32_t one = mode == DRY_RETRACT ? 0xffffffffu : 1u;
    for (int i = lane; i < c.n_land; i += 32) {
      atomicAdd(a.cov_out + (size_t)S->land_slot[i] * a.cov_stride + a.cov_off + S->land_nt[i], one);
      if (mode == DRY_RETRACT && a.cov_out2) atomicAdd(a.cov_out2 + (size_t)S->land_slot[i] * REC_WORDS + REC_COV + S->land_nt[i], one);
    }
    if (lane < SS_WALK) { if (mode == DRY_RETRACT) W->st[lane] -= S->cnt[lane]; else W->st[lane] += S->cnt[lane]; }
    if (mode == DRY_RETRACT && lane == 0) a.in_exact[rec] = a.flag_after;
    if (a.ext && c.n_stage) ext_flush(a, c, lane);
   }
  }
  __syncwarp();
  if (lane < SS_COUNT && W->st[lane]) {
    atomicAdd(&a.st->stats[lane], W->st[lane]);
    if (mode == DRY_RETRACT && a.st2) atomicAdd(&a.st2->stats[lane], W->st[lane]);
  }
}

// The reservation rows of the records [r_begin, r_end) -- what classify would write -- ahead of the epoch: a pure
// function of the text (a GPU that waits for the table of a sharded epoch has time for it).
__global__ void __launch_bounds__(DRY_THREADS) stitch_rows_kernel(StitchArgs a) {
  __shared__ uint32_t keep_s[DRY_WARPS][ROW_WORDS];
  uint32_t* keep = keep_s[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * DRY_THREADS + threadIdx.x) >> 5, n_warps = (gridDim.x * DRY_THREADS) >> 5;
  for (uint32_t rec = a.r_begin + gw; rec < a.r_end; rec += n_warps) {
    const uint32_t ls = __ldg(a.seq_start + rec), le = __ldg(a.seq_end + rec);
    const uint32_t len = le > ls ? le - ls : 0u;
    const int n_pos = len >= (uint32_t)a.k ? (int)(len - a.k + 1) : 0;
    uint32_t* row = a.rows + (size_t)(rec - a.rows_base) * ROW_WORDS;
    int n = 0;
    if (n_pos > 0 && n_pos <= POS_CAP) {
      line_reservations<3, false>(a, a.packed, ls, len, rec, lane, keep, &n, ROW_WORDS - 1);
      __syncwarp();
      if (lane == 0) row[0] = n < ROW_WORDS ? (uint32_t)n : 255u;
      if (lane + 1 < ROW_WORDS && lane < n) row[1 + lane] = keep[lane];
    } else if (lane == 0) {
      row[0] = len ? 255u : 0u;
    }
    __syncwarp();
  }
}

// Which records outside the exact set may have seen something else than T0?  dirty[] / dirty_max[] hold, per
// reservation slot, the first / last record of the exact set that wrote a junction under that slot in the run
// that just ended.  For a record R:
//   no write before R under any of its slots      -> T0 is what it would have seen                     (EX_QUIET)
//   writes before R, none after R                 -> the LIVE table is what it would have seen on its keys: it is
//                                                    walked again there                                (EX_RECHECK)
//   writes before and after R                     -> it joins the exact set                            (taint_mark)
__global__ void __launch_bounds__(DRY_THREADS) stitch_verify_kernel(StitchArgs a) {
  __shared__ uint32_t keep_s[DRY_WARPS][RES_CAP];
  uint32_t* keep = keep_s[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * DRY_THREADS + threadIdx.x) >> 5, n_warps = (gridDim.x * DRY_THREADS) >> 5;
  unsigned long long added = 0;
  for (uint32_t rec = a.r_begin + gw; rec < a.r_end; rec += n_warps) {
    const uint8_t fl = a.in_exact[rec];
    if (fl != EX_QUIET && fl != EX_SETTLED) continue;
    const uint32_t* row = a.rows + (size_t)(rec - a.rows_base) * ROW_WORDS;
    const uint32_t n_row = __ldg(row);
    const uint32_t grec = (uint32_t)a.rec_base + rec;  // the marks hold global record indices
    bool before = false, after = false;
    if (n_row < ROW_WORDS) {
      if (lane < (int)n_row) {
        const uint32_t sl = __ldg(row + 1 + lane) & ROW_SLOT_MASK;
        before = __ldcg(a.dirty + sl) < grec;
        after = __ldcg(a.dirty_max + sl) > grec + 1u;
      }
    } else {  // more slots than a row holds: list them again
      const uint32_t ls = __ldg(a.seq_start + rec), le = __ldg(a.seq_end + rec);
      int n = 0;
      line_reservations<3, false>(a, a.packed, ls, le - ls, rec, lane, keep, &n);
      __syncwarp();
      if (n > RES_CAP) before = after = true;  // (and more than fit here: the ordered kernel takes the line)
      for (int i = lane; i < n && i < RES_CAP; i += 32) {
        before |= __ldcg(a.dirty + (keep[i] & ROW_SLOT_MASK)) < grec;
        after |= __ldcg(a.dirty_max + (keep[i] & ROW_SLOT_MASK)) > grec + 1u;
      }
      __syncwarp();
    }
    before = __any_sync(0xffffffffu, before);
    after = __any_sync(0xffffffffu, after);
    uint8_t nf = EX_QUIET;
    if (before) nf = (after || !a.recheck) ? a.taint_mark : (uint8_t)EX_RECHECK;
    if (lane == 0 && nf != fl) a.in_exact[rec] = nf;
    if (lane == 0 && before && nf == a.taint_mark) added++;
  }
  if (lane == 0 && added) atomicAdd(&a.st->stats[SS_TAINTED], added);
}

// in_exact flags of [r_begin, r_end) -> 0/1 counts (then exclusive prefix sum, then the ascending list)
__global__ void exact_flags_kernel(const uint8_t* __restrict__ in_exact, uint32_t r_begin, uint32_t n, uint32_t* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint8_t f = in_exact[r_begin + i];
    out[i] = (f >= EX_MEMBER && f <= EX_RETRACTED) ? 1u : 0u;
  }
}
// fallback of an epoch that applied at classify time: every record that committed is queued for retraction
__global__ void exact_mark_applied_kernel(uint8_t* __restrict__ in_exact, uint32_t r_begin, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (in_exact[r_begin + i] == EX_QUIET || in_exact[r_begin + i] >= EX_SETTLED) in_exact[r_begin + i] = EX_COMMITTED;
}
__global__ void exact_list_kernel(const uint8_t* __restrict__ in_exact, uint32_t r_begin, uint32_t n,
                                  const uint32_t* __restrict__ prefix, uint32_t* __restrict__ list, uint32_t* __restrict__ count_out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint8_t f = in_exact[r_begin + i];
    const bool in = f >= EX_MEMBER && f <= EX_RETRACTED;
    if (in) list[prefix[i]] = r_begin + i;
    if (i == n - 1) *count_out = prefix[i] + (in ? 1u : 0u);
  }
}

// ---- table maintenance ---------------------------------------------------------------------------
__global__ void stitch_rehash_kernel(const unsigned long long* __restrict__ okeys, const uint32_t* __restrict__ orecs,
                                     const unsigned long long* __restrict__ ostamps, unsigned long long ocap,
                                     unsigned long long* keys, uint32_t* recs, unsigned long long* stamps,
                                     unsigned long long cap) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= ocap;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long key = okeys[i];
    unsigned long long h;
    if (i == ocap) {  // the home of the KEY_EMPTY k-mer moves to the new last slot
      h = cap;
      keys[cap] = key;
    } else {
      if (key == KEY_EMPTY) continue;
      h = mix64(key) & (cap - 1);
      while (atomicCAS(keys + h, KEY_EMPTY, key) != KEY_EMPTY) h = (h + 1) & (cap - 1);
    }
    for (int f = 0; f < REC_WORDS; f++) recs[h * REC_WORDS + f] = orecs[i * REC_WORDS + f];
    stamps[h] = ostamps[i];
  }
}

struct JunctionOut {  // == faucet_junction_rec (include/faucet_gpu.h)
  unsigned long long kmer;
  uint8_t dist[5], cov[4], linked[5], pad[2];
  unsigned long long rank;
};

// Creation order without a sort: stamp = (record index << STAMP_SHIFT | n-th creation of that record), and the
// n-th creations of a record are dense, so  rank = (#junctions created by earlier records) + n.
//   count:   hist[record]++ for every junction          scan: exclusive prefix sum of hist
//   emit:    out[prefix[record] + n] = junction
__global__ void stitch_count_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ stamps,
                                    unsigned long long cap, unsigned int special, uint32_t* __restrict__ hist) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= cap;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const bool occ = i == cap ? special != 0 : keys[i] != KEY_EMPTY;
    if (occ) atomicAdd(hist + (stamps[i] >> STAMP_SHIFT), 1u);
  }
}

constexpr int SCAN_CHUNK = 4096;  // elements per CTA of the prefix sum (256 threads x 16)
__global__ void __launch_bounds__(256) scan_reduce_kernel(const uint32_t* __restrict__ in, unsigned long long n,
                                                          uint32_t* __restrict__ block_sums) {
  const unsigned long long base = (unsigned long long)blockIdx.x * SCAN_CHUNK;
  uint32_t v = 0;
  for (int i = threadIdx.x; i < SCAN_CHUNK; i += 256)
    if (base + i < n) v += in[base + i];
  __shared__ uint32_t ws[8];
  v = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < 8; i++) t += ws[i];
    block_sums[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t* __restrict__ sums, unsigned long long n) {
  __shared__ uint32_t wtot[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (unsigned long long base = 0; base < n; base += 1024) {
    const unsigned long long i = base + threadIdx.x;
    const uint32_t v = i < n ? sums[i] : 0u;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      const uint32_t t = wtot[threadIdx.x];
      uint32_t sx = t;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, sx, o);
        if (threadIdx.x >= o) sx += y;
      }
      wtot[threadIdx.x] = sx - t;
    }
    __syncthreads();
    const uint32_t excl = carry_s + wtot[threadIdx.x >> 5] + x - v;
    if (i < n) sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
}
// in-place exclusive scan of one chunk + the chunk's offset
__global__ void __launch_bounds__(256) scan_apply_kernel(uint32_t* __restrict__ data, unsigned long long n,
                                                         const uint32_t* __restrict__ block_sums) {
  const unsigned long long base = (unsigned long long)blockIdx.x * SCAN_CHUNK + (unsigned long long)threadIdx.x * 16;
  uint32_t v[16], t = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) { v[i] = base + i < n ? data[base + i] : 0u; t += v[i]; }
  uint32_t x = t;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  __shared__ uint32_t ws[8];
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
  __syncthreads();
  uint32_t off = block_sums[blockIdx.x] + x - t;
  for (int i = 0; i < (int)(threadIdx.x >> 5); i++) off += ws[i];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    if (base + i < n) data[base + i] = off;
    off += v[i];
  }
}

__global__ void stitch_emit_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ recs,
                                   const unsigned long long* __restrict__ stamps, unsigned long long cap,
                                   unsigned int special, const uint32_t* __restrict__ prefix, JunctionOut* __restrict__ out) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= cap;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long key = keys[i];
    const bool occ = i == cap ? special != 0 : key != KEY_EMPTY;
    if (!occ) continue;
    const unsigned long long stamp = stamps[i];
    const unsigned long long rank = (unsigned long long)prefix[stamp >> STAMP_SHIFT] + (stamp & STAMP_LOW);
    const uint32_t* r = recs + i * REC_WORDS;
    JunctionOut jo;
    jo.kmer = key;
    for (int f = 0; f < 5; f++) { jo.dist[f] = (uint8_t)r[REC_DIST + f]; jo.linked[f] = (r[REC_LINK] >> f) & 1u; }
    for (int f = 0; f < 4; f++) jo.cov[f] = (uint8_t)(r[REC_COV + f] > 255u ? 255u : r[REC_COV + f]);
    jo.pad[0] = jo.pad[1] = 0;
    jo.rank = rank;
    out[rank] = jo;
  }
}

}  // namespace faucet
