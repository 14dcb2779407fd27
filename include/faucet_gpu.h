/* faucet_gpu.h -- C ABI of libfaucet_gpu.so: the B200 (sm_100a) implementation of Faucet's two-pass
 * streaming k-mer hot path (pass 1 "Bloom load", pass 2 "junction scan").
 *
 * The reference (Shamir-Lab/Faucet) has no FFI; the boundary is the handful of C++ calls that
 * src/Faucet.cpp's main() makes.  Each entry point below names the reference call it replaces.
 * Plain pointers and sizes only; all host pointers are caller-owned; every function returns 0 on
 * success or a negative FAUCET_E_* code, with faucet_gpu_last_error() giving the message.
 * The library is callable from one host thread at a time and owns its CUDA streams internally.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * FAUCET_E_NO_DEVICE.
 */
#ifndef FAUCET_GPU_H
#define FAUCET_GPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FAUCET_E_NO_DEVICE -1
#define FAUCET_E_CUDA -2
#define FAUCET_E_ARG -3
#define FAUCET_E_IO -4
#define FAUCET_E_NOMEM -5
#define FAUCET_E_STATE -6

/* utils/Junction.h:10-53 (14-byte POD) + its key + the creation rank the host needs to re-insert
 * records into std::unordered_map in the reference's insertion order. */
typedef struct {
  uint64_t kmer;          /* oriented k-mer, ReadKmer::getKmer (utils/ReadKmer.cpp:50-57) */
  uint8_t dist[5];        /* utils/Junction.h:18 */
  uint8_t cov[4];         /* utils/Junction.h:12 */
  uint8_t linked[5];      /* utils/Junction.h:19 */
  uint8_t pad[2];
  uint64_t creation_rank; /* 0,1,2,... in the order the reference would have created them */
} faucet_junction_rec;    /* 32 bytes */

/* counters printed by ReadScanner::printScanSummary / scanReads (src/ReadScanner.cpp:19-27,352-358) */
typedef struct {
  uint64_t n_junctions, nb_jcheck_kmer, nb_no_juncs, nb_processed, nb_skipped, reads_no_errors,
      reads_processed, unambiguous_reads;
} faucet_scan_stats;

/* what load_two_filters prints (utils/Bloom.cpp:345-349) */
typedef struct {
  uint64_t reads_processed, unambiguous_reads, kmers;
  uint64_t fresh_kmers;    /* occurrences NOT found in bloo1 (they were added to bloo1, not bloo2) */
  double weight1, weight2; /* Bloom::weight() of bloo1 / bloo2 after the load */
} faucet_load_stats;

/* device-side timings of the last call, milliseconds (CUDA events on the library's stream) */
typedef struct {
  float h2d_ms, parse_ms, load_ms, scan_ms, stitch_ms, d2h_ms, total_ms;
  uint64_t kernel_launches;
  uint64_t stitch_rounds, stitch_deferred; /* reservation rounds / deferred records of the last scan */
  uint64_t stitch_phase_ns[8];             /* ns in phase 1, barrier 1, phase 2, barrier 2, then phase-1 line fetch /
                                              reservations / lookups and phase-2 check (warp 0 of the grid) */
  /* epochs of the last scan: epochs run entirely through the ordered kernel / through classify-execute-
   * verify-apply; records the ordered kernel ran (re-runs included); records the read-only walk covered; ordered runs
   * of exact sets (>= classify epochs with a non-empty set); epochs that fell back to the ordered kernel (table growth);
   * records classified not quiet; records the ordered kernel found to have written */
  uint64_t epochs_exact, epochs_classify, exact_records, dry_records, epoch_iterations, epoch_fallbacks,
      nonquiet_records, writer_records;
} faucet_timings;

/* ---- lifecycle -------------------------------------------------------------------------- */
int faucet_gpu_init(int device);          /* cudaSetDevice + stream creation; idempotent */
void faucet_gpu_shutdown(void);
const char* faucet_gpu_last_error(void);
int faucet_gpu_device_count(void);        /* 0 when no CUDA device / driver */
const char* faucet_gpu_version(void);

/* ---- Bloom geometry, host side, bit-for-bit what the reference derives ------------------- */
/* getBloomFilterFromReads: p1 = brents_fun(my_func, fp, .5, 1e-4, 1000) then
 * create_bloom_filter_optimal(estimated_kmers, (float)p1)   (src/Faucet.cpp:197-223) */
int faucet_geometry_from_reads(uint64_t estimated_kmers, uint64_t singletons, float fp, double* p1_out,
                               int* log2_tai_out, int* n_hash_out);
/* Bloom::create_bloom_filter_optimal (utils/Bloom.cpp:229-247) */
int faucet_geometry_optimal(uint64_t estimated_items, float fp, int* log2_tai_out, int* n_hash_out);
/* Bloom::create_bloom_filter_2_hash (utils/Bloom.cpp:206-226), used only with -bloom_file */
int faucet_geometry_2_hash(uint64_t estimated_items, float fp, int* log2_tai_out, int* n_hash_out);

/* ---- pass 1: replaces load_two_filters(bloo1, bloo2, file, fastq, mercy=false) ------------
 * (utils/Bloom.h:294, utils/Bloom.cpp:267-350).  bloo2_out: tai/8 bytes, reference bit layout
 * (bit h <-> byte h>>3, mask 1<<(h&7)); bloo1_out may be NULL.  The arrays are OVERWRITTEN (the
 * reference's filters start zeroed, utils/Bloom.cpp:184-186). */
int faucet_gpu_load_two_filters(const char* reads_path, int fastq, int k, int log2_tai, int n_hash,
                                uint8_t* bloo2_out, uint8_t* bloo1_out, faucet_load_stats* stats);
/* same, over FASTA/FASTQ text already in host memory (pinned or pageable) */
int faucet_gpu_load_two_filters_mem(const char* text, size_t n, int fastq, int k, int log2_tai,
                                    int n_hash, uint8_t* bloo2_out, uint8_t* bloo1_out,
                                    faucet_load_stats* stats);

/* ---- pass 2: replaces ReadScanner::scanReads(fastq, paired_ends, no_cleaning) -------------
 * (src/ReadScanner.h:66-67, src/ReadScanner.cpp:284-359) together with the JunctionMap it fills
 * (utils/JunctionMap.h:61).  short_pf / long_pf are in/out pair filters
 * (src/Faucet.cpp:265-281), either may be NULL.  *recs_out is allocated by the library, sorted by
 * creation_rank; release it with faucet_gpu_free. */
int faucet_gpu_scan(const char* reads_path, int fastq, int paired_ends, int no_cleaning, int k, int j,
                    int max_spacer_dist, const uint8_t* bloo2, int log2_tai, int n_hash,
                    uint8_t* short_pf, int spf_log2_tai, int spf_n_hash, uint8_t* long_pf,
                    int lpf_log2_tai, int lpf_n_hash, faucet_junction_rec** recs_out,
                    uint64_t* n_recs_out, faucet_scan_stats* stats);
int faucet_gpu_scan_mem(const char* text, size_t n, int fastq, int paired_ends, int no_cleaning, int k,
                        int j, int max_spacer_dist, const uint8_t* bloo2, int log2_tai, int n_hash,
                        uint8_t* short_pf, int spf_log2_tai, int spf_n_hash, uint8_t* long_pf,
                        int lpf_log2_tai, int lpf_n_hash, faucet_junction_rec** recs_out,
                        uint64_t* n_recs_out, faucet_scan_stats* stats);
/* Pass 2 WITHOUT the text: after faucet_gpu_load_two_filters[_mem] ran with the "retain_planes" tuning set,
 * the parsed planes of every batch of that stream (validity bits, 2-bit codes, record table: 3/8 byte per
 * text byte) are still in HBM, so the scan of the SAME stream -- what src/Faucet.cpp:220,241-245 does when
 * -read_scan_file is not given -- needs no second upload and no second parse.  bloo2 may be NULL (the device
 * copy pass 1 left behind is used).  FAUCET_E_STATE when nothing (or too much: "retain_budget" bytes) was
 * retained; the caller then falls back to faucet_gpu_scan[_mem]. */
int faucet_gpu_scan_retained(int paired_ends, int no_cleaning, int k, int j, int max_spacer_dist,
                             const uint8_t* bloo2, int log2_tai, int n_hash, uint8_t* short_pf,
                             int spf_log2_tai, int spf_n_hash, uint8_t* long_pf, int lpf_log2_tai,
                             int lpf_n_hash, faucet_junction_rec** recs_out, uint64_t* n_recs_out,
                             faucet_scan_stats* stats);
void faucet_gpu_free(void* p);

/* ---- downstream of the scan: the Bloom walks of the contig build ----------------
 * Batched form of JunctionMap::getValidJExtension (utils/JunctionMap.cpp:474-490), which findNeighbor (:231-462) asks
 * at every step between two junctions: for each oriented k-mer kmers[i], masks_out[i] bit nt (0..3, A C T G) = the
 * forward extension by nt is in bloo2, bit 4+nt = it also passes the depth-j check.  getValidJExtension's answer is
 * then: no bit of the high nibble -> -1, exactly one -> its index, more -> -2.  bloo2 = NULL uses the device copy
 * the last load / scan call left behind. */
int faucet_gpu_query_ext_masks(const uint64_t* kmers, uint64_t n, int k, int j, const uint8_t* bloo2, int log2_tai,
                               int n_hash, uint8_t* masks_out);

/* ---- tuning / introspection -------------------------------------------------------------- */
/* bytes of read text per pipelined device batch (default 256 MiB, ramping up from 32 MiB; tests use tiny values to exercise the
 * multi-batch path) and the timestamp epoch length (default 2^32-2) */
int faucet_gpu_set_batch_bytes(size_t bytes);
int faucet_gpu_set_epoch_limit(uint64_t stamps);
int faucet_gpu_get_timings(faucet_timings* out);
/* Knobs (none changes a result; tests use them to force every code path).  Setting one drops the cached session.
 *  stitch:  "stitch_exec" 1 = the ordered part runs as a dataflow (per-slot predecessor lists, no grid barrier; default),
 *           0 = in rounds of windowed reservations; "flow_chunk" records per dependency sort;
 *           "epoch_mode" 0 = every record through the ordered executor (default), 1 = adaptive epochs (ordered executor
 *           while most records write, then read-only classify / ordered exact set / verify / apply), 2 = classify epochs
 *           from the first record on;
 *           "epoch0" / "epoch_max" first / largest epoch in records; "epoch_switch_pct", "epoch_shrink_pct",
 *           "epoch_grow_pct" the thresholds of the epoch controller; "epoch_recheck" 1 = a record with earlier but no later
 *           writes under its slots is walked again on the live table (default), 0 = it joins the exact set; "table_cap0" initial junction-table slots (power of
 *           two, grows by rehash at load 1/2); "res_log2" log2 of the reservation-table entries; "stitch_w0" /
 *           "stitch_w_max" initial / maximal records per round; "stitch_shrink_den" / "stitch_grow_den" window
 *           adaptation; "stitch_blocks" resident CTAs per SM (2..4); "ext_cap0" u64 words of the extension-list buffer
 *           that feeds the long pair filter; "shard_force_abort" 1 = a sharded epoch's first exact run reports that the
 *           table must grow (tests of the fallback to the serial path); "dry_lazy" 1 = the read-only walks of the epochs look
 *           junction keys up as they reach them, 0 = they park the lookups of the whole line first, 2 = lazy while the key
 *           array fits L2 (default)
 *  scan:    "scan_memo" 1 = scan_flags caches the extension masks of every k-mer it has computed (default), 0 = every
 *           position from the Bloom filter; "memo_shift" cache entries = Bloom bits >> memo_shift (8 bytes each)
 *  load:    "load_sub_bytes0" / "load_sub_bytes" first / largest sub-batch of pass 1; "load_memo_log2" pass 1 caches
 *           saturated k-mers when the filter has at least 2^this bits (default 29: filters that do not fit L2)
 *  passes:  "retain_planes" 1 = pass 1 keeps the parsed planes in HBM for faucet_gpu_scan_retained, within
 *           "retain_budget" bytes (default 64 GiB) */
int faucet_gpu_set_tuning(const char* name, uint64_t value);

/* ---- device-resident stage API (bench.py "value": inputs already in HBM) ------------------
 * A session owns the device buffers for one (k, geometry, j, spacer) configuration. */
typedef struct faucet_session faucet_session;
int faucet_session_create(faucet_session** out, int k, int log2_tai, int n_hash, int j,
                          int max_spacer_dist, size_t max_text_bytes);
void faucet_session_destroy(faucet_session* s);
/* copy text into the session's device batch buffer (host or device source pointer) */
int faucet_session_set_text(faucet_session* s, const void* text, size_t n, int src_is_device);
int faucet_session_reset_filters(faucet_session* s);           /* zero bloo1/bloo2/stamps */
int faucet_session_parse(faucet_session* s, int fastq);        /* text -> 2-bit + validity planes */
int faucet_session_load(faucet_session* s);                    /* pass 1 over the parsed batch */
int faucet_session_scan_flags(faucet_session* s);              /* pass 2, order-free part */
int faucet_session_scan_flags_records(faucet_session* s, uint32_t r_begin, uint32_t r_end); /* ... of these records only */
int faucet_session_stitch(faucet_session* s, int paired_ends, int no_cleaning,
                          uint64_t* n_junctions_out);          /* pass 2, stream-order part */
/* multi-batch form of the stitch: begin once (pair filters may be NULL), then one call per parsed and
 * flagged batch; get_junctions gathers the map (and writes the short pair filter back) */
int faucet_session_stitch_begin(faucet_session* s, int paired_ends, int no_cleaning, uint8_t* short_pf,
                                int spf_log2_tai, int spf_n_hash, uint8_t* long_pf, int lpf_log2_tai,
                                int lpf_n_hash);
int faucet_session_stitch_batch(faucet_session* s);
/* The dependency sort of the dataflow executor (per-record predecessor lists) is a pure function of the parsed text:
 * it may be computed ahead of the stitch of the batch -- pass 1 does so when it retains the planes, and in a
 * multi-GPU job the GPU that owns a shard does it before rank 0 imports the shard.  Optional: stitch_batch sorts
 * itself when the batch comes without. */
int faucet_session_flow_prepare(faucet_session* s);
/* ... of the first n records only; concurrent != 0: on a second stream, next to what is queued on the session's */
int faucet_session_flow_prepare_records(faucet_session* s, uint32_t n, int concurrent);
int faucet_session_load_stats(faucet_session* s, faucet_load_stats* out, uint64_t total_lines);
int faucet_session_set_profiling(faucet_session* s, int on);   /* per-kernel CUDA-event timing */
int faucet_session_get_bloom(faucet_session* s, uint8_t* bloo2_out, uint8_t* bloo1_out);
int faucet_session_set_bloom(faucet_session* s, const uint8_t* bloo2);
int faucet_session_read_bloom(faucet_session* s, uint8_t* bloo2_out); /* device bloo2 as it stands (e.g. after the OR all-reduce) */
int faucet_session_get_junctions(faucet_session* s, faucet_junction_rec** recs_out, uint64_t* n_out,
                                 faucet_scan_stats* stats);
int faucet_session_sync(faucet_session* s);
void* faucet_session_stream(faucet_session* s);                /* cudaStream_t the kernels run on */
/* CUDA-event timing of the stages on the session stream */
int faucet_session_timer_start(faucet_session* s);
int faucet_session_timer_stop_ms(faucet_session* s, float* ms_out);
uint64_t faucet_session_kernel_launches(faucet_session* s);
/* event-timed duration (ms) accumulated per named kernel since the last reset:
 * 0=parse 1=load_A 2=load_B 3=scan_flags 4=stitch (ordered kernel) 5=stitch_dry (read-only walks) 6=stitch_verify (verify + exact-set list)
 * 7=stitch_flow_prep (dependency sort of the dataflow executor) 8=shard_copy (a sharded epoch's copy of the owner's table)
 * 9=shard_merge (the owner's merge of the per-GPU coverage counts) */
int faucet_session_kernel_ms(faucet_session* s, int which, float* ms_out, uint64_t* launches_out);

/* ---- multi-GPU: one process per GPU of one NVSwitch box, peer HBM mapped through CUDA IPC -----
 * (DESIGN.md section 6).  Exact sharded pass 1: shard g = g-th contiguous range of the stream.
 *   prepare_multi; parse; bloo1_local            (every rank, over its shard)
 *   export/open_peers(FAUCET_BUF_BLOO1_LOCAL, FAUCET_BUF_BLOOM, planes...)   handles exchanged by the caller
 *   [barrier] prefix_or; load; get_bloom(NULL,NULL) [barrier] or_allreduce [barrier]
 * Pass 2: scan_flags on every rank; rank 0 stitches its own shard, then import_planes(r) + stitch_batch
 * for r = 1..N-1 (the junction map lives on rank 0). */
enum { FAUCET_BUF_INVAL = 0, FAUCET_BUF_PACKED, FAUCET_BUF_FLAGS, FAUCET_BUF_SEQ_START, FAUCET_BUF_SEQ_END,
       FAUCET_BUF_BLOO1_LOCAL, FAUCET_BUF_BLOOM,
       FAUCET_BUF_FLOW_ROWS, FAUCET_BUF_FLOW_PREDS, /* the dependency sort of the shard (faucet_session_flow_prepare);
                                                       exported again per scan -- an all-zero handle = none */
       FAUCET_BUF_TBL_KEYS, FAUCET_BUF_TBL_RECS, FAUCET_BUF_JSLOT, /* junction table of the stitch (sharded epoch) */
       FAUCET_BUF_TBL_PACK,                                       /* the owner's table packed for the other GPUs */
       FAUCET_BUF_EXACT_LIST, FAUCET_BUF_COV_DELTA,                /* a rank's exact-set list / coverage counts of the epoch */
       FAUCET_BUF_COUNT };
#define FAUCET_IPC_HANDLE_BYTES 64
int faucet_session_prepare_multi(faucet_session* s);   /* allocates every exportable buffer */
int faucet_session_export(faucet_session* s, int what, void* handle_out /* 64 bytes */);
int faucet_session_open_peers(faucet_session* s, int what, const void* handles /* n_ranks x 64 bytes */,
                              int n_ranks, int my_rank);
int faucet_session_close_peers(faucet_session* s);
int faucet_session_bloo1_local(faucet_session* s);
int faucet_session_prefix_or(faucet_session* s);
int faucet_session_or_allreduce(faucet_session* s);
int faucet_session_import_planes(faucet_session* s, int peer_rank, size_t n_text, uint32_t n_recs, int fastq);
int faucet_session_batch_info(faucet_session* s, size_t* n_text, uint32_t* n_recs);
/* ---- the stitch ACROSS GPUs: the sharded epoch (faucet_b200/csrc/shard.cuh; sequenced by faucet_b200/multi.py) ----
 * The owner (rank 0) runs the first records of its shard through the ordered executor; from there on every rank classifies
 * ITS OWN records read-only against a replica of the owner's table, the few records that are not quiet (the exact set)
 * are executed in stream order on every replica, and the owner finally merges the per-rank coverage counts.  Result =
 * the serial stitch, bit for bit (src/ReadScanner.cpp:61-231, utils/JunctionMap.cpp:533-570).
 *   every rank:  stitch_begin;  owner: stitch_records(0, r0, 0)
 *   shard_info -> all-gather -> open_peers(TBL_PACK, JSLOT) -> shard_begin -> open_peers(EXACT_LIST, TBL_KEYS, COV_DELTA)
 *   repeat { all-gather (n_exact, need_grow) [any need_grow: shard_abort, serial path]; stop when no list grew;
 *            shard_execute(n_exact of every rank, iteration);  shard_verify }
 *   shard_finish -> all-gather stats -> owner: shard_merge (others: shard_end)
 * A rank's exact-set list lives in FAUCET_BUF_EXACT_LIST, two buffers n_recs entries apart used alternately by iteration
 * parity (a rank may write its next list while a peer still reads the current one).  stitch_records(begin, end, advance):
 * the ordered executor over a record range of the resident batch (advance: the batch is complete, later batches follow). */
#define FAUCET_SHARD_INFO_BYTES 512
#define FAUCET_SHARD_STATS 32
int faucet_session_stitch_records(faucet_session* s, uint32_t begin, uint32_t end, int advance);
int faucet_session_shard_rows(faucet_session* s, uint32_t r_begin); /* optional: reservation rows ahead of shard_begin */
int faucet_session_shard_info(faucet_session* s, uint32_t r_begin, int is_owner, void* info_out /* FAUCET_SHARD_INFO_BYTES */);
int faucet_session_shard_begin(faucet_session* s, const void* infos /* n_ranks x FAUCET_SHARD_INFO_BYTES */, int n_ranks,
                               int my_rank, int owner, uint32_t* n_exact_out);
int faucet_session_shard_execute(faucet_session* s, const uint32_t* n_exact /* n_ranks */, int iter, int* need_grow_out);
int faucet_session_shard_verify(faucet_session* s, uint32_t* n_exact_out);
int faucet_session_shard_finish(faucet_session* s, uint64_t* stats_out /* FAUCET_SHARD_STATS */);
int faucet_session_shard_merge(faucet_session* s, const uint64_t* stats_all /* n_ranks x FAUCET_SHARD_STATS */);
int faucet_session_shard_end(faucet_session* s);
int faucet_session_shard_abort(faucet_session* s);
/* host helper: n_shards contiguous, record-aligned, byte-balanced ranges of a FASTA/FASTQ text;
 * offsets_out has n_shards + 1 entries */
int faucet_host_plan_shards(const char* text, size_t n, int fastq, int n_shards, uint64_t* offsets_out);

#ifdef __cplusplus
}
#endif
#endif
