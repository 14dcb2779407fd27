// libfaucet_gpu.so: sessions, batching and the C ABI declared in include/faucet_gpu.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/faucet_gpu.h"
#include "kmer.cuh"
#include "load.cuh"
#include "multi.cuh"
#include "parse.cuh"
#include "scan.cuh"
#include "pair_filter_host.hpp"
#include "stitch.cuh"
#include "flow.cuh"
#include "shard.cuh"

using namespace faucet;

namespace {

struct Global {
  bool inited = false;
  int device = 0;
  int sm_count = 148;
  std::string err;
  size_t batch_bytes = (size_t)256 << 20;  // text bytes per pipelined device batch of the whole-pass entry points
  uint64_t epoch_limit = 0xfffffffeull;
  unsigned long long table_cap0 = 1ull << 22;  // initial junction-table slots (grows by rehash)
  int res_log2 = 24;                           // reservation table entries (u32 each)
  uint32_t stitch_w_max = 1u << 15, stitch_w0 = 2048, stitch_shrink_den = 4, stitch_grow_den = 10;
  int stitch_blocks = 3;  // resident stitch CTAs per SM the kernel is compiled for (2, 3 or 4)
  // epochs of the stitch (stitch.cuh): 0 = every record through the ordered executor (the default: with the dataflow
  // executor it is the fastest on every BASELINE config), 1 = adaptive (ordered while most records write, then
  // classify / execute / verify / apply), 2 = classify from the first record on (tests)
  int epoch_mode = 0;
  uint32_t epoch0 = 8192, epoch_max = 1u << 20;  // first / largest epoch, records
  uint32_t epoch_switch_pct = 30;                // an ordered epoch with fewer writers than this switches to classify epochs
  uint32_t epoch_shrink_pct = 14, epoch_grow_pct = 6;  // exact-set share above / below which the epoch halves / doubles
  int stitch_exec = 1;                           // ordered executor: 1 = dataflow (flow.cuh), 0 = rounds with grid barriers (stitch.cuh)
  uint32_t flow_chunk = 1u << 22;                // records per dependency sort of the dataflow executor
  bool epoch_recheck = true;                     // records with earlier but no later writes under their slots are walked again on
                                                 // the live table instead of joining the exact set
  int load_memo_log2 = 29;            // pass 1 caches saturated k-mers when the filter has at least 2^this bits
  int memo_shift = 1;                 // memo entries = Bloom bits >> memo_shift (8 bytes each): load <= ~0.3 of the 8-probe cache
  bool scan_memo = true;              // scan_flags looks the extension masks of a k-mer up before it computes them (scan.cuh)
  bool retain_planes = false;         // pass 1 keeps the parsed planes of every batch in HBM for faucet_gpu_scan_retained
  size_t retain_budget = (size_t)64 << 30;
  unsigned long long ext_cap0 = 1ull << 24;
  int dry_lazy = 2;                   // read-only walks look junction keys up as they reach them instead of parking the lookups of
                                      // the whole line first: 1 = always (10.3 vs 13.5 ms per 1.8 M records at configs[1], keys in
                                      // L2), 0 = never, 2 = while the key array fits L2 (configs[2]: 5 GB table, 288 vs ~130 ms)
  bool shard_force_abort = false;     // tests: the first exact run of a sharded epoch reports "table must grow"
  size_t load_sub_bytes0 = (size_t)1 << 20, load_sub_bytes = (size_t)64 << 20;  // first / largest load sub-batch
  faucet_timings tim{};
  faucet_session* cached = nullptr;
} g;

int fail(int code, const std::string& msg) {
  g.err = msg;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(FAUCET_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
  } while (0)

constexpr size_t TAIL_MAX = (size_t)1 << 24;  // longest partial record carried between batches
constexpr size_t TEXT_PAD = 2 * PARSE_CHUNK;
enum { KT_PARSE = 0, KT_LOAD_A, KT_LOAD_B, KT_SCAN, KT_STITCH, KT_DRY, KT_VERIFY, KT_FLOW_PREP, KT_SHARD_COPY, KT_SHARD_MERGE, KT_COUNT };

}  // namespace

struct faucet_session {
  int k = 0, log2_tai = 0, n_hash = 0, j = 0, max_spacer = 0;
  size_t cap = 0;  // bytes of text one batch may hold
  cudaStream_t stream = nullptr;
  // batch text: two buffers so that the H2D copy of the next batch (copy stream) overlaps the kernels
  // of the current one (whole-pass entry points); the second one is allocated on first use
  uint8_t* d_textbufs[3] = {nullptr, nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  cudaStream_t aux_stream = nullptr;   // a sort running next to the session's stream (flow_prepare_records)
  cudaEvent_t ev_copied[3] = {nullptr, nullptr, nullptr};
  uint8_t* d_text = nullptr;  // the current batch
  size_t n = 0;               // bytes in the current batch
  bool fastq = false, parsed = false, final_batch = true;
  uint32_t *d_inval = nullptr, *d_packed = nullptr, *d_skipA = nullptr, *d_pend = nullptr, *d_chunk = nullptr;
  ParseCounters* d_pctr = nullptr;
  LoadCounters* d_lctr = nullptr;
  uint2* d_complex = nullptr;
  uint32_t complex_cap = 0;
  ParseCounters h_pctr{};
  // pass 1 state
  unsigned long long* d_fused = nullptr;
  uint32_t* d_stamps = nullptr;
  uint32_t stamp_base = 1;
  // pass 2 state
  uint32_t* d_bloom = nullptr;  // plain bloo2
  uint32_t* d_bloom1 = nullptr; // plain bloo1 (only materialised on request)
  unsigned long long* d_memo = nullptr;  // per-k-mer extension masks under the current bloo2 (scan_flags_memo_kernel)
  uint64_t memo_entries = 0;
  bool memo_dirty = true;       // bloo2 (or j) changed since the memo was last cleared
  int memo_kind = 0;            // what the table holds: 1 = pass 1's saturated k-mers, 2 = pass 2's extension masks
  uint8_t* d_flags = nullptr;
  // record table of the parsed batch (sequence line of record r = [seq_start[r], seq_end[r]))
  uint32_t *d_seq_start = nullptr, *d_seq_end = nullptr;
  size_t rec_cap = 0;
  uint32_t n_recs = 0;          // records in the parsed batch
  // pass 2 stitch: junction table + reservation state on the device (stitch.cuh)
  bool stitching = false;
  int paired = 0, no_cleaning = 1;
  StitchState* d_st = nullptr;
  unsigned long long *d_keys = nullptr, *d_jstamps = nullptr;
  uint32_t* d_recs = nullptr;   // REC_WORDS u32 per slot
  unsigned long long tbl_cap = 0;
  uint32_t* d_res = nullptr;
  uint32_t* d_jslot = nullptr;  // one bit per reservation slot: a junction lives under it (stitch.cuh)
  uint32_t *d_dirty = nullptr, *d_dirty_max = nullptr;  // epochs: first / last record that wrote a junction under each reservation slot
  uint32_t* d_rows = nullptr;   // epochs: reservation slots of every record of the epoch (classify writes, verify reads)
  size_t rows_cap = 0;
  bool dry_ready = false;
  // dataflow executor (flow.cuh)
  uint32_t *d_frows = nullptr, *d_fpreds = nullptr, *d_fdone = nullptr, *d_fcounts = nullptr, *d_fcount_sums = nullptr;
  size_t flow_cap = 0;
  unsigned long long *d_fpairs = nullptr, *d_fpairs2 = nullptr;
  uint32_t *d_fhist = nullptr, *d_fhist_sums = nullptr;
  size_t fpairs_cap = 0, fhist_cap = 0;
  unsigned int* d_fbig = nullptr;
  const void* flow_fn = nullptr;
  int flow_grid = 0;
  // a dependency sort prepared ahead for the whole parsed batch (faucet_session_flow_prepare): in the session's own
  // buffers, or -- pass 2 on retained planes -- next to the planes pass 1 kept
  bool prep_valid = false, prep_big = false;
  uint32_t prep_n = 0;
  const uint32_t *prep_rows = nullptr, *prep_preds = nullptr;
  uint8_t* d_in_exact = nullptr;  // epochs: per record of the batch, member of the exact set
  size_t in_exact_cap = 0;
  uint32_t *d_list = nullptr, *d_eprefix = nullptr, *d_eprefix_sums = nullptr, *d_count = nullptr;  // the exact set as an ascending list
  size_t list_cap = 0;
  unsigned long long* d_snap_keys = nullptr;  // epochs: the table as it stood when the epoch began (T0)
  uint32_t* d_snap_recs = nullptr;
  unsigned long long snap_cap = 0;
  StitchState* d_st_snap = nullptr;
  StitchState h_st{};           // host copy of the stitch state after the last ordered run / apply
  uint32_t* d_spf_snap = nullptr;
  uint32_t ep_size = 0;         // records of the next epoch
  bool ep_exact = true;         // the next epoch runs entirely through the ordered kernel
  struct EpochStats { uint64_t exact_epochs = 0, classify_epochs = 0, exact_runs = 0, dry_records = 0, iterations = 0, fallbacks = 0, nonquiet = 0; } ep;
  uint32_t* d_deferred[2] = {nullptr, nullptr};
  uint32_t w_max = 0, deferred_cap = 0;
  uint32_t* d_spf = nullptr;    // device copy of the short pair filter
  uint8_t* h_spf = nullptr;     // caller's array (written back by stitch_end / get_junctions)
  int spf_log2 = 0, spf_nh = 0;
  unsigned long long* d_ext = nullptr;
  unsigned long long ext_cap = 0;
  std::vector<uint64_t> h_ext;
  LongPairFilter lpf;
  uint64_t rec_base = 0;        // global index of the first record of the current batch
  faucet_junction_rec* h_recs = nullptr;  // pinned staging of the collected junction map (grow-only)
  size_t h_recs_cap = 0, n_recs_out = 0;
  faucet_scan_stats sstats{};
  int stitch_grid = 0;
  // planes of the batches of the last pass 1 (tuning "retain_planes"): pass 2 can run without the text
  struct Retained { size_t n; uint32_t n_recs; bool fastq; uint32_t *inval, *packed, *seq_start, *seq_end;
                    uint32_t *rows, *preds; bool prep, big; };  // + the dependency sort of the batch, if pass 1 prepared one
  std::vector<Retained> retained;
  struct Arena { uint8_t* p; size_t cap, used; };  // device blocks the retained planes live in; kept across passes
  std::vector<Arena> arena;
  size_t retained_bytes = 0;
  bool retained_valid = false;
  uint64_t retained_lines = 0;
  char* h_stage[3] = {nullptr, nullptr, nullptr};  // pinned staging buffers of the file reader (whole-pass entry points)
  size_t h_stage_cap = 0;
  uint32_t *d_hist = nullptr, *d_hist_sums = nullptr;  // junction-creation counts per record (ordering)
  unsigned long long hist_cap = 0;
  faucet_junction_rec* d_out = nullptr;
  size_t out_cap = 0;
  // sharded epoch (shard.cuh): this GPU's own coverage counts / scan counters of its quiet records, a peer's exact list
  uint32_t* d_cov_delta = nullptr;          // 4 u32 per table slot
  unsigned long long cov_delta_cap = 0;
  StitchState* d_st_quiet = nullptr;
  uint32_t rows_ahead_n = 0, rows_ahead_begin = 0;  // d_rows holds the rows of the records [begin, n) of the parsed batch (shard_rows)
  bool rows_ahead = false;
  // the gathered exact set: a small local batch of the lines of all its members (shard.cuh)
  uint32_t *mb_inval = nullptr, *mb_packed = nullptr, *mb_seq_start = nullptr, *mb_seq_end = nullptr, *mb_gid = nullptr, *mb_span = nullptr,
           *mb_span_sums = nullptr;
  uint8_t* mb_flags = nullptr;
  size_t mb_entries_cap = 0, mb_pos_cap = 0;
  PackedJunction* d_tbl_pack = nullptr;     // owner: T0 as the other GPUs fetch it (16 bytes per junction)
  unsigned long long tbl_pack_cap = 0;      // entries
  unsigned int* d_pack_n = nullptr;
  PackedJunction* d_tbl_in = nullptr;       // others: the local copy of the owner's pack
  unsigned long long tbl_in_cap = 0;
  const uint32_t* cur_gid = nullptr;        // global record indices of the batch the executor runs (NULL: rec_base + index)
  struct Shard {
    bool active = false, snapshot = false, ran = false;
    int owner = 0, iter = 0;
    uint32_t r_begin = 0;                   // this GPU's records [r_begin, n_recs) belong to the epoch
    uint64_t rec_base[MAX_PEERS + 1] = {};  // global index of record 0 of every shard; [n_ranks] = records of the stream
    uint32_t n_recs[MAX_PEERS] = {};
  } shard;
  // multi-GPU: peer buffers mapped through CUDA IPC (multi.cuh)
  uint32_t* d_b1local = nullptr;            // OR of every k-mer of this GPU's shard (plain layout)
  int n_ranks = 1, rank = 0;
  void* peer[FAUCET_BUF_COUNT][MAX_PEERS] = {};
  char peer_handle[FAUCET_BUF_COUNT][MAX_PEERS][FAUCET_IPC_HANDLE_BYTES] = {};
  bool peers_open = false;
  const void* stitch_fn = nullptr;
  // bookkeeping
  uint64_t launches = 0;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  bool profile = false;
  struct Ev { int which; cudaEvent_t a, b; };
  std::vector<Ev> evs;
  float kms[KT_COUNT] = {0};
  uint64_t kn[KT_COUNT] = {0};

  uint64_t tai() const { return 1ull << log2_tai; }
};

namespace {

struct KTimer {  // optional CUDA-event bracket around one kernel
  faucet_session* s; int which; cudaEvent_t a = nullptr, b = nullptr;
  KTimer(faucet_session* s_, int w) : s(s_), which(w) {
    if (s->profile) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s->stream); }
  }
  ~KTimer() {
    if (s->profile) { cudaEventRecord(b, s->stream); s->evs.push_back({which, a, b}); }
  }
};

void drain_events(faucet_session* s) {
  for (auto& e : s->evs) {
    float ms = 0;
    cudaEventSynchronize(e.b);
    cudaEventElapsedTime(&ms, e.a, e.b);
    s->kms[e.which] += ms;
    s->kn[e.which]++;
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  s->evs.clear();
}

template <class T>
int dmalloc(T** p, size_t count) {
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) return fail(FAUCET_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  return 0;
}

int ensure_load_buffers(faucet_session* s) {
  if (s->d_fused) return 0;
  int rc;
  if ((rc = dmalloc(&s->d_fused, s->tai() / 32))) return rc;
  if ((rc = dmalloc(&s->d_stamps, s->tai()))) return rc;
  if ((rc = dmalloc(&s->d_pend, s->cap / 32 + TEXT_PAD))) return rc;
  return faucet_session_reset_filters(s);
}

int ensure_scan_buffers(faucet_session* s) {
  int rc;
  if (!s->d_bloom && (rc = dmalloc(&s->d_bloom, s->tai() / 32))) return rc;
  if (!s->d_flags && (rc = dmalloc(&s->d_flags, s->cap + TEXT_PAD))) return rc;
  return 0;
}

#define DISPATCH_NH(kernel, nh, grid, block, stream, args)                     \
  switch (nh) {                                                                \
    case 1: kernel<1><<<grid, block, 0, stream>>>(args); break;                \
    case 2: kernel<2><<<grid, block, 0, stream>>>(args); break;                \
    case 3: kernel<3><<<grid, block, 0, stream>>>(args); break;                \
    case 4: kernel<4><<<grid, block, 0, stream>>>(args); break;                \
    case 5: kernel<5><<<grid, block, 0, stream>>>(args); break;                \
    case 6: kernel<6><<<grid, block, 0, stream>>>(args); break;                \
    default: kernel<0><<<grid, block, 0, stream>>>(args); break;               \
  }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(FAUCET_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

const char* faucet_gpu_last_error(void) { return g.err.c_str(); }
const char* faucet_gpu_version(void) { return "faucet_b200 0.1 (sm_100a)"; }

int faucet_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int faucet_gpu_init(int device) {
  if (g.inited && g.device == device) return 0;
  int n = faucet_gpu_device_count();
  if (n <= 0) return fail(FAUCET_E_NO_DEVICE, "no CUDA device visible: libfaucet_gpu has no CPU fallback");
  if (device < 0 || device >= n) return fail(FAUCET_E_ARG, "device index out of range");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  g.sm_count = prop.multiProcessorCount;
  g.device = device;

  g.inited = true;
  return 0;
}

void faucet_gpu_shutdown(void) {
  if (g.cached) { faucet_session_destroy(g.cached); g.cached = nullptr; }
  g.inited = false;
}

int faucet_gpu_set_batch_bytes(size_t bytes) {
  if (bytes < 1024) return fail(FAUCET_E_ARG, "batch too small");
  if (bytes > ((size_t)3 << 30)) return fail(FAUCET_E_ARG, "batch must stay below 3 GiB (32-bit offsets)");
  g.batch_bytes = bytes;
  if (g.cached) { faucet_session_destroy(g.cached); g.cached = nullptr; }
  return 0;
}
int faucet_gpu_set_epoch_limit(uint64_t stamps) {
  if (stamps < 64 || stamps > 0xfffffffeull) return fail(FAUCET_E_ARG, "epoch limit out of range");
  g.epoch_limit = stamps;
  return 0;
}
int faucet_gpu_get_timings(faucet_timings* out) { *out = g.tim; return 0; }
int faucet_gpu_set_tuning(const char* name, uint64_t value) {
  std::string n(name ? name : "");
  if (n == "table_cap0") {
    if (value < 64 || (value & (value - 1))) return fail(FAUCET_E_ARG, "table_cap0 must be a power of two >= 64");
    g.table_cap0 = value;
  } else if (n == "res_log2") {
    if (value < 8 || value > 24) return fail(FAUCET_E_ARG, "res_log2 out of range (8..24)");
    g.res_log2 = (int)value;
  } else if (n == "stitch_w_max") {
    if (value < 1 || value > (1u << 22)) return fail(FAUCET_E_ARG, "stitch_w_max out of range");
    g.stitch_w_max = (uint32_t)value;
  } else if (n == "stitch_shrink_den" || n == "stitch_grow_den") {
    if (value < 1 || value > 1000) return fail(FAUCET_E_ARG, "stitch window thresholds out of range");
    (n == "stitch_shrink_den" ? g.stitch_shrink_den : g.stitch_grow_den) = (uint32_t)value;
  } else if (n == "stitch_blocks") {
    if (value < 2 || value > 4) return fail(FAUCET_E_ARG, "stitch_blocks must be 2, 3 or 4");
    g.stitch_blocks = (int)value;
  } else if (n == "epoch_mode") {
    if (value > 2) return fail(FAUCET_E_ARG, "epoch_mode must be 0 (one ordered run), 1 (adaptive) or 2 (always classify)");
    g.epoch_mode = (int)value;
  } else if (n == "stitch_exec") {
    if (value > 1) return fail(FAUCET_E_ARG, "stitch_exec must be 0 (rounds) or 1 (dataflow)");
    g.stitch_exec = (int)value;
  } else if (n == "flow_chunk") {
    if (value < 1 || value > (1u << 24)) return fail(FAUCET_E_ARG, "flow_chunk out of range");
    g.flow_chunk = (uint32_t)value;
  } else if (n == "epoch_recheck") {
    g.epoch_recheck = value != 0;
  } else if (n == "epoch0" || n == "epoch_max") {
    if (value < 1 || value > (1u << 30)) return fail(FAUCET_E_ARG, "epoch size out of range");
    (n == "epoch0" ? g.epoch0 : g.epoch_max) = (uint32_t)value;
    if (g.epoch_max < g.epoch0) g.epoch_max = g.epoch0;
  } else if (n == "epoch_switch_pct" || n == "epoch_shrink_pct" || n == "epoch_grow_pct") {
    if (value > 100) return fail(FAUCET_E_ARG, "percentage out of range");
    (n == "epoch_switch_pct" ? g.epoch_switch_pct : n == "epoch_shrink_pct" ? g.epoch_shrink_pct : g.epoch_grow_pct) = (uint32_t)value;
  } else if (n == "scan_memo") {
    g.scan_memo = value != 0;
  } else if (n == "load_memo_log2") {
    g.load_memo_log2 = (int)value;
  } else if (n == "memo_shift") {
    if (value > 16) return fail(FAUCET_E_ARG, "memo_shift out of range");
    g.memo_shift = (int)value;
  } else if (n == "retain_planes") {
    g.retain_planes = value != 0;
  } else if (n == "retain_budget") {
    g.retain_budget = (size_t)value;
  } else if (n == "stitch_w0") {
    if (value < 1) return fail(FAUCET_E_ARG, "stitch_w0 out of range");
    g.stitch_w0 = (uint32_t)value;
  } else if (n == "load_sub_bytes0") {
    if (value < 32) return fail(FAUCET_E_ARG, "load_sub_bytes0 out of range");
    g.load_sub_bytes0 = value;
  } else if (n == "load_sub_bytes") {
    if (value < 32) return fail(FAUCET_E_ARG, "load_sub_bytes out of range");
    g.load_sub_bytes = value;
  } else if (n == "ext_cap0") {
    if (value < 64) return fail(FAUCET_E_ARG, "ext_cap0 out of range");
    g.ext_cap0 = value;
  } else if (n == "dry_lazy") {
    if (value > 2) return fail(FAUCET_E_ARG, "dry_lazy must be 0, 1 or 2 (while the keys fit L2)");
    g.dry_lazy = (int)value;
  } else if (n == "shard_force_abort") {
    g.shard_force_abort = value != 0;
  } else {
    return fail(FAUCET_E_ARG, "unknown tuning knob: " + n);
  }
  if (g.cached) { faucet_session_destroy(g.cached); g.cached = nullptr; }  // sessions size their buffers at creation
  return 0;
}
void faucet_gpu_free(void* p) { free(p); }

// ---- sessions ----------------------------------------------------------------------------------

int faucet_session_create(faucet_session** out, int k, int log2_tai, int n_hash, int j, int max_spacer_dist,
                          size_t max_text_bytes) {
  if (!g.inited) { int rc = faucet_gpu_init(0); if (rc) return rc; }
  if (k < 2 || k > 32) return fail(FAUCET_E_ARG, "k must be in [2,32]");
  if (log2_tai < 6 || log2_tai > 37) return fail(FAUCET_E_ARG, "log2_tai must be in [6,37]");  // word index in 32 bits
  if (n_hash < 1 || n_hash > MAX_NHASH) return fail(FAUCET_E_ARG, "n_hash must be in [1,10]");
  if (j < 0 || j > MAX_J) return fail(FAUCET_E_ARG, "j must be in [0,4]");
  if (max_text_bytes > ((size_t)3 << 30)) return fail(FAUCET_E_ARG, "a batch must stay below 3 GiB");
  faucet_session* s = new faucet_session();
  s->k = k; s->log2_tai = log2_tai; s->n_hash = n_hash; s->j = j; s->max_spacer = max_spacer_dist;
  s->cap = ((max_text_bytes + TAIL_MAX + PARSE_CHUNK - 1) / PARSE_CHUNK) * PARSE_CHUNK;
  int rc = 0;
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete s; return fail(FAUCET_E_CUDA, cudaGetErrorString(e)); }
  cudaEventCreate(&s->t0);
  cudaEventCreate(&s->t1);
  size_t words = s->cap / 32 + TEXT_PAD;
  s->complex_cap = (uint32_t)(s->cap / 64 + 16);
  s->rec_cap = s->cap / 4 + 16;  // a record is at least a header and a sequence line
  if ((rc = dmalloc(&s->d_seq_start, s->rec_cap)) || (rc = dmalloc(&s->d_seq_end, s->rec_cap)) ||
      (rc = dmalloc(&s->d_textbufs[0], s->cap + TEXT_PAD)) || (rc = dmalloc(&s->d_inval, words)) ||
      (rc = dmalloc(&s->d_packed, 2 * words)) || (rc = dmalloc(&s->d_skipA, words)) ||
      (rc = dmalloc(&s->d_chunk, s->cap / PARSE_CHUNK + 16)) || (rc = dmalloc(&s->d_pctr, 1)) ||
      (rc = dmalloc(&s->d_lctr, 1)) || (rc = dmalloc(&s->d_complex, s->complex_cap))) {
    faucet_session_destroy(s);
    return rc;
  }
  cudaMemsetAsync(s->d_inval, 0xff, words * 4, s->stream);
  cudaMemsetAsync(s->d_packed, 0, 2 * words * 4, s->stream);
  cudaMemsetAsync(s->d_lctr, 0, sizeof(LoadCounters), s->stream);
  cudaMemsetAsync(s->d_textbufs[0], '\n', s->cap + TEXT_PAD, s->stream);
  s->d_text = s->d_textbufs[0];
  cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 3; i++) cudaEventCreateWithFlags(&s->ev_copied[i], cudaEventDisableTiming);
  *out = s;
  return 0;
}

static void retained_free(faucet_session* s);
static int memo_acquire(faucet_session* s, int kind, int* qbits);

void faucet_session_destroy(faucet_session* s) {
  if (!s) return;
  if (s->stream) cudaStreamSynchronize(s->stream);
  drain_events(s);
  if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); }
  if (s->aux_stream) { cudaStreamSynchronize(s->aux_stream); cudaStreamDestroy(s->aux_stream); }
  for (int i = 0; i < 3; i++) { cudaFree(s->d_textbufs[i]); if (s->ev_copied[i]) cudaEventDestroy(s->ev_copied[i]); }
  cudaFree(s->d_inval); cudaFree(s->d_packed); cudaFree(s->d_skipA);
  cudaFree(s->d_pend); cudaFree(s->d_chunk); cudaFree(s->d_pctr); cudaFree(s->d_lctr);
  cudaFree(s->d_complex); cudaFree(s->d_fused); cudaFree(s->d_stamps); cudaFree(s->d_bloom); cudaFree(s->d_memo);
  retained_free(s);
  for (int i = 0; i < 3; i++) if (s->h_stage[i]) cudaFreeHost(s->h_stage[i]);
  if (s->h_recs) cudaFreeHost(s->h_recs);
  cudaFree(s->d_bloom1); cudaFree(s->d_flags); cudaFree(s->d_seq_start); cudaFree(s->d_seq_end);
  cudaFree(s->d_st); cudaFree(s->d_keys); cudaFree(s->d_jstamps); cudaFree(s->d_recs); cudaFree(s->d_res); cudaFree(s->d_jslot); cudaFree(s->d_dirty); cudaFree(s->d_dirty_max); cudaFree(s->d_rows);
  cudaFree(s->d_frows); cudaFree(s->d_fpreds); cudaFree(s->d_fdone); cudaFree(s->d_fcounts); cudaFree(s->d_fcount_sums);
  cudaFree(s->d_fpairs); cudaFree(s->d_fpairs2); cudaFree(s->d_fhist); cudaFree(s->d_fhist_sums); cudaFree(s->d_fbig); cudaFree(s->d_in_exact);
  cudaFree(s->d_list); cudaFree(s->d_eprefix); cudaFree(s->d_eprefix_sums); cudaFree(s->d_count); cudaFree(s->d_snap_keys); cudaFree(s->d_snap_recs);
  cudaFree(s->d_st_snap); cudaFree(s->d_spf_snap); cudaFree(s->d_cov_delta); cudaFree(s->d_st_quiet); cudaFree(s->d_tbl_pack); cudaFree(s->d_pack_n); cudaFree(s->d_tbl_in);
  cudaFree(s->mb_inval); cudaFree(s->mb_packed); cudaFree(s->mb_seq_start); cudaFree(s->mb_seq_end); cudaFree(s->mb_gid); cudaFree(s->mb_span); cudaFree(s->mb_span_sums); cudaFree(s->mb_flags);
  cudaFree(s->d_deferred[0]); cudaFree(s->d_deferred[1]); cudaFree(s->d_spf); cudaFree(s->d_ext);
  faucet_session_close_peers(s);
  cudaFree(s->d_b1local); cudaFree(s->d_hist); cudaFree(s->d_hist_sums); cudaFree(s->d_out);
  if (s->t0) cudaEventDestroy(s->t0);
  if (s->t1) cudaEventDestroy(s->t1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

void* faucet_session_stream(faucet_session* s) { return (void*)s->stream; }
int faucet_session_sync(faucet_session* s) {
  CU(cudaStreamSynchronize(s->stream));
  drain_events(s);
  return 0;
}
uint64_t faucet_session_kernel_launches(faucet_session* s) { return s->launches; }
int faucet_session_timer_start(faucet_session* s) { CU(cudaEventRecord(s->t0, s->stream)); return 0; }
int faucet_session_timer_stop_ms(faucet_session* s, float* ms_out) {
  CU(cudaEventRecord(s->t1, s->stream));
  CU(cudaEventSynchronize(s->t1));
  CU(cudaEventElapsedTime(ms_out, s->t0, s->t1));
  return 0;
}
int faucet_session_set_profiling(faucet_session* s, int on) {
  drain_events(s);
  s->profile = on != 0;
  for (int i = 0; i < KT_COUNT; i++) { s->kms[i] = 0; s->kn[i] = 0; }
  return 0;
}
int faucet_session_kernel_ms(faucet_session* s, int which, float* ms_out, uint64_t* launches_out) {
  if (which < 0 || which >= KT_COUNT) return fail(FAUCET_E_ARG, "bad kernel id");
  drain_events(s);
  *ms_out = s->kms[which];
  if (launches_out) *launches_out = s->kn[which];
  return 0;
}

// copies one batch of text into text buffer `buf` on stream `st` (not yet the current batch)
static int stage_text(faucet_session* s, int buf, const void* text, size_t n, cudaMemcpyKind kind, cudaStream_t st) {
  if (n > s->cap - TAIL_MAX) return fail(FAUCET_E_ARG, "text larger than the session batch capacity");
  if (!s->d_textbufs[buf]) {
    int rc = dmalloc(&s->d_textbufs[buf], s->cap + TEXT_PAD);
    if (rc) return rc;
  }
  uint8_t* dst = s->d_textbufs[buf];
  if (n) CU(cudaMemcpyAsync(dst, text, n, kind, st));
  // bytes past the end must not look like bases of a previous, longer batch
  CU(cudaMemsetAsync(dst + n, '\n', TEXT_PAD, st));
  return 0;
}
static void select_text(faucet_session* s, int buf, size_t n) {
  s->d_text = s->d_textbufs[buf];
  s->n = n;
  s->parsed = false;
}

int faucet_session_set_text(faucet_session* s, const void* text, size_t n, int src_is_device) {
  int rc = stage_text(s, 0, text, n, src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s->stream);
  if (rc) return rc;
  select_text(s, 0, n);
  return 0;
}

int faucet_session_reset_filters(faucet_session* s) {
  if (!s->d_fused) return ensure_load_buffers(s);
  CU(cudaMemsetAsync(s->d_fused, 0, s->tai() / 32 * 8, s->stream));
  CU(cudaMemsetAsync(s->d_stamps, 0xff, s->tai() * 4, s->stream));
  CU(cudaMemsetAsync(s->d_lctr, 0, sizeof(LoadCounters), s->stream));
  s->stamp_base = 1;
  s->memo_dirty = true;  // pass 1's cache of saturated k-mers describes the filters that were just emptied
  return 0;
}

static int parse_batch(faucet_session* s, bool fastq, bool final_batch) {
  s->prep_valid = false;  // a dependency sort belongs to one parsed batch
  s->rows_ahead = false;
  s->fastq = fastq;
  s->final_batch = final_batch;
  uint32_t n_chunks = (uint32_t)((s->n + PARSE_CHUNK - 1) / PARSE_CHUNK);
  if (n_chunks == 0) n_chunks = 1;
  size_t words = (size_t)n_chunks * (PARSE_CHUNK / 32);
  CU(cudaMemsetAsync(s->d_pctr, 0, sizeof(ParseCounters), s->stream));
  CU(cudaMemsetAsync(s->d_skipA, 0, (words + 2) * 4, s->stream));
  // records of this batch <= lines / period (+1 for a ragged tail)
  const size_t rec_bound = std::min(s->rec_cap, s->n / (fastq ? 4 : 2) + 2);
  CU(cudaMemsetAsync(s->d_seq_start, 0, rec_bound * 4, s->stream));
  CU(cudaMemsetAsync(s->d_seq_end, 0, rec_bound * 4, s->stream));
  {
    KTimer kt(s, KT_PARSE);
    parse_count_kernel<<<n_chunks, PARSE_THREADS, 0, s->stream>>>(s->d_text, s->n, s->d_chunk);
    parse_scan_kernel<<<1, 1024, 0, s->stream>>>(s->d_chunk, n_chunks, s->d_pctr);
    ParseArgs a;
    a.text = s->d_text; a.n = s->n; a.inval = s->d_inval; a.packed = s->d_packed; a.skipA = s->d_skipA;
    a.chunk_prefix = s->d_chunk; a.ctr = s->d_pctr; a.complex_list = s->d_complex; a.complex_cap = s->complex_cap;
    a.period_mask = fastq ? 3 : 1; a.final_batch = final_batch ? 1 : 0; a.k = s->k;
    a.seq_start = s->d_seq_start; a.seq_end = s->d_seq_end; a.rec_shift = fastq ? 2 : 1; a.rec_cap = (uint32_t)rec_bound;
    parse_planes_kernel<<<n_chunks, PARSE_THREADS, 0, s->stream>>>(a);
    s->launches += 3;
  }
  // the two guard words after the covered range stay "invalid"
  CU(cudaMemsetAsync(s->d_inval + words, 0xff, 8, s->stream));
  CU(cudaMemcpyAsync(&s->h_pctr, s->d_pctr, sizeof(ParseCounters), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  int rc = check_launch("parse");
  if (rc) return rc;
  if (s->h_pctr.complex_overflow) return fail(FAUCET_E_NOMEM, "too many multi-segment lines in one batch");
  {  // records = header lines that belong to this batch (a ragged tail counts: std::getline semantics)
    const uint64_t period = fastq ? 4 : 2;
    uint64_t lines = s->h_pctr.total_newlines;
    if (final_batch) {
      uint8_t last = '\n';
      if (s->n) CU(cudaMemcpy(&last, s->d_text + s->n - 1, 1, cudaMemcpyDeviceToHost));
      if (s->n && last != '\n') lines++;
      s->n_recs = (uint32_t)((lines + period - 1) / period);
    } else {
      s->n_recs = (uint32_t)(lines / period);
    }
    if (s->n_recs > rec_bound) return fail(FAUCET_E_NOMEM, "too many records in one batch");
  }
  s->parsed = true;
  return 0;
}

int faucet_session_parse(faucet_session* s, int fastq) { return parse_batch(s, fastq != 0, true); }

int faucet_session_load(faucet_session* s) {
  if (!s->parsed) return fail(FAUCET_E_STATE, "faucet_session_load before faucet_session_parse");
  int rc = ensure_load_buffers(s);
  if (rc) return rc;
  if ((uint64_t)s->stamp_base + s->n + 1 > g.epoch_limit) {
    stamps_epoch_kernel<<<g.sm_count * 8, 256, 0, s->stream>>>(s->d_stamps, s->tai());
    s->launches++;
    s->stamp_base = 1;
  }
  LoadArgs a;
  a.inval = s->d_inval; a.packed = s->d_packed; a.skipA = s->d_skipA; a.pend = s->d_pend;
  a.n_words = (uint32_t)((s->n + 31) / 32);
  a.fused = s->d_fused; a.stamps = s->d_stamps; a.tai_mask = s->tai() - 1; a.base = s->stamp_base;
  a.k = s->k; a.n_hash = s->n_hash; a.ctr = s->d_lctr; a.text = s->d_text; a.complex_list = s->d_complex;
  a.n_complex = s->h_pctr.n_complex;
  a.memo = nullptr; a.memo_mask = 0; a.memo_qbits = 0;
  if (g.scan_memo && s->log2_tai >= g.load_memo_log2) {  // filters that do not fit L2: cache the saturated k-mers (load.cuh)
    int qb = 0;
    if ((rc = memo_acquire(s, 1, &qb))) return rc;
    if (s->d_memo) { a.memo = s->d_memo; a.memo_mask = s->memo_entries - 1; a.memo_qbits = qb; }
  }

  // Sub-batches: kernel A treats "all bits already in bloo1 when the sub-batch starts" as contained and
  // only the rest touches the 4-byte-per-bit stamp array, so short early sub-batches (bloo1 fills
  // fast) keep almost every k-mer of a deep-coverage stream off the stamps.  Any partition by start
  // offset is exact: stamps are global and monotone (load.cuh).
  const uint32_t sub0 = (uint32_t)std::max<size_t>(1, g.load_sub_bytes0 / 32), sub_max = (uint32_t)std::max<size_t>(sub0, g.load_sub_bytes / 32);
  uint32_t sub = sub0;
  for (uint32_t wb = 0; wb < a.n_words; ) {
    a.w_begin = wb;
    a.w_end = (uint32_t)std::min<uint64_t>((uint64_t)wb + sub, a.n_words);
    const int grid = (int)std::min<uint32_t>(g.sm_count * 8, (a.w_end - a.w_begin + 7) / 8);
    {
      KTimer kt(s, KT_LOAD_A);
      const int grid_a = (int)std::min<uint32_t>(g.sm_count * LOAD_CTAS_PER_SM, (a.w_end - a.w_begin + 7) / 8);  // one full wave
      if (a.memo) {
        switch (s->n_hash) {
          case 1: load_A_kernel<1, true><<<grid_a, LOAD_THREADS, 0, s->stream>>>(a); break;
          case 2: load_A_kernel<2, true><<<grid_a, LOAD_THREADS, 0, s->stream>>>(a); break;
          case 3: load_A_kernel<3, true><<<grid_a, LOAD_THREADS, 0, s->stream>>>(a); break;
          case 4: load_A_kernel<4, true><<<grid_a, LOAD_THREADS, 0, s->stream>>>(a); break;
          default: load_A_kernel<0, true><<<grid_a, LOAD_THREADS, 0, s->stream>>>(a); break;
        }
      } else {
        DISPATCH_NH(load_A_kernel, s->n_hash, grid_a, LOAD_THREADS, s->stream, a);
      }
      s->launches++;
    }
    if (a.n_complex) {
      load_complex_kernel<0><<<std::min<uint32_t>(g.sm_count * 8, (a.n_complex + 7) / 8), LOAD_THREADS, 0, s->stream>>>(a);
      s->launches++;
    }
    {
      KTimer kt(s, KT_LOAD_B);
      DISPATCH_NH(load_B_kernel, s->n_hash, grid, LOAD_THREADS, s->stream, a);
      s->launches++;
    }
    if (a.n_complex) {
      load_complex_kernel<1><<<std::min<uint32_t>(g.sm_count * 8, (a.n_complex + 7) / 8), LOAD_THREADS, 0, s->stream>>>(a);
      s->launches++;
    }
    wb = a.w_end;
    sub = std::min<uint64_t>((uint64_t)sub * 2, sub_max);
  }
  s->stamp_base += (uint32_t)s->n + 1;
  return check_launch("load");
}

// split the fused filters into plain arrays on the device; copies to the host if pointers are given
int faucet_session_get_bloom(faucet_session* s, uint8_t* bloo2_out, uint8_t* bloo1_out) {
  if (!s->d_fused) return fail(FAUCET_E_STATE, "no load pass has run in this session");
  int rc;
  if (!s->d_bloom && (rc = dmalloc(&s->d_bloom, s->tai() / 32))) return rc;
  if (bloo1_out && !s->d_bloom1 && (rc = dmalloc(&s->d_bloom1, s->tai() / 32))) return rc;
  CU(cudaMemsetAsync(&s->d_lctr->weight1, 0, 16, s->stream));
  bloom_split_kernel<<<g.sm_count * 8, 256, 0, s->stream>>>(s->d_fused, s->tai() / 32, bloo1_out ? s->d_bloom1 : nullptr,
                                                           s->d_bloom, s->d_lctr);
  s->memo_dirty = true;
  s->launches++;
  if (bloo2_out) CU(cudaMemcpyAsync(bloo2_out, s->d_bloom, s->tai() / 8, cudaMemcpyDeviceToHost, s->stream));
  if (bloo1_out) CU(cudaMemcpyAsync(bloo1_out, s->d_bloom1, s->tai() / 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return check_launch("bloom_split");
}

int faucet_session_read_bloom(faucet_session* s, uint8_t* bloo2_out) {
  if (!s->d_bloom) return fail(FAUCET_E_STATE, "no bloo2 on the device");
  CU(cudaMemcpyAsync(bloo2_out, s->d_bloom, s->tai() / 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int faucet_session_set_bloom(faucet_session* s, const uint8_t* bloo2) {
  int rc;
  if (!s->d_bloom && (rc = dmalloc(&s->d_bloom, s->tai() / 32))) return rc;
  CU(cudaMemcpyAsync(s->d_bloom, bloo2, s->tai() / 8, cudaMemcpyHostToDevice, s->stream));
  s->memo_dirty = true;
  return 0;
}

int faucet_session_load_stats(faucet_session* s, faucet_load_stats* out, uint64_t total_lines) {
  LoadCounters c;
  CU(cudaMemcpyAsync(&c, s->d_lctr, sizeof c, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  out->kmers = c.kmers;
  out->unambiguous_reads = c.segments;
  out->fresh_kmers = c.fresh;
  const uint64_t period = s->fastq ? 4 : 2;
  out->reads_processed = (total_lines + period - 1) / period;
  // Bloom::weight() divides two floats (utils/Bloom.cpp:191-203)
  out->weight1 = (double)((float)(long)c.weight1 / (float)s->tai());
  out->weight2 = (double)((float)(long)c.weight2 / (float)s->tai());
  return 0;
}

// The k-mer cache: one buffer, used by pass 2 (kind 2: extension masks, scan.cuh) and, for filters that do not fit L2, by
// pass 1 (kind 1: saturated k-mers, load.cuh).  Cleared when its user changes or when what it caches went stale.  Sized from the filter (~ estimated k-mers); a cache, so a
// short table only costs recomputation.
static int memo_acquire(faucet_session* s, int kind, int* qbits) {
  if (!s->d_memo) {
    uint64_t want = std::max<uint64_t>(s->tai() >> g.memo_shift, (uint64_t)1 << 20);
    want = std::min<uint64_t>(want, (uint64_t)1 << 30);
    while (want >= ((uint64_t)1 << 20) && cudaMalloc((void**)&s->d_memo, want * 8) != cudaSuccess) {
      cudaGetLastError();
      s->d_memo = nullptr;
      want >>= 1;
    }
    if (s->d_memo) { s->memo_entries = want; s->memo_dirty = true; }
  }
  if (!s->d_memo) return 0;
  if (s->memo_dirty || s->memo_kind != kind) {
    CU(cudaMemsetAsync(s->d_memo, 0xff, s->memo_entries * 8, s->stream));
    s->memo_dirty = false;
    s->memo_kind = kind;
  }
  int lg = 0;
  while (((uint64_t)1 << lg) < s->memo_entries) lg++;
  *qbits = 64 - lg;
  return 0;
}

int faucet_session_scan_flags(faucet_session* s) { return faucet_session_scan_flags_records(s, 0, s->n_recs); }

// the flag bytes of the records [r_begin, r_end) only (a sharded epoch's owner flags the records of its ordered prefix
// first, the others once the table is on its way to the other GPUs); (0, n_recs) = the whole batch
int faucet_session_scan_flags_records(faucet_session* s, uint32_t r_begin, uint32_t r_end) {
  if (!s->parsed) return fail(FAUCET_E_STATE, "faucet_session_scan_flags before faucet_session_parse");
  if (r_begin > r_end || r_end > s->n_recs) return fail(FAUCET_E_ARG, "bad record range");
  int rc = ensure_scan_buffers(s);
  if (rc) return rc;
  // byte range of those records: from the sequence line of the first one to the end of the sequence line of the last
  uint32_t b0 = 0, b1 = (uint32_t)s->n;
  if (r_begin == r_end) return 0;
  if (r_begin > 0) CU(cudaMemcpyAsync(&b0, s->d_seq_start + r_begin, 4, cudaMemcpyDeviceToHost, s->stream));
  if (r_end < s->n_recs) CU(cudaMemcpyAsync(&b1, s->d_seq_end + r_end - 1, 4, cudaMemcpyDeviceToHost, s->stream));
  if (r_begin > 0 || r_end < s->n_recs) CU(cudaStreamSynchronize(s->stream));
  ScanArgs a;
  a.inval = s->d_inval; a.packed = s->d_packed; a.w_begin = b0 / 32; a.n_words = (uint32_t)(((size_t)b1 + 31) / 32);
  a.bloom = s->d_bloom; a.wmask = (uint32_t)((s->tai() - 1) >> 5); a.k = s->k; a.j = s->j; a.n_hash = s->n_hash; a.flags = s->d_flags;
  const int grid = g.sm_count * SCAN_CTAS_PER_SM * 2;  // two full waves of resident CTAs
  a.memo = nullptr; a.memo_mask = 0; a.dbg = nullptr;
  static unsigned long long* d_dbg = nullptr;
  if (getenv("FAUCET_SCAN_DEBUG")) {
    if (!d_dbg) { cudaMalloc((void**)&d_dbg, 64); }
    cudaMemsetAsync(d_dbg, 0, 64, s->stream);
    a.dbg = d_dbg;
  }
  if (g.scan_memo) {
    int qb = 0;
    if ((rc = memo_acquire(s, 2, &qb))) return rc;
    if (s->d_memo) { a.memo = s->d_memo; a.memo_mask = s->memo_entries - 1; a.memo_qbits = qb; }
  }
  {
    KTimer kt(s, KT_SCAN);
    if (a.memo) { DISPATCH_NH(scan_flags_memo_kernel, s->n_hash, grid, SCAN_THREADS, s->stream, a); }
    else { DISPATCH_NH(scan_flags_kernel, s->n_hash, grid, SCAN_THREADS, s->stream, a); }
    s->launches++;
  }
  if (a.dbg) {
    unsigned long long h[3];
    cudaStreamSynchronize(s->stream);
    cudaMemcpy(h, a.dbg, 24, cudaMemcpyDeviceToHost);
    fprintf(stderr, "scan_flags debug: %llu lanes missed, %llu warp passes through the long path, %llu failed inserts\n", h[0], h[1], h[2]);
  }
  return check_launch("scan_flags");
}

// ---- pass 2, stream-order part: GPU junction table (stitch.cuh) ------------------------------------

static int stitch_alloc_into(faucet_session* s, unsigned long long cap, unsigned long long** keys, uint32_t** recs,
                             unsigned long long** stamps) {
  *keys = nullptr; *recs = nullptr; *stamps = nullptr;
  int rc;
  if ((rc = dmalloc(keys, cap + 1)) || (rc = dmalloc(recs, (cap + 1) * REC_WORDS)) || (rc = dmalloc(stamps, cap + 1))) {
    cudaFree(*keys); cudaFree(*recs); cudaFree(*stamps);
    *keys = nullptr; *recs = nullptr; *stamps = nullptr;
    return rc;
  }
  CU(cudaMemsetAsync(*keys, 0xff, (cap + 1) * 8, s->stream));
  CU(cudaMemsetAsync(*recs, 0, (cap + 1) * REC_WORDS * 4, s->stream));
  CU(cudaMemsetAsync(*stamps, 0, (cap + 1) * 8, s->stream));
  return 0;
}

static int stitch_alloc_table(faucet_session* s, unsigned long long cap) {
  int rc = stitch_alloc_into(s, cap, &s->d_keys, &s->d_recs, &s->d_jstamps);
  if (rc) return rc;
  s->tbl_cap = cap;
  return 0;
}

// doubles the table by rehash; on failure the old table stays in place
static int stitch_grow_table(faucet_session* s) {
  const unsigned long long ocap = s->tbl_cap;
  if (ocap * 2 >= (1ull << 31)) return fail(FAUCET_E_NOMEM, "junction table would exceed 2^31 slots");
  unsigned long long *nk, *ns;
  uint32_t* nr;
  int rc = stitch_alloc_into(s, ocap * 2, &nk, &nr, &ns);
  if (rc) return rc;
  stitch_rehash_kernel<<<g.sm_count * 8, 256, 0, s->stream>>>(s->d_keys, s->d_recs, s->d_jstamps, ocap, nk, nr, ns, ocap * 2);
  s->launches++;
  CU(cudaStreamSynchronize(s->stream));
  if ((rc = check_launch("stitch_rehash"))) { cudaFree(nk); cudaFree(nr); cudaFree(ns); return rc; }
  cudaFree(s->d_keys); cudaFree(s->d_recs); cudaFree(s->d_jstamps);
  s->d_keys = nk; s->d_recs = nr; s->d_jstamps = ns; s->tbl_cap = ocap * 2;
  return 0;
}

int faucet_session_stitch_begin(faucet_session* s, int paired_ends, int no_cleaning, uint8_t* short_pf,
                                int spf_log2_tai, int spf_n_hash, uint8_t* long_pf, int lpf_log2_tai,
                                int lpf_n_hash) {
  int rc;
  s->paired = paired_ends; s->no_cleaning = no_cleaning;
  if (!s->d_st && ((rc = dmalloc(&s->d_st, 1)) || (rc = dmalloc(&s->d_st_snap, 1)) || (rc = dmalloc(&s->d_count, 1)))) return rc;
  if (!s->d_keys) {
    if ((rc = stitch_alloc_table(s, g.table_cap0))) return rc;
  } else {  // a new scan starts from an empty JunctionMap
    CU(cudaMemsetAsync(s->d_keys, 0xff, (s->tbl_cap + 1) * 8, s->stream));
    CU(cudaMemsetAsync(s->d_recs, 0, (s->tbl_cap + 1) * REC_WORDS * 4, s->stream));
  }
  s->w_max = g.stitch_w_max;
  if (!s->stitch_grid) {
    int per_sm = 0;
    const size_t smem = STITCH_WARPS * sizeof(WarpScratch);
#if FAUCET_STITCH_THREADS >= 512
    s->stitch_fn = (const void*)stitch_kernel<1>;
#else
    s->stitch_fn = g.stitch_blocks == 4 ? (const void*)stitch_kernel<4> : g.stitch_blocks == 3 ? (const void*)stitch_kernel<3> : (const void*)stitch_kernel<2>;
#endif
    CU(cudaFuncSetAttribute(s->stitch_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s->stitch_fn, STITCH_THREADS, smem));
    if (per_sm < 1) return fail(FAUCET_E_CUDA, "stitch_kernel cannot be made resident");
    s->stitch_grid = per_sm * g.sm_count;
  }
  // one record per warp per round
  s->w_max = std::min<uint32_t>(s->w_max, (uint32_t)s->stitch_grid * STITCH_WARPS);
  if (!s->d_res) {
    if ((rc = dmalloc(&s->d_res, (size_t)1 << g.res_log2))) return rc;
    CU(cudaMemsetAsync(s->d_res, 0xff, ((size_t)1 << g.res_log2) * 4, s->stream));
  }
  if (!s->d_jslot && (rc = dmalloc(&s->d_jslot, ((size_t)1 << g.res_log2) / 32 + 1))) return rc;
  CU(cudaMemsetAsync(s->d_jslot, 0, (((size_t)1 << g.res_log2) / 32 + 1) * 4, s->stream));
  if (s->w_max > s->deferred_cap) {
    for (int i = 0; i < 2; i++) {
      cudaFree(s->d_deferred[i]); s->d_deferred[i] = nullptr;
      if ((rc = dmalloc(&s->d_deferred[i], s->w_max))) return rc;
    }
    s->deferred_cap = s->w_max;
  }
  CU(cudaMemsetAsync(s->d_st, 0, sizeof(StitchState), s->stream));
  unsigned int w0 = std::min(g.stitch_w0, s->w_max);
  CU(cudaMemcpyAsync(&s->d_st->W, &w0, 4, cudaMemcpyHostToDevice, s->stream));
  s->h_st = StitchState();
  s->h_st.W = w0;
  // short pair filter: adds only (src/ReadScanner.cpp:208-225) => atomicOr on a device copy
  cudaFree(s->d_spf); s->d_spf = nullptr; s->h_spf = nullptr;
  cudaFree(s->d_spf_snap); s->d_spf_snap = nullptr;
  if (short_pf && !no_cleaning) {
    size_t words = ((size_t)1 << spf_log2_tai) / 32;
    if (words == 0) words = 1;
    if ((rc = dmalloc(&s->d_spf, words)) || (rc = dmalloc(&s->d_spf_snap, words))) return rc;
    CU(cudaMemcpyAsync(s->d_spf, short_pf, ((size_t)1 << spf_log2_tai) / 8, cudaMemcpyHostToDevice, s->stream));
    s->h_spf = short_pf; s->spf_log2 = spf_log2_tai; s->spf_nh = spf_n_hash;
  }
  // long pair filter: sequential per mate pair => host, fed by the extension lists the kernel emits
  s->lpf = LongPairFilter();
  if (long_pf && paired_ends && !no_cleaning) {
    s->lpf.init(long_pf, lpf_log2_tai, lpf_n_hash, s->k);
    if (!s->d_ext) {
      s->ext_cap = g.ext_cap0;
      if ((rc = dmalloc(&s->d_ext, s->ext_cap))) return rc;
    }
  }
  s->h_ext.clear();
  s->rec_base = 0;
  s->n_recs_out = 0;
  std::memset(&s->sstats, 0, sizeof s->sstats);
  s->ep_size = g.epoch0;
  s->ep_exact = true;
  s->ep = faucet_session::EpochStats();
  s->stitching = true;
  return 0;
}

static void stitch_fill_args(faucet_session* s, StitchArgs& a) {
  std::memset(&a, 0, sizeof a);
  a.inval = s->d_inval; a.packed = s->d_packed; a.flags = s->d_flags;
  a.seq_start = s->d_seq_start; a.seq_end = s->d_seq_end; a.n_recs = s->n_recs; a.rec_base = s->rec_base; a.gid = s->cur_gid;
  a.k = s->k; a.j = s->j; a.spacer = s->max_spacer; a.no_cleaning = s->no_cleaning; a.paired = s->paired;
  a.keys = s->d_keys; a.recs = s->d_recs; a.stamps = s->d_jstamps; a.cap = s->tbl_cap;
  a.res = s->d_res; a.res_mask = (uint32_t)(((size_t)1 << g.res_log2) - 1);
  a.deferred[0] = s->d_deferred[0]; a.deferred[1] = s->d_deferred[1];
  a.st = s->d_st; a.special = &s->d_st->special; a.jslot = s->d_jslot;
  a.spf = s->d_spf; a.spf_mask = s->d_spf ? ((1ull << s->spf_log2) - 1) : 0; a.spf_nh = s->spf_nh;
  a.ext = s->lpf.enabled() ? s->d_ext : nullptr; a.ext_cap = s->ext_cap;
  a.w_min = std::min<uint32_t>(64, s->w_max); a.w_max = s->w_max;
  a.shrink_den = g.stitch_shrink_den; a.grow_den = g.stitch_grow_den;
  a.cov_stride = REC_WORDS; a.cov_off = REC_COV;
  a.lazy = g.dry_lazy == 1 || (g.dry_lazy == 2 && (s->tbl_cap + 1) * 8 <= ((size_t)64 << 20)) ? 1 : 0;
}

// moves what the kernels left in the extension-list buffer to the host
static int stitch_drain_ext(faucet_session* s) {
  unsigned long long used = 0;
  CU(cudaMemcpyAsync(&used, &s->d_st->ext_used, 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (used) {
    const size_t at = s->h_ext.size();
    s->h_ext.resize(at + used);
    CU(cudaMemcpy(s->h_ext.data() + at, s->d_ext, used * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemsetAsync(&s->d_st->ext_used, 0, 8, s->stream));
  }
  return 0;
}

// The ordered kernel over the entries [begin, end) of `list` (NULL: the records themselves), relaunched until done.
// allow_grow: grow the table when a round could overfill it; otherwise *need_grow is set and the caller decides
// (inside a classify epoch the slots of the snapshot must stay valid).
static int stitch_run_ordered(faucet_session* s, const uint32_t* list, uint32_t begin, uint32_t end, bool mark_dirty,
                              bool allow_grow, bool* need_grow) {
  if (need_grow) *need_grow = false;
  struct { unsigned int next, nd[2]; } z = {begin, {0, 0}};
  CU(cudaMemcpyAsync(&s->d_st->next, &z, sizeof z, cudaMemcpyHostToDevice, s->stream));
  while (true) {
    StitchArgs a;
    stitch_fill_args(s, a);
    a.n_recs = end; a.list = list;
    a.dirty = mark_dirty ? s->d_dirty : nullptr; a.dirty_max = mark_dirty ? s->d_dirty_max : nullptr;
    void* params[] = {&a};
    {
      KTimer kt(s, KT_STITCH);
      CU(cudaLaunchCooperativeKernel(s->stitch_fn, dim3(s->stitch_grid), dim3(STITCH_THREADS), params,
                                     STITCH_WARPS * sizeof(WarpScratch), s->stream));
      s->launches++;
    }
    CU(cudaMemcpyAsync(&s->h_st, s->d_st, sizeof(StitchState), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    int rc = check_launch("stitch");
    if (rc) return rc;
    const unsigned int status = s->h_st.status;
    if (status == ST_DONE) break;
    // an aborted round leaves its reservations behind
    CU(cudaMemsetAsync(s->d_res, 0xff, ((size_t)1 << g.res_log2) * 4, s->stream));
    if (status == ST_GROW_TABLE) {
      if (!allow_grow) { *need_grow = true; return 0; }
      if ((rc = stitch_grow_table(s))) return rc;
    } else if (status == ST_DRAIN_EXT) {
      if (s->h_st.ext_used == 0) {  // one round alone does not fit: a bigger buffer
        cudaFree(s->d_ext); s->d_ext = nullptr;
        s->ext_cap *= 2;
        if ((rc = dmalloc(&s->d_ext, s->ext_cap))) return rc;
      } else if ((rc = stitch_drain_ext(s))) {
        return rc;
      }
    } else {
      return fail(FAUCET_E_CUDA, "stitch kernel returned an unknown status");
    }
  }
  return 0;
}

// in-place exclusive prefix sum of n u32 (stitch.cuh's scan kernels); sums = scratch of n / SCAN_CHUNK + 2 entries
static void exclusive_scan_u32(faucet_session* s, uint32_t* data, unsigned long long n, uint32_t* sums) {
  const unsigned long long nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  scan_reduce_kernel<<<(unsigned)nb, 256, 0, s->stream>>>(data, n, sums);
  scan_sums_kernel<<<1, 1024, 0, s->stream>>>(sums, nb);
  scan_apply_kernel<<<(unsigned)nb, 256, 0, s->stream>>>(data, n, sums);
  s->launches += 3;
}

static int stitch_run_ordered(faucet_session* s, const uint32_t* list, uint32_t begin, uint32_t end, bool mark_dirty,
                              bool allow_grow, bool* need_grow);

static int flow_init(faucet_session* s) {
  if (s->flow_fn) return 0;
  int rc, per_sm = 0;
  const size_t smem = STITCH_WARPS * sizeof(WarpScratch);
  s->flow_fn = (const void*)stitch_flow_kernel<FAUCET_FLOW_BLOCKS>;
  CU(cudaFuncSetAttribute(s->flow_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s->flow_fn, STITCH_THREADS, smem));
  if (per_sm < 1) return fail(FAUCET_E_CUDA, "stitch_flow_kernel cannot be made resident");
  s->flow_grid = per_sm * g.sm_count;
  if ((rc = dmalloc(&s->d_fbig, 64))) return rc;  // [0] = long-line count, [32] = the ticket counter (its own 128-byte line)
  return 0;
}

static int flow_ensure(faucet_session* s, uint32_t n) {
  if (n <= s->flow_cap) return 0;
  int rc;
  cudaFree(s->d_frows); cudaFree(s->d_fpreds); cudaFree(s->d_fdone); cudaFree(s->d_fcounts); cudaFree(s->d_fcount_sums);
  s->d_frows = s->d_fpreds = s->d_fdone = s->d_fcounts = s->d_fcount_sums = nullptr;
  s->prep_valid = false;
  s->flow_cap = (size_t)n + n / 4 + 1024;
  if ((rc = dmalloc(&s->d_frows, s->flow_cap * ROW_WORDS)) || (rc = dmalloc(&s->d_fpreds, s->flow_cap * ROW_WORDS)) ||
      (rc = dmalloc(&s->d_fdone, s->flow_cap)) || (rc = dmalloc(&s->d_fcounts, s->flow_cap + 1)) ||
      (rc = dmalloc(&s->d_fcount_sums, s->flow_cap / SCAN_CHUNK + 4)))
    return rc;
  return 0;
}

// The dependency sort of the dataflow executor (flow.cuh) for the entries [b0, b0 + n) of `list` (NULL: the records
// themselves) of the parsed batch: rows and predecessors into the session's buffers.  A pure function of the text
// planes -- it does not touch the junction table -- so a whole batch can be prepared ahead of its stitch: during pass 1
// (retained planes), or by the GPU that owns the shard (multi-GPU).  *big: a line has more slots than a row holds.
static int flow_prepare(faucet_session* s, const uint32_t* list, uint32_t b0, uint32_t n, bool* big_out) {
  int rc;
  if ((rc = flow_init(s)) || (rc = flow_ensure(s, n))) return rc;
  const int grid = g.sm_count * 8;
  StitchArgs a;
  stitch_fill_args(s, a);
  a.list = list ? list + b0 : nullptr;
  FlowArgs f;
  std::memset(&f, 0, sizeof f);
  f.n = n; f.begin = b0; f.rows = s->d_frows; f.preds = s->d_fpreds; f.done = s->d_fdone; f.counts = s->d_fcounts; f.big = s->d_fbig;
  uint32_t tail[2] = {0, 0};  // last count and its prefix: their sum is the number of pairs
  unsigned int big = 0;
  {
    KTimer kt(s, KT_FLOW_PREP);
    CU(cudaMemsetAsync(s->d_fbig, 0, 4, s->stream));
    flow_rows_kernel<<<grid, DRY_THREADS, 0, s->stream>>>(a, f);
    CU(cudaMemcpyAsync(&tail[0], s->d_fcounts + n - 1, 4, cudaMemcpyDeviceToHost, s->stream));
    exclusive_scan_u32(s, s->d_fcounts, n, s->d_fcount_sums);
    CU(cudaMemcpyAsync(&tail[1], s->d_fcounts + n - 1, 4, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(&big, s->d_fbig, 4, cudaMemcpyDeviceToHost, s->stream));
    s->launches++;
  }
  CU(cudaStreamSynchronize(s->stream));
  if ((rc = check_launch("flow_rows"))) return rc;
  *big_out = big != 0;
  if (big) return 0;
  const uint32_t n_pairs = tail[0] + tail[1];
  f.n_pairs = n_pairs;
  f.n_sub = (n_pairs + RADIX_SUB - 1) / RADIX_SUB;
  if (n_pairs > s->fpairs_cap) {
    cudaFree(s->d_fpairs); cudaFree(s->d_fpairs2); s->d_fpairs = s->d_fpairs2 = nullptr;
    s->fpairs_cap = (size_t)n_pairs + n_pairs / 4 + 1024;
    if ((rc = dmalloc(&s->d_fpairs, s->fpairs_cap)) || (rc = dmalloc(&s->d_fpairs2, s->fpairs_cap))) return rc;
  }
  if ((size_t)f.n_sub * 256 > s->fhist_cap) {
    cudaFree(s->d_fhist); cudaFree(s->d_fhist_sums); s->d_fhist = s->d_fhist_sums = nullptr;
    s->fhist_cap = (size_t)f.n_sub * 256 + 1024;
    if ((rc = dmalloc(&s->d_fhist, s->fhist_cap)) || (rc = dmalloc(&s->d_fhist_sums, s->fhist_cap / SCAN_CHUNK + 4))) return rc;
  }
  f.pairs = s->d_fpairs; f.pairs2 = s->d_fpairs2; f.hist = s->d_fhist;
  {
    KTimer kt(s, KT_FLOW_PREP);
    flow_pairs_kernel<<<grid, 256, 0, s->stream>>>(f);
    s->launches++;
    if (n_pairs) {
      const unsigned rgrid = (f.n_sub + RADIX_THREADS / 32 - 1) / (RADIX_THREADS / 32);
      for (int shift = 0; shift < g.res_log2; shift += 8) {  // stable LSD passes over the slot
        f.shift = shift;
        radix_hist_kernel<<<rgrid, RADIX_THREADS, 0, s->stream>>>(f);
        exclusive_scan_u32(s, s->d_fhist, (unsigned long long)f.n_sub * 256, s->d_fhist_sums);
        radix_scatter_kernel<<<rgrid, RADIX_THREADS, 0, s->stream>>>(f);
        std::swap(f.pairs, f.pairs2);
        s->launches += 2;
      }
      flow_preds_kernel<<<grid, 256, 0, s->stream>>>(f);
      s->launches++;
    }
  }
  return check_launch("flow_prepare");
}

// prepares the whole parsed batch ahead of its stitch (stage API; pass 1 with retained planes; shard owners)
int faucet_session_flow_prepare(faucet_session* s) { return faucet_session_flow_prepare_records(s, s->n_recs, 0); }

// ... or its first n records (a sorted prefix serves every shorter prefix: predecessors are earlier records).
// concurrent != 0: the sort runs on a second stream, next to what is queued on the session's (a sharded epoch's owner
// sorts the records of its ordered prefix while scan_flags flags them: one is memory-bound, the other instruction-bound);
// the call returns when the sort is complete.
int faucet_session_flow_prepare_records(faucet_session* s, uint32_t n, int concurrent) {
  if (!s->parsed) return fail(FAUCET_E_STATE, "flow_prepare before parse");
  if (n > s->n_recs) return fail(FAUCET_E_ARG, "bad record count");
  s->prep_valid = false;
  if (g.stitch_exec == 0 || n == 0 || n > g.flow_chunk) return 0;  // nothing to keep: the stitch sorts itself
  bool big = false;
  cudaStream_t main_stream = s->stream;
  if (concurrent) {
    if (!s->aux_stream) CU(cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
    s->stream = s->aux_stream;
  }
  int rc = flow_prepare(s, nullptr, 0, n, &big);
  if (!rc && concurrent && cudaStreamSynchronize(s->aux_stream) != cudaSuccess) rc = fail(FAUCET_E_CUDA, "flow_prepare (second stream)");
  s->stream = main_stream;
  if (rc) return rc;
  s->prep_valid = true; s->prep_n = n; s->prep_big = big;
  s->prep_rows = s->d_frows; s->prep_preds = s->d_fpreds;
  return 0;
}

// The ordered execution of the entries [begin, end) of `list` (NULL: the records themselves) by the dataflow
// executor (flow.cuh): dependency sort (unless the batch came with one), then one persistent kernel, in chunks of
// g.flow_chunk records.  Same contract as stitch_run_ordered, which still takes the lists that hold a line with
// more slots than a row.
static int stitch_run_flow(faucet_session* s, const uint32_t* list, uint32_t begin, uint32_t end, bool mark_dirty,
                           bool allow_grow, bool* need_grow) {
  if (need_grow) *need_grow = false;
  if (g.stitch_exec == 0) return stitch_run_ordered(s, list, begin, end, mark_dirty, allow_grow, need_grow);
  int rc;
  if ((rc = flow_init(s))) return rc;
  for (uint32_t b0 = begin; b0 < end;) {
    const uint32_t n = std::min<uint32_t>(end - b0, g.flow_chunk);
    // (a dependency sort of the whole batch also serves any prefix of it: predecessors are earlier records)
    const bool prepared = !list && b0 == 0 && s->prep_valid && n <= s->prep_n && end <= g.flow_chunk;
    bool big = false;
    const uint32_t *rows = s->prep_rows, *preds = s->prep_preds;
    if (prepared) {
      big = s->prep_big;
      if ((rc = flow_ensure(s, n))) return rc;  // (the done flags; a prepared sort in the session's own buffers fits already)
      if (!s->prep_valid) return fail(FAUCET_E_STATE, "prepared dependency sort lost");
    } else {
      if ((rc = flow_prepare(s, list, b0, n, &big))) return rc;
      s->prep_valid = false;  // the buffers now describe this chunk
      rows = s->d_frows; preds = s->d_fpreds;
    }
    if (big) {  // a line with more slots than a row holds: this chunk goes through the rounds
      if ((rc = stitch_run_ordered(s, list, b0, b0 + n, mark_dirty, allow_grow, need_grow))) return rc;
      if (need_grow && *need_grow) return 0;
      b0 += n;
      continue;
    }
    StitchArgs a;
    stitch_fill_args(s, a);
    a.list = list ? list + b0 : nullptr;
    a.n_recs = n;
    a.dirty = mark_dirty ? s->d_dirty : nullptr; a.dirty_max = mark_dirty ? s->d_dirty_max : nullptr;
    FlowArgs f;
    std::memset(&f, 0, sizeof f);
    f.n = n; f.begin = b0; f.rows = const_cast<uint32_t*>(rows); f.preds = const_cast<uint32_t*>(preds); f.done = s->d_fdone;
    f.next = s->d_fbig + 32;
    CU(cudaMemsetAsync(s->d_fdone, 0, (size_t)n * 4, s->stream));
    // ---- run; relaunched after the table grew / the extension lists were drained
    while (true) {
      struct { unsigned int next, nd[2], W, round, status; } z = {0, {0, 0}, s->h_st.W, s->h_st.round, ST_DONE};
      CU(cudaMemcpyAsync(&s->d_st->next, &z, sizeof z, cudaMemcpyHostToDevice, s->stream));
      CU(cudaMemsetAsync(s->d_fbig + 32, 0, 4, s->stream));
      void* params[] = {&a, &f};
      {
        KTimer kt(s, KT_STITCH);
        CU(cudaLaunchCooperativeKernel(s->flow_fn, dim3(s->flow_grid), dim3(STITCH_THREADS), params,
                                       STITCH_WARPS * sizeof(WarpScratch), s->stream));
        s->launches++;
      }
      CU(cudaMemcpyAsync(&s->h_st, s->d_st, sizeof(StitchState), cudaMemcpyDeviceToHost, s->stream));
      CU(cudaStreamSynchronize(s->stream));
      if ((rc = check_launch("stitch_flow"))) return rc;
      const unsigned int status = s->h_st.status;
      if (status == ST_DONE) break;
      if (status == ST_GROW_TABLE) {
        if (!allow_grow) { *need_grow = true; return 0; }
        if ((rc = stitch_grow_table(s))) return rc;
        a.keys = s->d_keys; a.recs = s->d_recs; a.stamps = s->d_jstamps; a.cap = s->tbl_cap;
      } else if (status == ST_DRAIN_EXT) {
        if (s->h_st.ext_used == 0) {
          cudaFree(s->d_ext); s->d_ext = nullptr;
          s->ext_cap *= 2;
          if ((rc = dmalloc(&s->d_ext, s->ext_cap))) return rc;
          a.ext = s->d_ext; a.ext_cap = s->ext_cap;
        } else if ((rc = stitch_drain_ext(s))) {
          return rc;
        }
      } else {
        return fail(FAUCET_E_CUDA, "stitch flow kernel returned an unknown status");
      }
    }
    b0 += n;
  }
  return 0;
}

static int stitch_ensure_epoch_buffers(faucet_session* s, uint32_t m) {
  int rc;
  if (!s->d_in_exact || s->n_recs > s->in_exact_cap) {  // (allocated even for an empty batch: a rank of a sharded job may hold no record)
    cudaFree(s->d_in_exact); s->d_in_exact = nullptr;
    s->in_exact_cap = (size_t)s->n_recs + s->n_recs / 4 + 1024;
    if ((rc = dmalloc(&s->d_in_exact, s->in_exact_cap))) return rc;
  }
  if (m > s->list_cap) {
    cudaFree(s->d_list); cudaFree(s->d_eprefix); cudaFree(s->d_eprefix_sums);
    s->d_list = nullptr; s->d_eprefix = nullptr; s->d_eprefix_sums = nullptr;
    s->list_cap = (size_t)m + m / 4 + 1024;
    if ((rc = dmalloc(&s->d_list, 2 * s->list_cap)) || (rc = dmalloc(&s->d_eprefix, s->list_cap)) ||  // (two lists: sharded epochs alternate)
        (rc = dmalloc(&s->d_eprefix_sums, s->list_cap / SCAN_CHUNK + 2)))
      return rc;
  }
  if (!s->d_dirty && ((rc = dmalloc(&s->d_dirty, (size_t)1 << g.res_log2)) || (rc = dmalloc(&s->d_dirty_max, (size_t)1 << g.res_log2)))) return rc;
  if ((size_t)m > s->rows_cap) {
    cudaFree(s->d_rows); s->d_rows = nullptr;
    s->rows_cap = (size_t)m + m / 4 + 1024;
    if ((rc = dmalloc(&s->d_rows, s->rows_cap * ROW_WORDS))) return rc;
  }
  if (!s->dry_ready) {
    CU(cudaFuncSetAttribute((const void*)stitch_dry_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DRY_WARPS * sizeof(DryScratch))));
    s->dry_ready = true;
  }
  if (s->snap_cap != s->tbl_cap) {
    cudaFree(s->d_snap_keys); cudaFree(s->d_snap_recs);
    s->d_snap_keys = nullptr; s->d_snap_recs = nullptr; s->snap_cap = 0;
    if ((rc = dmalloc(&s->d_snap_keys, s->tbl_cap + 1)) || (rc = dmalloc(&s->d_snap_recs, (s->tbl_cap + 1) * REC_WORDS))) return rc;
    s->snap_cap = s->tbl_cap;
  }
  return 0;
}

// T0 := the live table (keys and records; creation stamps of T0 keys never change), the stitch state and the short pair filter
static int stitch_snapshot(faucet_session* s) {
  CU(cudaMemcpyAsync(s->d_snap_keys, s->d_keys, (s->tbl_cap + 1) * 8, cudaMemcpyDeviceToDevice, s->stream));
  CU(cudaMemcpyAsync(s->d_snap_recs, s->d_recs, (s->tbl_cap + 1) * REC_WORDS * 4, cudaMemcpyDeviceToDevice, s->stream));
  CU(cudaMemcpyAsync(s->d_st_snap, s->d_st, sizeof(StitchState), cudaMemcpyDeviceToDevice, s->stream));
  if (s->d_spf) CU(cudaMemcpyAsync(s->d_spf_snap, s->d_spf, std::max<size_t>(4, ((size_t)1 << s->spf_log2) / 8), cudaMemcpyDeviceToDevice, s->stream));
  return 0;
}
static int stitch_restore(faucet_session* s, size_t h_ext_mark) {
  CU(cudaMemcpyAsync(s->d_keys, s->d_snap_keys, (s->tbl_cap + 1) * 8, cudaMemcpyDeviceToDevice, s->stream));
  CU(cudaMemcpyAsync(s->d_recs, s->d_snap_recs, (s->tbl_cap + 1) * REC_WORDS * 4, cudaMemcpyDeviceToDevice, s->stream));
  CU(cudaMemcpyAsync(s->d_st, s->d_st_snap, sizeof(StitchState), cudaMemcpyDeviceToDevice, s->stream));
  if (s->d_spf) CU(cudaMemcpyAsync(s->d_spf, s->d_spf_snap, std::max<size_t>(4, ((size_t)1 << s->spf_log2) / 8), cudaMemcpyDeviceToDevice, s->stream));
  s->h_ext.resize(h_ext_mark);
  return 0;
}

// One epoch of the classify / execute / verify / apply scheme (stitch.cuh) over the records [r, r + m).
static int stitch_epoch_classify(faucet_session* s, uint32_t r, uint32_t m, uint32_t* n_exact_out) {
  int rc;
  // headroom the ordered kernel itself would ask for, before the snapshot pins the slots
  while (s->h_st.n_entries + s->h_st.max_need * (g.stitch_exec ? (unsigned long long)s->stitch_grid * STITCH_WARPS : s->w_max) > s->tbl_cap / 2)
    if ((rc = stitch_grow_table(s))) return rc;
  if ((rc = stitch_ensure_epoch_buffers(s, m))) return rc;
  const int grid = g.sm_count * 8;
  const size_t dry_smem = DRY_WARPS * sizeof(DryScratch);
  StitchArgs d;
  stitch_fill_args(s, d);
  // classify walks the live table, which is T0 until the exact set runs; with no pair filter to feed, the quiet
  // records commit their coverage counts right away
  const bool fused = !d.spf && !d.ext;
  d.in_exact = s->d_in_exact; d.r_begin = r; d.r_end = r + m; d.dirty = s->d_dirty; d.dirty_max = s->d_dirty_max;
  d.rows = s->d_rows; d.rows_base = r; d.recheck = g.epoch_recheck ? 1 : 0;
  d.taint_mark = fused ? EX_COMMITTED : EX_MEMBER;
  // a walk of the records flagged `want`, on the live table or on the snapshot
  auto dry = [&](int mode, uint8_t want, uint8_t after, bool live, uint32_t x0, uint32_t x1) {
    StitchArgs w = d;
    w.dry_mode = mode; w.want_flag = want; w.flag_after = after; w.r_begin = x0; w.r_end = x1;
    if (!live) { w.keys = s->d_snap_keys; w.recs = s->d_snap_recs; w.special = &s->d_st_snap->special; }
    w.cov_out = s->d_recs;
    KTimer kt(s, KT_DRY);
    stitch_dry_kernel<<<grid, DRY_THREADS, dry_smem, s->stream>>>(w);
    s->launches++;
  };
  CU(cudaMemsetAsync(s->d_in_exact + r, 0, m, s->stream));
  dry(fused ? DRY_CLASSIFY_APPLY : DRY_CLASSIFY, 0, 0, true, r, r + m);
  if ((rc = stitch_snapshot(s))) return rc;  // T0 (+ the counts the quiet records just added)
  const size_t h_ext_mark = s->h_ext.size();
  const unsigned n_blocks = (m + SCAN_CHUNK - 1) / SCAN_CHUNK;
  uint32_t n_exact = 0, n_prev = 0;
  for (int iter = 0;; iter++) {
    {
      KTimer kt(s, KT_VERIFY);
      exact_flags_kernel<<<grid, 256, 0, s->stream>>>(s->d_in_exact, r, m, s->d_eprefix);
      scan_reduce_kernel<<<n_blocks, 256, 0, s->stream>>>(s->d_eprefix, m, s->d_eprefix_sums);
      scan_sums_kernel<<<1, 1024, 0, s->stream>>>(s->d_eprefix_sums, n_blocks);
      scan_apply_kernel<<<n_blocks, 256, 0, s->stream>>>(s->d_eprefix, m, s->d_eprefix_sums);
      exact_list_kernel<<<grid, 256, 0, s->stream>>>(s->d_in_exact, r, m, s->d_eprefix, s->d_list, s->d_count);
      s->launches += 5;
    }
    CU(cudaMemcpyAsync(&n_exact, s->d_count, 4, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if ((rc = check_launch("stitch_list"))) return rc;
    if (iter == 0) s->ep.nonquiet += n_exact;
    if (n_exact == 0 || (iter > 0 && n_exact == n_prev)) break;  // nothing (more) joined: the live table is final
    if (iter > 0 && (rc = stitch_restore(s, h_ext_mark))) return rc;
    CU(cudaMemsetAsync(s->d_dirty, 0xff, ((size_t)1 << g.res_log2) * 4, s->stream));
    CU(cudaMemsetAsync(s->d_dirty_max, 0, ((size_t)1 << g.res_log2) * 4, s->stream));
    bool need_grow = false;
    if ((rc = stitch_run_flow(s, s->d_list, 0, n_exact, true, false, &need_grow))) return rc;
    s->ep.exact_runs += n_exact;
    s->ep.iterations++;
    if (need_grow) {  // the slots of T0 are about to move: undo the epoch, grow, and take it in order instead
      if ((rc = stitch_restore(s, h_ext_mark))) return rc;
      if (fused) {  // every record that committed under T0 takes it back
        exact_mark_applied_kernel<<<grid, 256, 0, s->stream>>>(s->d_in_exact, r, m);
        s->launches++;
        dry(DRY_RETRACT, EX_COMMITTED, EX_RETRACTED, false, r, r + m);
      }
      if ((rc = stitch_grow_table(s))) return rc;
      s->ep.fallbacks++;
      *n_exact_out = m;
      return stitch_run_flow(s, nullptr, r, r + m, false, true, nullptr);
    }
    {
      KTimer kt(s, KT_VERIFY);
      stitch_verify_kernel<<<grid, DRY_THREADS, 0, s->stream>>>(d);
      s->launches++;
    }
    if (d.recheck) dry(DRY_RECHECK, EX_RECHECK, 0, true, r, r + m);
    if (fused) {  // what joined had committed under T0: take that back, from the live table and from the snapshot
      StitchArgs w = d;
      w.dry_mode = DRY_RETRACT; w.want_flag = EX_COMMITTED; w.flag_after = EX_RETRACTED;
      w.keys = s->d_snap_keys; w.recs = s->d_snap_recs; w.special = &s->d_st_snap->special;
      w.cov_out = s->d_recs; w.cov_out2 = s->d_snap_recs; w.st2 = s->d_st_snap;
      KTimer kt(s, KT_DRY);
      stitch_dry_kernel<<<grid, DRY_THREADS, dry_smem, s->stream>>>(w);
      s->launches++;
    }
    n_prev = n_exact;
  }
  // commit what has not committed yet.  EX_QUIET records saw T0; EX_SETTLED records saw the live table.
  if (fused) {
    if (n_exact) {  // (no exact set: nothing was written, nobody is settled)
      dry(DRY_RETRACT, EX_SETTLED, EX_SETTLED, false, r, r + m);
      dry(DRY_APPLY, EX_SETTLED, 0, true, r, r + m);
    }
  } else {
    for (int pass = 0; pass < 2; pass++) {
      if (pass == 1 && !n_exact) break;
      const uint8_t want = pass ? EX_SETTLED : EX_QUIET;
      if (!d.ext) { dry(DRY_APPLY, want, 0, pass == 1, r, r + m); continue; }
      // bounded launches: a quiet record emits at most LAND_CAP extensions plus chunk headers
      const uint32_t per_rec = LAND_CAP + LAND_CAP / (EXT_STAGE - 1) + 2;
      for (uint32_t x = r; x < r + m;) {
        if ((rc = stitch_drain_ext(s))) return rc;
        const uint32_t fit = (uint32_t)std::max<unsigned long long>(1, s->ext_cap / per_rec);
        const uint32_t x1 = (uint32_t)std::min<uint64_t>((uint64_t)x + fit, (uint64_t)r + m);
        dry(DRY_APPLY, want, 0, pass == 1, x, x1);
        x = x1;
      }
    }
  }
  CU(cudaMemcpyAsync(&s->h_st, s->d_st, sizeof(StitchState), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if ((rc = check_launch("stitch_apply"))) return rc;
  if (s->h_st.stats[SS_DRY_ERROR]) return fail(FAUCET_E_CUDA, "stitch: apply met a record that is not quiet (internal error)");
  *n_exact_out = n_exact;
  return 0;
}

// runs the stitch over the parsed + flagged batch that is resident in the session, epoch by epoch
int faucet_session_stitch_batch(faucet_session* s) {
  if (!s->stitching) return fail(FAUCET_E_STATE, "stitch_begin not called");
  if (!s->parsed) return fail(FAUCET_E_STATE, "stitch_batch before parse");
  const bool want_ext = s->lpf.enabled();
  s->h_ext.clear();
  int rc;
  for (uint32_t r = 0; r < s->n_recs;) {
    const uint32_t m = g.epoch_mode == 0 ? s->n_recs - r : (uint32_t)std::min<uint64_t>(s->ep_size, s->n_recs - r);
    const bool classify = g.epoch_mode == 2 || (g.epoch_mode == 1 && !s->ep_exact);
    if (!classify) {
      const unsigned long long w0 = s->h_st.stats[SS_WRITERS];
      if ((rc = stitch_run_flow(s, nullptr, r, r + m, false, true, nullptr))) return rc;
      const unsigned long long writers = s->h_st.stats[SS_WRITERS] - w0;
      s->ep.exact_epochs++; s->ep.exact_runs += m;
      s->ep_size = (uint32_t)std::min<uint64_t>((uint64_t)s->ep_size * 2, g.epoch_max);
      if (writers * 100 < (unsigned long long)m * g.epoch_switch_pct) s->ep_exact = false;
    } else {
      uint32_t n_exact = 0;
      if ((rc = stitch_epoch_classify(s, r, m, &n_exact))) return rc;
      s->ep.classify_epochs++; s->ep.dry_records += m;
      // the exact set grows with the epoch (every write taints the later records of the epoch that share its slot)
      if ((uint64_t)n_exact * 2 > m) { s->ep_exact = true; s->ep_size = std::max<uint32_t>(s->ep_size / 2, g.epoch0); }
      else if ((uint64_t)n_exact * 100 > (uint64_t)m * g.epoch_shrink_pct) s->ep_size = std::max<uint32_t>(s->ep_size / 2, g.epoch0);
      else if ((uint64_t)n_exact * 100 < (uint64_t)m * g.epoch_grow_pct) s->ep_size = (uint32_t)std::min<uint64_t>((uint64_t)s->ep_size * 2, g.epoch_max);
    }
    r += m;
  }
  if (want_ext) {
    if ((rc = stitch_drain_ext(s))) return rc;
    s->lpf.process_batch(s->h_ext.data(), s->h_ext.size(), s->n_recs, s->rec_base);
  }
  s->rec_base += s->n_recs;
  return 0;
}

// gathers the junction map (sorted into creation order) and the counters; writes the short pair
// filter back to the caller's array
static int stitch_finish(faucet_session* s) {
  StitchState st;
  CU(cudaMemcpyAsync(&st, s->d_st, sizeof st, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  const size_t n = (size_t)st.n_entries;
  int rc;
  // creation order on the device: counting sort over record indices (stitch.cuh)
  const unsigned long long n_hist = s->rec_base + 1;
  const unsigned long long n_blocks = (n_hist + SCAN_CHUNK - 1) / SCAN_CHUNK;
  if (n_hist > s->hist_cap) {
    cudaFree(s->d_hist); cudaFree(s->d_hist_sums); s->d_hist = nullptr; s->d_hist_sums = nullptr;
    s->hist_cap = n_hist + n_hist / 4;
    if ((rc = dmalloc(&s->d_hist, s->hist_cap)) || (rc = dmalloc(&s->d_hist_sums, s->hist_cap / SCAN_CHUNK + 2))) return rc;
  }
  if (n > s->out_cap) {
    cudaFree(s->d_out); s->d_out = nullptr;
    s->out_cap = n + n / 4 + 1024;
    if ((rc = dmalloc(&s->d_out, s->out_cap))) return rc;
  }
  static_assert(sizeof(JunctionOut) == sizeof(faucet_junction_rec), "record layouts must agree");
  if (n) {
    CU(cudaMemsetAsync(s->d_hist, 0, n_hist * 4, s->stream));
    stitch_count_kernel<<<g.sm_count * 8, 256, 0, s->stream>>>(s->d_keys, s->d_jstamps, s->tbl_cap, st.special, s->d_hist);
    scan_reduce_kernel<<<(unsigned)n_blocks, 256, 0, s->stream>>>(s->d_hist, n_hist, s->d_hist_sums);
    scan_sums_kernel<<<1, 1024, 0, s->stream>>>(s->d_hist_sums, n_blocks);
    scan_apply_kernel<<<(unsigned)n_blocks, 256, 0, s->stream>>>(s->d_hist, n_hist, s->d_hist_sums);
    stitch_emit_kernel<<<g.sm_count * 8, 256, 0, s->stream>>>(s->d_keys, s->d_recs, s->d_jstamps, s->tbl_cap, st.special,
                                                            s->d_hist, (JunctionOut*)s->d_out);
    s->launches += 5;
  }
  if (n > s->h_recs_cap) {
    if (s->h_recs) cudaFreeHost(s->h_recs);
    s->h_recs = nullptr; s->h_recs_cap = 0;
    const size_t want = n + n / 4 + 1024;
    CU(cudaHostAlloc((void**)&s->h_recs, want * sizeof(faucet_junction_rec), cudaHostAllocDefault));
    s->h_recs_cap = want;
  }
  s->n_recs_out = n;
  if (n) CU(cudaMemcpyAsync(s->h_recs, s->d_out, n * sizeof(JunctionOut), cudaMemcpyDeviceToHost, s->stream));
  if (s->h_spf) CU(cudaMemcpyAsync(s->h_spf, s->d_spf, ((size_t)1 << s->spf_log2) / 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if ((rc = check_launch("stitch_collect"))) return rc;
  faucet_scan_stats& o = s->sstats;
  o.n_junctions = n;
  o.nb_jcheck_kmer = st.stats[SS_JCHECK]; o.nb_no_juncs = st.stats[SS_NOJUNC]; o.nb_processed = st.stats[SS_PROCESSED];
  o.nb_skipped = st.stats[SS_SKIPPED]; o.reads_no_errors = st.stats[SS_NOERR]; o.unambiguous_reads = st.stats[SS_UNAMBIG];
  o.reads_processed = s->rec_base;
  g.tim.stitch_rounds = st.stats[SS_ROUNDS];
  g.tim.stitch_deferred = st.stats[SS_DEFERRED];
  for (int i = 0; i < 8; i++) g.tim.stitch_phase_ns[i] = st.stats[SS_T_PHASE1 + i];
  g.tim.epochs_exact = s->ep.exact_epochs; g.tim.epochs_classify = s->ep.classify_epochs;
  g.tim.exact_records = s->ep.exact_runs; g.tim.dry_records = s->ep.dry_records; g.tim.epoch_iterations = s->ep.iterations;
  g.tim.epoch_fallbacks = s->ep.fallbacks; g.tim.nonquiet_records = s->ep.nonquiet; g.tim.writer_records = st.stats[SS_WRITERS];
  return 0;
}

int faucet_session_stitch(faucet_session* s, int paired_ends, int no_cleaning, uint64_t* n_junctions_out) {
  int rc = faucet_session_stitch_begin(s, paired_ends, no_cleaning, nullptr, 0, 0, nullptr, 0, 0);
  if (rc) return rc;
  if ((rc = faucet_session_stitch_batch(s))) return rc;
  if (n_junctions_out) {
    unsigned long long n = 0;
    CU(cudaMemcpy(&n, &s->d_st->n_entries, 8, cudaMemcpyDeviceToHost));
    *n_junctions_out = n;
  }
  return 0;
}

int faucet_session_get_junctions(faucet_session* s, faucet_junction_rec** recs_out, uint64_t* n_out,
                                 faucet_scan_stats* stats) {
  if (!s->stitching) return fail(FAUCET_E_STATE, "no stitch has run in this session");
  int rc = stitch_finish(s);
  if (rc) return rc;
  const size_t nr = s->n_recs_out;
  if (recs_out) {
    *recs_out = (faucet_junction_rec*)malloc(std::max<size_t>(1, nr) * sizeof(faucet_junction_rec));
    if (!*recs_out) return fail(FAUCET_E_NOMEM, "malloc");
    if (nr) memcpy(*recs_out, s->h_recs, nr * sizeof(faucet_junction_rec));
  }
  if (n_out) *n_out = nr;
  if (stats) *stats = s->sstats;
  return 0;
}

// ---- multi-GPU (multi.cuh) -------------------------------------------------------------------------

static void* session_buffer(faucet_session* s, int what) {
  switch (what) {
    case FAUCET_BUF_INVAL: return s->d_inval;
    case FAUCET_BUF_PACKED: return s->d_packed;
    case FAUCET_BUF_FLAGS: return s->d_flags;
    case FAUCET_BUF_SEQ_START: return s->d_seq_start;
    case FAUCET_BUF_SEQ_END: return s->d_seq_end;
    case FAUCET_BUF_BLOO1_LOCAL: return s->d_b1local;
    case FAUCET_BUF_BLOOM: return s->d_bloom;
    case FAUCET_BUF_FLOW_ROWS: return s->prep_valid && !s->prep_big && s->prep_n == s->n_recs ? s->d_frows : nullptr;
    case FAUCET_BUF_FLOW_PREDS: return s->prep_valid && !s->prep_big && s->prep_n == s->n_recs ? s->d_fpreds : nullptr;
    case FAUCET_BUF_TBL_KEYS: return s->d_keys;
    case FAUCET_BUF_TBL_RECS: return s->d_recs;
    case FAUCET_BUF_TBL_PACK: return s->d_tbl_pack;
    case FAUCET_BUF_JSLOT: return s->d_jslot;
    case FAUCET_BUF_EXACT_LIST: return s->d_list;
    case FAUCET_BUF_COV_DELTA: return s->d_cov_delta;
    default: return nullptr;
  }
}

int faucet_session_prepare_multi(faucet_session* s) {
  int rc;
  if ((rc = ensure_load_buffers(s)) || (rc = ensure_scan_buffers(s))) return rc;
  if (!s->d_b1local && (rc = dmalloc(&s->d_b1local, s->tai() / 32))) return rc;
  CU(cudaMemsetAsync(s->d_b1local, 0, s->tai() / 8, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int faucet_session_export(faucet_session* s, int what, void* handle_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == FAUCET_IPC_HANDLE_BYTES, "handle size");
  void* p = session_buffer(s, what);
  if (!p && what >= FAUCET_BUF_FLOW_ROWS && what < FAUCET_BUF_COUNT) {  // optional buffers (no prepared sort, not the owner
    memset(handle_out, 0, FAUCET_IPC_HANDLE_BYTES);                     // of the table, ...): an all-zero handle says so
    return 0;
  }
  if (!p) return fail(FAUCET_E_STATE, "buffer not allocated yet (call faucet_session_prepare_multi first)");
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, p));
  memcpy(handle_out, &h, sizeof h);
  return 0;
}

int faucet_session_open_peers(faucet_session* s, int what, const void* handles, int n_ranks, int my_rank) {
  if (n_ranks < 1 || n_ranks > MAX_PEERS || my_rank < 0 || my_rank >= n_ranks) return fail(FAUCET_E_ARG, "bad rank layout");
  if (what < 0 || what >= FAUCET_BUF_COUNT) return fail(FAUCET_E_ARG, "bad buffer id");
  s->n_ranks = n_ranks; s->rank = my_rank;
  static const char zero[FAUCET_IPC_HANDLE_BYTES] = {0};
  for (int r = 0; r < n_ranks; r++) {
    if (r == my_rank) { s->peer[what][r] = session_buffer(s, what); continue; }
    const char* hb = (const char*)handles + (size_t)r * FAUCET_IPC_HANDLE_BYTES;
    // (a buffer may be exported again -- the dependency sort moves when a batch outgrows it: same handle, same mapping)
    if (s->peer[what][r] && !memcmp(hb, s->peer_handle[what][r], FAUCET_IPC_HANDLE_BYTES)) continue;
    if (s->peer[what][r]) { cudaIpcCloseMemHandle(s->peer[what][r]); s->peer[what][r] = nullptr; }
    memcpy(s->peer_handle[what][r], hb, FAUCET_IPC_HANDLE_BYTES);
    if (!memcmp(hb, zero, FAUCET_IPC_HANDLE_BYTES)) continue;  // the peer has nothing to show
    cudaIpcMemHandle_t h;
    memcpy(&h, hb, sizeof h);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer[what][r] = p;
  }
  s->peers_open = true;
  return 0;
}

int faucet_session_close_peers(faucet_session* s) {
  if (!s->peers_open) return 0;
  for (int w = 0; w < FAUCET_BUF_COUNT; w++)
    for (int r = 0; r < s->n_ranks; r++) {
      if (r != s->rank && s->peer[w][r]) cudaIpcCloseMemHandle(s->peer[w][r]);
      s->peer[w][r] = nullptr;
      memset(s->peer_handle[w][r], 0, FAUCET_IPC_HANDLE_BYTES);
    }
  s->peers_open = false;
  return 0;
}

// step 1: OR every k-mer of the parsed batch into this GPU's shard-wide bit array
int faucet_session_bloo1_local(faucet_session* s) {
  if (!s->parsed) return fail(FAUCET_E_STATE, "bloo1_local before parse");
  if (!s->d_b1local) return fail(FAUCET_E_STATE, "call faucet_session_prepare_multi first");
  LoadArgs a{};
  a.inval = s->d_inval; a.packed = s->d_packed; a.n_words = (uint32_t)((s->n + 31) / 32);
  a.tai_mask = s->tai() - 1; a.k = s->k; a.n_hash = s->n_hash;
  const int grid = g.sm_count * 8;
  switch (s->n_hash) {
    case 1: bloom_add_all_kernel<1><<<grid, LOAD_THREADS, 0, s->stream>>>(a, s->d_b1local); break;
    case 2: bloom_add_all_kernel<2><<<grid, LOAD_THREADS, 0, s->stream>>>(a, s->d_b1local); break;
    case 3: bloom_add_all_kernel<3><<<grid, LOAD_THREADS, 0, s->stream>>>(a, s->d_b1local); break;
    case 4: bloom_add_all_kernel<4><<<grid, LOAD_THREADS, 0, s->stream>>>(a, s->d_b1local); break;
    default: bloom_add_all_kernel<0><<<grid, LOAD_THREADS, 0, s->stream>>>(a, s->d_b1local); break;
  }
  s->launches++;
  return check_launch("bloom_add_all");
}

// step 2: bloo1 := OR of the shard arrays of the GPUs before this one (bloo2 := empty, stamps fresh)
int faucet_session_prefix_or(faucet_session* s) {
  if (!s->peers_open) return fail(FAUCET_E_STATE, "peers not opened");
  int rc = faucet_session_reset_filters(s);
  if (rc) return rc;
  PeerPtrs p{};
  for (int r = 0; r < s->rank; r++) p.in[r] = (const uint32_t*)s->peer[FAUCET_BUF_BLOO1_LOCAL][r];
  bloom_prefix_or_kernel<<<g.sm_count * 8, 256, 0, s->stream>>>(p, s->rank, s->d_fused, s->d_stamps, s->tai() / 32);
  s->launches++;
  return check_launch("bloom_prefix_or");
}

// step 4: in-place OR all-reduce of the per-shard bloo2 arrays (d_bloom of every GPU).  The caller
// puts a cross-process barrier before (all shards loaded and split) and after (all ranges written).
int faucet_session_or_allreduce(faucet_session* s) {
  if (!s->peers_open) return fail(FAUCET_E_STATE, "peers not opened");
  s->memo_dirty = true;  // every rank's bloo2 is rewritten (by its peers too)
  const uint64_t n_words = s->tai() / 32;
  if (n_words % 4) return fail(FAUCET_E_ARG, "Bloom filter too small for the multi-GPU reduce");
  PeerPtrs p{};
  for (int r = 0; r < s->n_ranks; r++) p.out[r] = (uint32_t*)s->peer[FAUCET_BUF_BLOOM][r];
  const uint64_t vecs = n_words / 4;
  const uint64_t v0 = vecs * s->rank / s->n_ranks, v1 = vecs * (s->rank + 1) / s->n_ranks;
  if (v1 > v0) {
    const int grid = (int)std::min<uint64_t>(g.sm_count * 8, (v1 - v0 + 255) / 256);
    bloom_or_allreduce_kernel<<<grid, 256, 0, s->stream>>>(p, s->n_ranks, v0 * 4, v1 * 4);
    s->launches++;
  }
  return check_launch("bloom_or_allreduce");
}

// pass 2 on GPU 0: pull the parsed + flagged planes of `peer_rank` over NVLink into this session
int faucet_session_import_planes(faucet_session* s, int peer_rank, size_t n_text, uint32_t n_recs, int fastq) {
  if (!s->peers_open || peer_rank < 0 || peer_rank >= s->n_ranks) return fail(FAUCET_E_STATE, "peer not opened");
  if (n_text > s->cap) return fail(FAUCET_E_ARG, "peer batch larger than this session's capacity");
  if (peer_rank != s->rank) {
    const size_t words = (n_text + 31) / 32 + 2;
    CU(cudaMemcpyAsync(s->d_inval, s->peer[FAUCET_BUF_INVAL][peer_rank], words * 4, cudaMemcpyDeviceToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_packed, s->peer[FAUCET_BUF_PACKED][peer_rank], (2 * words + 2) * 4, cudaMemcpyDeviceToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_flags, s->peer[FAUCET_BUF_FLAGS][peer_rank], (words - 2) * 32, cudaMemcpyDeviceToDevice, s->stream));  // bytes or 8 plane words per 32
    CU(cudaMemcpyAsync(s->d_seq_start, s->peer[FAUCET_BUF_SEQ_START][peer_rank], (size_t)n_recs * 4, cudaMemcpyDeviceToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_seq_end, s->peer[FAUCET_BUF_SEQ_END][peer_rank], (size_t)n_recs * 4, cudaMemcpyDeviceToDevice, s->stream));
    // the dependency sort of the shard, if its owner prepared one (faucet_session_flow_prepare)
    s->prep_valid = false;
    const void *pr = s->peer[FAUCET_BUF_FLOW_ROWS][peer_rank], *pp = s->peer[FAUCET_BUF_FLOW_PREDS][peer_rank];
    if (pr && pp && n_recs && n_recs <= g.flow_chunk) {
      int rc = flow_init(s);
      if (!rc) rc = flow_ensure(s, n_recs);
      if (rc) return rc;
      CU(cudaMemcpyAsync(s->d_frows, pr, (size_t)n_recs * ROW_WORDS * 4, cudaMemcpyDeviceToDevice, s->stream));
      CU(cudaMemcpyAsync(s->d_fpreds, pp, (size_t)n_recs * ROW_WORDS * 4, cudaMemcpyDeviceToDevice, s->stream));
      s->prep_valid = true; s->prep_n = n_recs; s->prep_big = false; s->prep_rows = s->d_frows; s->prep_preds = s->d_fpreds;
    }
  }
  s->n = n_text; s->n_recs = n_recs; s->fastq = fastq != 0; s->parsed = true;
  return 0;
}

int faucet_session_batch_info(faucet_session* s, size_t* n_text, uint32_t* n_recs) {
  if (!s->parsed) return fail(FAUCET_E_STATE, "no parsed batch");
  *n_text = s->n; *n_recs = s->n_recs;
  return 0;
}

// ---- sharded epoch (shard.cuh): the stitch across GPUs ------------------------------------------------

// the ordered executor over the records [begin, end) of the resident batch (stitch_batch = all of them, by epochs)
int faucet_session_stitch_records(faucet_session* s, uint32_t begin, uint32_t end, int advance) {
  if (!s->stitching) return fail(FAUCET_E_STATE, "stitch_begin not called");
  if (!s->parsed) return fail(FAUCET_E_STATE, "stitch_records before parse");
  if (begin > end || end > s->n_recs) return fail(FAUCET_E_ARG, "bad record range");
  if (s->lpf.enabled()) return fail(FAUCET_E_STATE, "stitch_records does not feed the long pair filter: use stitch_batch");
  int rc;
  if (begin < end && (rc = stitch_run_flow(s, nullptr, begin, end, false, true, nullptr))) return rc;
  s->ep.exact_epochs++; s->ep.exact_runs += end - begin;
  if (advance) s->rec_base += s->n_recs;
  return 0;
}

namespace {
struct ShardInfo {  // what the ranks tell each other before a sharded epoch (FAUCET_SHARD_INFO_BYTES per rank)
  uint64_t n_text, rec_base, tbl_cap;
  uint32_t n_recs, r_begin, fastq, eligible;
  StitchState st;
};
static_assert(sizeof(ShardInfo) <= FAUCET_SHARD_INFO_BYTES, "ShardInfo must fit its blob");

// what every read-only kernel of the epoch gets: own planes, own replica, own count array / counter block
void shard_args(faucet_session* s, StitchArgs& d) {
  stitch_fill_args(s, d);
  d.in_exact = s->d_in_exact; d.r_begin = s->shard.r_begin; d.r_end = s->n_recs;
  d.dirty = s->d_dirty; d.dirty_max = s->d_dirty_max;
  d.rows = s->d_rows; d.rows_base = s->shard.r_begin; d.recheck = 1; d.taint_mark = EX_COMMITTED;
  d.st = s->d_st_quiet; d.special = &s->d_st->special;
  d.cov_out = s->d_cov_delta; d.cov_stride = 4; d.cov_off = 0; d.cov_out2 = nullptr; d.st2 = nullptr;
  d.rows_ready = s->rows_ahead && s->rows_ahead_n == s->n_recs && s->rows_ahead_begin == s->shard.r_begin ? 1 : 0;
}
void shard_dry(faucet_session* s, const StitchArgs& d, int mode, uint8_t want, uint8_t after, bool live) {
  StitchArgs w = d;
  w.dry_mode = mode; w.want_flag = want; w.flag_after = after;
  if (!live) { w.keys = s->d_snap_keys; w.recs = s->d_snap_recs; w.special = &s->d_st_snap->special; }
  KTimer kt(s, KT_DRY);
  stitch_dry_kernel<<<g.sm_count * 8, DRY_THREADS, DRY_WARPS * sizeof(DryScratch), s->stream>>>(w);
  s->launches++;
}
// the ascending list of this GPU's members of the exact set -> list buffer `sel` (d_list + sel x n_recs: the ranks
// alternate between two buffers, so a rank may write its next list while a peer still reads the current one); their number
int shard_list(faucet_session* s, int sel, uint32_t* n_out) {
  const uint32_t r = s->shard.r_begin, m = s->n_recs - r;
  *n_out = 0;
  if (!m) return 0;
  const int grid = g.sm_count * 8;
  const unsigned n_blocks = (m + SCAN_CHUNK - 1) / SCAN_CHUNK;
  {
    KTimer kt(s, KT_VERIFY);
    exact_flags_kernel<<<grid, 256, 0, s->stream>>>(s->d_in_exact, r, m, s->d_eprefix);
    scan_reduce_kernel<<<n_blocks, 256, 0, s->stream>>>(s->d_eprefix, m, s->d_eprefix_sums);
    scan_sums_kernel<<<1, 1024, 0, s->stream>>>(s->d_eprefix_sums, n_blocks);
    scan_apply_kernel<<<n_blocks, 256, 0, s->stream>>>(s->d_eprefix, m, s->d_eprefix_sums);
    exact_list_kernel<<<grid, 256, 0, s->stream>>>(s->d_in_exact, r, m, s->d_eprefix, s->d_list + (size_t)sel * s->n_recs, s->d_count);
    s->launches += 5;
  }
  CU(cudaMemcpyAsync(n_out, s->d_count, 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return check_launch("shard_list");
}
}  // namespace

// Step 0 (every rank, after stitch_begin; the owner after its ordered prefix [0, r_begin)): what this rank brings to the
// epoch.  The owner first makes room in its table for what the exact set may create (the table cannot grow while
// replicas of it exist).  eligible = 0: this scan cannot be sharded (pair filters are fed): every rank must then take
// the serial path.
int faucet_session_shard_info(faucet_session* s, uint32_t r_begin, int is_owner, void* info_out) {
  if (!s->stitching) return fail(FAUCET_E_STATE, "stitch_begin not called");
  if (!s->parsed) return fail(FAUCET_E_STATE, "shard_info before parse");
  if (r_begin > s->n_recs) return fail(FAUCET_E_ARG, "bad epoch start");
  int rc;
  ShardInfo in;
  std::memset(&in, 0, sizeof in);
  CU(cudaMemcpyAsync(&s->h_st, s->d_st, sizeof(StitchState), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (is_owner) {
    if ((rc = flow_init(s))) return rc;
    // (the executor itself asks for n_entries + bound <= cap / 2 before every record; what is left above that is the room
    // for the junctions the exact set may create -- if it runs out, the epoch is called off: shard_abort)
    const unsigned long long bound = std::max<unsigned long long>(s->h_st.max_need, 1024) * (unsigned long long)s->flow_grid * STITCH_WARPS;
    while (s->h_st.n_entries + s->h_st.n_entries / 8 + bound + 65536 > s->tbl_cap / 2)
      if ((rc = stitch_grow_table(s))) return rc;
  }
  if (is_owner) {  // T0 as the other GPUs will fetch it
    if (s->tbl_pack_cap != s->tbl_cap / 2 + 2) {
      cudaFree(s->d_tbl_pack); s->d_tbl_pack = nullptr;
      s->tbl_pack_cap = s->tbl_cap / 2 + 2;
      if ((rc = dmalloc(&s->d_tbl_pack, s->tbl_pack_cap))) return rc;
    }
    if (!s->d_pack_n && (rc = dmalloc(&s->d_pack_n, 1))) return rc;
    StitchArgs a;
    stitch_fill_args(s, a);
    unsigned int n_packed = 0;
    {
      KTimer kt(s, KT_SHARD_COPY);
      CU(cudaMemsetAsync(s->d_pack_n, 0, 4, s->stream));
      shard_pack_kernel<<<g.sm_count * 8, 256, 0, s->stream>>>(a, s->d_tbl_pack, s->d_pack_n);
      s->launches++;
    }
    CU(cudaMemcpyAsync(&n_packed, s->d_pack_n, 4, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if ((rc = check_launch("shard_pack"))) return rc;
    if (n_packed != s->h_st.n_entries) return fail(FAUCET_E_CUDA, "sharded epoch: packed junction count differs from the table's entry counter");
  }
  in.n_text = s->n; in.rec_base = s->rec_base; in.tbl_cap = s->tbl_cap; in.n_recs = s->n_recs; in.r_begin = r_begin;
  in.fastq = s->fastq; in.eligible = (s->d_spf || s->lpf.enabled() || g.stitch_exec == 0) ? 0u : 1u;
  in.st = s->h_st;
  std::memset(info_out, 0, FAUCET_SHARD_INFO_BYTES);
  std::memcpy(info_out, &in, sizeof in);
  return 0;
}

// Optional, any time between parse and shard_begin: the reservation rows of this rank's records of the epoch (a pure
// function of the text), e.g. while the owner is still busy with its prefix; classify then does not compute them.
int faucet_session_shard_rows(faucet_session* s, uint32_t r_begin) {
  if (!s->parsed) return fail(FAUCET_E_STATE, "shard_rows before parse");
  if (r_begin > s->n_recs) return fail(FAUCET_E_ARG, "bad epoch start");
  const uint32_t m = s->n_recs - r_begin;
  s->rows_ahead = false;
  if (!m) return 0;
  int rc;
  if ((size_t)m > s->rows_cap) {
    cudaFree(s->d_rows); s->d_rows = nullptr;
    s->rows_cap = (size_t)m + m / 4 + 1024;
    if ((rc = dmalloc(&s->d_rows, s->rows_cap * ROW_WORDS))) return rc;
  }
  StitchArgs d;
  stitch_fill_args(s, d);
  d.rows = s->d_rows; d.rows_base = r_begin; d.r_begin = r_begin; d.r_end = s->n_recs;
  {
    KTimer kt(s, KT_FLOW_PREP);
    stitch_rows_kernel<<<g.sm_count * 8, DRY_THREADS, 0, s->stream>>>(d);
    s->launches++;
  }
  s->rows_ahead = true; s->rows_ahead_n = s->n_recs; s->rows_ahead_begin = r_begin;
  return check_launch("shard_rows");
}

// Step 1: replicate T0 (ranks other than the owner), classify this rank's records of the epoch against it, snapshot, list
// the exact set.  Needs the owner's FAUCET_BUF_TBL_KEYS / TBL_RECS / JSLOT opened after its shard_info call.
int faucet_session_shard_begin(faucet_session* s, const void* infos, int n_ranks, int my_rank, int owner, uint32_t* n_exact_out) {
  if (!s->peers_open || n_ranks != s->n_ranks || my_rank != s->rank) return fail(FAUCET_E_STATE, "peers not opened for this rank layout");
  if (owner < 0 || owner >= n_ranks) return fail(FAUCET_E_ARG, "bad owner");
  int rc;
  ShardInfo in[MAX_PEERS];
  for (int r = 0; r < n_ranks; r++) {
    std::memcpy(&in[r], (const char*)infos + (size_t)r * FAUCET_SHARD_INFO_BYTES, sizeof(ShardInfo));
    if (!in[r].eligible) return fail(FAUCET_E_STATE, "this scan cannot be sharded (pair filters / round executor)");
  }
  auto& sh = s->shard;
  sh = faucet_session::Shard();
  sh.owner = owner; sh.r_begin = in[my_rank].r_begin;
  uint64_t base = in[0].rec_base;
  for (int r = 0; r < n_ranks; r++) { sh.rec_base[r] = base; sh.n_recs[r] = in[r].n_recs; base += in[r].n_recs; }
  sh.rec_base[n_ranks] = base;
  if (base >= 0xfffffffeull) return fail(FAUCET_E_ARG, "sharded epoch: more than 2^32 - 2 records");
  if (in[my_rank].n_recs != s->n_recs) return fail(FAUCET_E_STATE, "shard_begin: the resident batch changed since shard_info");
  const unsigned long long cap = in[owner].tbl_cap;
  if (my_rank != owner) {
    if (s->tbl_cap != cap) {
      cudaFree(s->d_keys); cudaFree(s->d_recs); cudaFree(s->d_jstamps);
      s->d_keys = nullptr; s->d_recs = nullptr; s->d_jstamps = nullptr; s->tbl_cap = 0;
      if ((rc = stitch_alloc_table(s, cap))) return rc;
    }
    const void *pp = s->peer[FAUCET_BUF_TBL_PACK][owner], *pj = s->peer[FAUCET_BUF_JSLOT][owner];
    if (!pp || !pj) return fail(FAUCET_E_STATE, "the owner's table is not opened");
    const unsigned long long n_in = in[owner].st.n_entries;
    if (n_in > s->tbl_in_cap) {
      cudaFree(s->d_tbl_in); s->d_tbl_in = nullptr;
      s->tbl_in_cap = n_in + n_in / 4 + 4096;
      if ((rc = dmalloc(&s->d_tbl_in, s->tbl_in_cap))) return rc;
    }
    {
      KTimer kt(s, KT_SHARD_COPY);
      CU(cudaMemsetAsync(s->d_keys, 0xff, (cap + 1) * 8, s->stream));
      CU(cudaMemsetAsync(s->d_recs, 0, (cap + 1) * REC_WORDS * 4, s->stream));
      if (n_in) CU(cudaMemcpyAsync(s->d_tbl_in, pp, n_in * sizeof(PackedJunction), cudaMemcpyDeviceToDevice, s->stream));
      CU(cudaMemcpyAsync(s->d_jslot, pj, (((size_t)1 << g.res_log2) / 32 + 1) * 4, cudaMemcpyDeviceToDevice, s->stream));
      if (n_in) {
        StitchArgs a;
        stitch_fill_args(s, a);
        a.cap = cap;
        shard_unpack_kernel<<<(unsigned)std::min<unsigned long long>((n_in + 255) / 256, (unsigned long long)g.sm_count * 16), 256, 0, s->stream>>>(a, s->d_tbl_in, (unsigned int)n_in);
        s->launches++;
      }
    }
    s->h_st = in[owner].st;
    std::memset(s->h_st.stats, 0, sizeof s->h_st.stats);  // the counters of T0 live on the owner
    s->h_st.ext_used = 0;
    CU(cudaMemcpyAsync(s->d_st, &s->h_st, sizeof(StitchState), cudaMemcpyHostToDevice, s->stream));
  } else if (s->tbl_cap != cap) {
    return fail(FAUCET_E_STATE, "shard_begin: the owner's table changed since shard_info");
  }
  s->rec_base = sh.rec_base[my_rank];
  const uint32_t m = s->n_recs - sh.r_begin;
  if ((rc = stitch_ensure_epoch_buffers(s, std::max<uint32_t>(s->n_recs, 1)))) return rc;  // (lists: two buffers n_recs apart)
  if (s->cov_delta_cap != cap) {
    cudaFree(s->d_cov_delta); s->d_cov_delta = nullptr; s->cov_delta_cap = 0;
    if ((rc = dmalloc(&s->d_cov_delta, (cap + 1) * 4))) return rc;
    s->cov_delta_cap = cap;
  }
  if (!s->d_st_quiet && (rc = dmalloc(&s->d_st_quiet, 1))) return rc;
  CU(cudaMemsetAsync(s->d_cov_delta, 0, (cap + 1) * 16, s->stream));
  CU(cudaMemsetAsync(s->d_st_quiet, 0, sizeof(StitchState), s->stream));
  CU(cudaMemsetAsync(s->d_in_exact, 0, s->n_recs ? s->n_recs : 1, s->stream));
  sh.active = true;
  if (m) {
    StitchArgs d;
    shard_args(s, d);
    shard_dry(s, d, DRY_CLASSIFY_APPLY, 0, 0, true);
  }
  if ((rc = stitch_snapshot(s))) return rc;
  sh.snapshot = true;
  s->ep.classify_epochs++; s->ep.dry_records += m;
  sh.iter = 0;
  if ((rc = shard_list(s, 0, n_exact_out))) return rc;
  s->ep.nonquiet += *n_exact_out;
  return 0;
}

// Step 2: the whole exact set -- n_exact[r] entries of rank r's list, r = 0 .. n_ranks-1, in that order -- through the
// ordered executor on this rank's replica (restored to T0 first when iter > 0).  The lines of all members are gathered
// into one small local batch first (shard.cuh).  *need_grow_out: the table would have to grow; every rank must then call
// shard_abort and the job falls back to the serial path.
int faucet_session_shard_execute(faucet_session* s, const uint32_t* n_exact, int iter, int* need_grow_out) {
  auto& sh = s->shard;
  if (!sh.active) return fail(FAUCET_E_STATE, "no sharded epoch is open");
  *need_grow_out = 0;
  int rc;
  if (iter != sh.iter) return fail(FAUCET_E_STATE, "shard_execute: iterations out of step");
  sh.iter = iter + 1;  // (the list shard_verify builds next goes to the other buffer)
  if (iter > 0 && (rc = stitch_restore(s, 0))) return rc;
  CU(cudaMemsetAsync(s->d_dirty, 0xff, ((size_t)1 << g.res_log2) * 4, s->stream));
  CU(cudaMemsetAsync(s->d_dirty_max, 0, ((size_t)1 << g.res_log2) * 4, s->stream));
  if (g.shard_force_abort && iter == 0) { *need_grow_out = 1; return 0; }
  sh.ran = true;
  GatherArgs ga;
  std::memset(&ga, 0, sizeof ga);
  ga.n_ranks = s->n_ranks;
  uint64_t n_total = 0;
  for (int r = 0; r < s->n_ranks; r++) {
    const bool me = r == s->rank;
    ga.first[r] = (uint32_t)n_total;
    n_total += n_exact[r];
    ga.rec_base[r] = (uint32_t)sh.rec_base[r];
    ga.inval[r] = me ? s->d_inval : (const uint32_t*)s->peer[FAUCET_BUF_INVAL][r];
    ga.packed[r] = me ? s->d_packed : (const uint32_t*)s->peer[FAUCET_BUF_PACKED][r];
    ga.flags[r] = me ? s->d_flags : (const uint8_t*)s->peer[FAUCET_BUF_FLAGS][r];
    ga.seq_start[r] = me ? s->d_seq_start : (const uint32_t*)s->peer[FAUCET_BUF_SEQ_START][r];
    ga.seq_end[r] = me ? s->d_seq_end : (const uint32_t*)s->peer[FAUCET_BUF_SEQ_END][r];
    ga.list[r] = me ? s->d_list : (const uint32_t*)s->peer[FAUCET_BUF_EXACT_LIST][r];
    if (ga.list[r]) ga.list[r] += (size_t)(iter & 1) * sh.n_recs[r];
    if (n_exact[r] && (!ga.list[r] || !ga.packed[r] || !ga.inval[r] || !ga.flags[r] || !ga.seq_start[r] || !ga.seq_end[r]))
      return fail(FAUCET_E_STATE, "a peer's exact list / planes are not opened");
  }
  ga.first[s->n_ranks] = (uint32_t)n_total;
  if (n_total >= (1ull << 31)) return fail(FAUCET_E_ARG, "exact set too large");
  const uint32_t n = (uint32_t)n_total;
  ga.n = n;
  if (!n) { s->ep.iterations++; return 0; }
  if (n > s->mb_entries_cap) {
    cudaFree(s->mb_seq_start); cudaFree(s->mb_seq_end); cudaFree(s->mb_gid); cudaFree(s->mb_span); cudaFree(s->mb_span_sums);
    s->mb_seq_start = s->mb_seq_end = s->mb_gid = s->mb_span = s->mb_span_sums = nullptr;
    s->mb_entries_cap = (size_t)n + n / 2 + 4096;
    if ((rc = dmalloc(&s->mb_seq_start, s->mb_entries_cap)) || (rc = dmalloc(&s->mb_seq_end, s->mb_entries_cap)) ||
        (rc = dmalloc(&s->mb_gid, s->mb_entries_cap)) || (rc = dmalloc(&s->mb_span, s->mb_entries_cap + 1)) ||
        (rc = dmalloc(&s->mb_span_sums, s->mb_entries_cap / SCAN_CHUNK + 4)))
      return rc;
  }
  ga.span = s->mb_span; ga.o_seq_start = s->mb_seq_start; ga.o_seq_end = s->mb_seq_end; ga.o_gid = s->mb_gid;
  uint32_t tail[2] = {0, 0};
  const int grid = (int)std::min<uint64_t>((uint64_t)g.sm_count * 8, (n + 7) / 8);
  {
    KTimer kt(s, KT_SHARD_COPY);
    shard_gather_spans_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(ga);
    CU(cudaMemcpyAsync(&tail[0], s->mb_span + n - 1, 4, cudaMemcpyDeviceToHost, s->stream));
    exclusive_scan_u32(s, s->mb_span, n, s->mb_span_sums);
    CU(cudaMemcpyAsync(&tail[1], s->mb_span + n - 1, 4, cudaMemcpyDeviceToHost, s->stream));
    s->launches++;
  }
  CU(cudaStreamSynchronize(s->stream));
  if ((rc = check_launch("shard_gather_spans"))) return rc;
  const uint64_t total = (uint64_t)tail[0] + tail[1];
  if (total >= (3ull << 30)) return fail(FAUCET_E_ARG, "exact set too large for one gathered batch");
  if (total + 4096 > s->mb_pos_cap) {
    cudaFree(s->mb_inval); cudaFree(s->mb_packed); cudaFree(s->mb_flags);
    s->mb_inval = s->mb_packed = nullptr; s->mb_flags = nullptr;
    s->mb_pos_cap = total + total / 2 + 65536;
    if ((rc = dmalloc(&s->mb_inval, s->mb_pos_cap / 32 + 64)) || (rc = dmalloc(&s->mb_packed, s->mb_pos_cap / 16 + 64)) ||
        (rc = dmalloc(&s->mb_flags, s->mb_pos_cap + 64)))
      return rc;
  }
  ga.o_inval = s->mb_inval; ga.o_packed = s->mb_packed; ga.o_flags = s->mb_flags;
  {
    KTimer kt(s, KT_SHARD_COPY);
    shard_gather_copy_kernel<<<grid, 256, 0, s->stream>>>(ga, (uint32_t)total);
    s->launches++;
  }
  // the executor runs the gathered batch as if it were the resident one
  struct Planes { uint32_t *inval, *packed, *seq_start, *seq_end; uint8_t* flags; uint32_t n_recs; uint64_t rec_base; } own =
      {s->d_inval, s->d_packed, s->d_seq_start, s->d_seq_end, s->d_flags, s->n_recs, s->rec_base};
  s->d_inval = s->mb_inval; s->d_packed = s->mb_packed; s->d_flags = s->mb_flags; s->d_seq_start = s->mb_seq_start; s->d_seq_end = s->mb_seq_end;
  s->n_recs = n; s->rec_base = 0; s->cur_gid = s->mb_gid; s->prep_valid = false;
  bool need_grow = false;
  rc = stitch_run_flow(s, nullptr, 0, n, true, false, &need_grow);
  s->d_inval = own.inval; s->d_packed = own.packed; s->d_seq_start = own.seq_start; s->d_seq_end = own.seq_end; s->d_flags = own.flags;
  s->n_recs = own.n_recs; s->rec_base = own.rec_base; s->cur_gid = nullptr; s->prep_valid = false;
  if (rc) return rc;
  s->ep.exact_runs += n;
  if (need_grow) { *need_grow_out = 1; return 0; }
  s->ep.iterations++;
  return 0;
}

// Step 3: which of this rank's quiet records may have seen something else than T0 (verify), the second look of those
// with earlier writes only (recheck), the commits of the ones that join the exact set taken back (retract); the new list.
int faucet_session_shard_verify(faucet_session* s, uint32_t* n_exact_out) {
  auto& sh = s->shard;
  if (!sh.active) return fail(FAUCET_E_STATE, "no sharded epoch is open");
  if (s->n_recs > sh.r_begin) {
    StitchArgs d;
    shard_args(s, d);
    {
      KTimer kt(s, KT_VERIFY);
      stitch_verify_kernel<<<g.sm_count * 8, DRY_THREADS, 0, s->stream>>>(d);
      s->launches++;
    }
    shard_dry(s, d, DRY_RECHECK, EX_RECHECK, 0, true);
    shard_dry(s, d, DRY_RETRACT, EX_COMMITTED, EX_RETRACTED, false);
  }
  return shard_list(s, sh.iter & 1, n_exact_out);
}

// Step 4: the settled records (they saw the replica as the last exact run left it) move their commit from the walk on T0
// to the walk on that table; this rank's counters of the epoch -> stats_out (FAUCET_SHARD_STATS entries).
int faucet_session_shard_finish(faucet_session* s, uint64_t* stats_out) {
  auto& sh = s->shard;
  if (!sh.active) return fail(FAUCET_E_STATE, "no sharded epoch is open");
  if (sh.ran && s->n_recs > sh.r_begin) {
    StitchArgs d;
    shard_args(s, d);
    shard_dry(s, d, DRY_RETRACT, EX_SETTLED, EX_SETTLED, false);
    shard_dry(s, d, DRY_APPLY, EX_SETTLED, 0, true);
  }
  StitchState q;
  CU(cudaMemcpyAsync(&q, s->d_st_quiet, sizeof q, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(&s->h_st, s->d_st, sizeof(StitchState), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  int rc = check_launch("shard_finish");
  if (rc) return rc;
  if (q.stats[SS_DRY_ERROR]) return fail(FAUCET_E_CUDA, "sharded epoch: apply met a record that is not quiet (internal error)");
  static_assert(FAUCET_SHARD_STATS >= SS_COUNT, "stats blob too small");
  for (int i = 0; i < FAUCET_SHARD_STATS; i++) stats_out[i] = i < SS_COUNT ? q.stats[i] : 0;
  return 0;
}

// Step 5 (owner; after a barrier): the count arrays of all ranks into the owner's table, their counters into its
// counter block.  stats_all: n_ranks x FAUCET_SHARD_STATS, what shard_finish returned on every rank.
int faucet_session_shard_merge(faucet_session* s, const uint64_t* stats_all) {
  auto& sh = s->shard;
  if (!sh.active) return fail(FAUCET_E_STATE, "no sharded epoch is open");
  if (s->rank != sh.owner) return fail(FAUCET_E_STATE, "shard_merge runs on the owner of the table");
  StitchArgs a;
  stitch_fill_args(s, a);
  {
    KTimer kt(s, KT_SHARD_MERGE);
    for (int r = 0; r < s->n_ranks; r++) {
      const unsigned long long* pk = r == s->rank ? s->d_keys : (const unsigned long long*)s->peer[FAUCET_BUF_TBL_KEYS][r];
      const uint4* pc = r == s->rank ? (const uint4*)s->d_cov_delta : (const uint4*)s->peer[FAUCET_BUF_COV_DELTA][r];
      if (!pk || !pc) return fail(FAUCET_E_STATE, "a peer's table / count array is not opened");
      shard_merge_kernel<<<g.sm_count * 8, 256, 0, s->stream>>>(pk, pc, s->tbl_cap, a, r == s->rank ? 1 : 0);
      s->launches++;
    }
  }
  unsigned long long add[SS_COUNT] = {0};
  for (int r = 0; r < s->n_ranks; r++)
    for (int i = 0; i < SS_WALK; i++) add[i] += stats_all[(size_t)r * FAUCET_SHARD_STATS + i];
  CU(cudaMemcpyAsync(&s->h_st, s->d_st, sizeof(StitchState), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  int rc = check_launch("shard_merge");
  if (rc) return rc;
  if (s->h_st.stats[SS_DRY_ERROR]) return fail(FAUCET_E_CUDA, "sharded epoch: a peer counted coverage on a junction the owner does not hold (internal error)");
  for (int i = 0; i < SS_WALK; i++) s->h_st.stats[i] += add[i];
  CU(cudaMemcpyAsync(s->d_st->stats, s->h_st.stats, sizeof(unsigned long long) * SS_WALK, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->rec_base = sh.rec_base[s->n_ranks];  // the whole stream is in the map now
  sh.active = false;
  return 0;
}

// closes the epoch on a rank that does not merge (every rank but the owner)
int faucet_session_shard_end(faucet_session* s) {
  s->shard.active = false;
  return 0;
}

// The epoch is called off (the table would have to grow): the replica goes back to T0 and the commits of the epoch are
// forgotten; the owner then runs its records [r_begin, n_recs) and the other shards through the serial path.
int faucet_session_shard_abort(faucet_session* s) {
  auto& sh = s->shard;
  if (!sh.active) return 0;
  int rc;
  if (sh.snapshot && (rc = stitch_restore(s, 0))) return rc;
  CU(cudaMemcpyAsync(&s->h_st, s->d_st, sizeof(StitchState), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->rec_base = sh.rec_base[s->rank];
  sh.active = false;
  s->ep.fallbacks++;
  return 0;
}

}  // extern "C"

// ---- whole-pass entry points ---------------------------------------------------------------------

static void retained_clear(faucet_session* s) {  // forgets the planes, keeps the device blocks for the next pass 1
  s->retained.clear();
  for (auto& a : s->arena) a.used = 0;
  s->retained_bytes = 0;
  s->retained_valid = false;
}
static void retained_free(faucet_session* s) {
  retained_clear(s);
  for (auto& a : s->arena) cudaFree(a.p);
  s->arena.clear();
}
static uint32_t* retained_alloc(faucet_session* s, size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  for (auto& a : s->arena)
    if (a.cap - a.used >= bytes) { uint8_t* p = a.p + a.used; a.used += bytes; return reinterpret_cast<uint32_t*>(p); }
  size_t have = 0;
  for (auto& a : s->arena) have += a.cap;
  if (have + bytes > g.retain_budget) return nullptr;
  // a new block: at least what is asked, normally 1 GiB (cudaMalloc synchronises the device, so blocks are few and kept)
  const size_t cap = std::max(bytes, std::min<size_t>((size_t)1 << 30, g.retain_budget - have));
  uint8_t* p = nullptr;
  if (cudaMalloc((void**)&p, cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  s->arena.push_back({p, cap, bytes});
  return reinterpret_cast<uint32_t*>(p);
}

// keeps a copy of the parsed planes of the current batch (D2D, on the session stream)
static int retained_push(faucet_session* s) {
  const size_t n_chunks = std::max<size_t>(1, (s->n + PARSE_CHUNK - 1) / PARSE_CHUNK);
  const size_t words = n_chunks * (PARSE_CHUNK / 32) + 2;  // what parse_batch covers, guard words included
  const size_t nrec = std::max<size_t>(1, s->n_recs);
  faucet_session::Retained b{s->n, s->n_recs, s->fastq, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, false, false};
  if (!(b.inval = retained_alloc(s, words * 4)) || !(b.packed = retained_alloc(s, (2 * words + 2) * 4)) ||
      !(b.seq_start = retained_alloc(s, nrec * 4)) || !(b.seq_end = retained_alloc(s, nrec * 4))) {
    retained_clear(s);
    return 1;  // not an error: pass 2 will need the text again
  }
  cudaMemcpyAsync(b.inval, s->d_inval, words * 4, cudaMemcpyDeviceToDevice, s->stream);
  cudaMemcpyAsync(b.packed, s->d_packed, (2 * words + 2) * 4, cudaMemcpyDeviceToDevice, s->stream);
  cudaMemcpyAsync(b.seq_start, s->d_seq_start, nrec * 4, cudaMemcpyDeviceToDevice, s->stream);
  cudaMemcpyAsync(b.seq_end, s->d_seq_end, nrec * 4, cudaMemcpyDeviceToDevice, s->stream);
  size_t extra = 0;
  if (s->prep_valid && s->prep_n == s->n_recs) {  // the stitch of this batch will not have to sort
    b.big = s->prep_big;
    if (b.big) b.prep = true;
    else if ((b.rows = retained_alloc(s, nrec * ROW_WORDS * 4)) && (b.preds = retained_alloc(s, nrec * ROW_WORDS * 4))) {
      cudaMemcpyAsync(b.rows, s->prep_rows, nrec * ROW_WORDS * 4, cudaMemcpyDeviceToDevice, s->stream);
      cudaMemcpyAsync(b.preds, s->prep_preds, nrec * ROW_WORDS * 4, cudaMemcpyDeviceToDevice, s->stream);
      b.prep = true;
      extra = 2 * nrec * ROW_WORDS * 4;
    }
  }
  s->retained.push_back(b);
  s->retained_bytes += (3 * words + 2 + 2 * nrec) * 4 + extra;
  return 0;
}

static int get_session(faucet_session** out, int k, int log2_tai, int n_hash, int j, int spacer) {
  faucet_session* c = g.cached;
  if (c && c->k == k && c->log2_tai == log2_tai && c->n_hash == n_hash && c->cap >= g.batch_bytes + TAIL_MAX) {
    if (c->j != j) c->memo_dirty = true;  // the depth of the j-check is part of what the memo holds
    c->j = j; c->max_spacer = spacer;
    *out = c;
    return 0;
  }
  if (c) { faucet_session_destroy(c); g.cached = nullptr; }
  int rc = faucet_session_create(&c, k, log2_tai, n_hash, j, spacer, g.batch_bytes);
  if (rc) return rc;
  g.cached = c;
  *out = c;
  return 0;
}

// ---- where the read text comes from: host memory, or a file streamed through pinned buffers -------
struct TextSource {
  size_t n = 0;               // total bytes
  bool ends_with_newline = true;
  virtual ~TextSource() {}
  virtual void plan(const std::vector<std::pair<size_t, size_t>>&) {}
  virtual bool ready(size_t) { return true; }                       // chunk j can be acquired without blocking
  virtual const char* acquire(size_t j, size_t off, size_t len) = 0;  // host bytes of chunk j (NULL: I/O error)
  virtual void release(size_t) {}                                    // the H2D copy of chunk j has completed
};
struct MemSource : TextSource {
  const char* text;
  MemSource(const char* t, size_t n_) : text(t) { n = n_; ends_with_newline = n_ == 0 || t[n_ - 1] == '\n'; }
  const char* acquire(size_t, size_t off, size_t) override { return text + off; }
};
// A reader thread freads chunk after chunk into three pinned buffers (kept by the session), at most three
// chunks ahead of the uploads: disk I/O, host-to-device copies and kernels overlap, host memory stays
// O(batch) whatever the size of the file (BASELINE configs[4]: ~200 GB of reads).
struct FileSource : TextSource {
  faucet_session* s;
  FILE* f = nullptr;
  std::vector<std::pair<size_t, size_t>> chunks;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  size_t n_ready = 0, n_freed = 0;
  bool stop = false, io_error = false, alloc_error = false;
  FileSource(faucet_session* s_, const char* path) : s(s_) {
    f = fopen(path, "rb");
    // the reference opens the ifstream unchecked and simply sees zero reads (utils/Bloom.cpp:268-269)
    if (!f) return;
    if (fseeko(f, 0, SEEK_END) == 0) {
      const off_t sz = ftello(f);
      if (sz > 0) {
        n = (size_t)sz;
        char last = '\n';
        if (fseeko(f, sz - 1, SEEK_SET) == 0 && fread(&last, 1, 1, f) == 1) ends_with_newline = last == '\n';
      }
    }
    fseeko(f, 0, SEEK_SET);
  }
  ~FileSource() override {
    {
      std::lock_guard<std::mutex> l(mu);
      stop = true;
    }
    cv.notify_all();
    if (th.joinable()) th.join();
    if (f) fclose(f);
  }
  void plan(const std::vector<std::pair<size_t, size_t>>& c) override {
    chunks = c;
    size_t need = 1;
    for (auto& x : chunks) need = std::max(need, x.second);
    if (need > s->h_stage_cap) {
      for (int i = 0; i < 3; i++) { if (s->h_stage[i]) cudaFreeHost(s->h_stage[i]); s->h_stage[i] = nullptr; }
      s->h_stage_cap = 0;
      for (int i = 0; i < 3; i++)
        if (cudaHostAlloc((void**)&s->h_stage[i], need, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); alloc_error = true; return; }
      s->h_stage_cap = need;
    }
    th = std::thread([this] {
      for (size_t j = 0; j < chunks.size(); j++) {
        {
          std::unique_lock<std::mutex> l(mu);
          cv.wait(l, [&] { return stop || j < n_freed + 3; });
          if (stop) return;
        }
        const size_t len = chunks[j].second;
        const bool ok = len == 0 || (f && fread(s->h_stage[j % 3], 1, len, f) == len);
        {
          std::lock_guard<std::mutex> l(mu);
          if (!ok) io_error = true;
          n_ready = j + 1;
        }
        cv.notify_all();
        if (!ok) return;
      }
    });
  }
  bool ready(size_t j) override {
    std::lock_guard<std::mutex> l(mu);
    return n_ready > j || io_error || alloc_error;
  }
  const char* acquire(size_t j, size_t, size_t) override {
    if (alloc_error) return nullptr;
    std::unique_lock<std::mutex> l(mu);
    cv.wait(l, [&] { return n_ready > j || io_error; });
    return io_error ? nullptr : s->h_stage[j % 3];
  }
  void release(size_t j) override {
    {
      std::lock_guard<std::mutex> l(mu);
      n_freed = std::max(n_freed, j + 1);
    }
    cv.notify_all();
  }
};

// Feeds the text through the session in batches cut at record boundaries.  `per_batch` runs the pass on
// the parsed batch.
//
// The host text is uploaded in CHUNKS that do not depend on where records end (32 MiB ramping up to the
// batch size, three device buffers), so the copy engine runs back to back: chunk j+1 is issued before
// batch j is even parsed.  A device batch = [pad][tail][chunk]: the chunk always lands at offset TAIL_MAX
// of its buffer; the tail -- the incomplete last record of the previous batch, from the cut point the
// parse reported -- is carried over device to device right in front of it; up to 15 '#' bytes in front
// of that make the batch start 16-byte aligned.  A batch starts on a record boundary, so those bytes only
// lengthen a header line, which nothing reads.
template <class F>
static int for_each_batch(faucet_session* s, TextSource& src, bool fastq, uint64_t* total_lines, F per_batch) {
  *total_lines = 0;
  const size_t n = src.n;
  const size_t room = s->cap - TAIL_MAX - 64;
  std::vector<std::pair<size_t, size_t>> chunks;  // (offset, length) in the text
  {
    // ramp up from 32 MiB (the first copy overlaps nothing) and down again at the end (the kernels of the last batch
    // run after the last copy: keep that batch small)
    const size_t small = std::min(room, (size_t)32 << 20);
    size_t ramp = small, off = 0;
    do {
      const size_t rest = n - off;
      const size_t len = rest <= small ? rest : std::min(ramp, std::max(rest / 2, small));
      chunks.push_back({off, len});
      off += len;
      ramp = std::min(room, ramp * 2);
    } while (off < n);
  }
  src.plan(chunks);
  int rc;
  size_t issued = 0;
  auto issue = [&]() -> int {
    const int b = (int)(issued % 3);
    if (!s->d_textbufs[b]) {
      int r = dmalloc(&s->d_textbufs[b], s->cap + TEXT_PAD);
      if (r) return r;
    }
    uint8_t* dst = s->d_textbufs[b] + TAIL_MAX;
    const char* host = src.acquire(issued, chunks[issued].first, chunks[issued].second);
    if (!host && chunks[issued].second) return fail(FAUCET_E_IO, "cannot read the reads file (or pin its staging buffers)");
    if (chunks[issued].second) CU(cudaMemcpyAsync(dst, host, chunks[issued].second, cudaMemcpyHostToDevice, s->copy_stream));
    // bytes past the end must not look like bases of a previous, longer batch
    CU(cudaMemsetAsync(dst + chunks[issued].second, '\n', TEXT_PAD, s->copy_stream));
    CU(cudaEventRecord(s->ev_copied[b], s->copy_stream));
    issued++;
    return 0;
  };
  if ((rc = issue())) return rc;
  const uint8_t* tail_src = nullptr;  // device address of the previous batch's unconsumed tail
  size_t tail_len = 0;
  for (size_t j = 0; j < chunks.size(); j++) {
    // buffer (j+1) % 3 last held batch j-2, whose kernels completed before the parse of batch j-1 returned.
    // (a file source that has not read chunk j+1 yet is not waited for here: batch j goes first)
    if (issued == j + 1 && issued < chunks.size() && src.ready(issued) && (rc = issue())) return rc;
    const int b = (int)(j % 3);
    const bool final_batch = j + 1 == chunks.size();
    CU(cudaStreamWaitEvent(s->stream, s->ev_copied[b], 0));
    uint8_t* chunk0 = s->d_textbufs[b] + TAIL_MAX;
    uint8_t* start = chunk0 - tail_len;
    const size_t pad = (size_t)(reinterpret_cast<uintptr_t>(start) & 15);
    if (tail_len) CU(cudaMemcpyAsync(start, tail_src, tail_len, cudaMemcpyDeviceToDevice, s->stream));
    if (pad) CU(cudaMemsetAsync(start - pad, '#', pad, s->stream));
    s->d_text = start - pad;
    s->n = pad + tail_len + chunks[j].second;
    s->parsed = false;
    if ((rc = parse_batch(s, fastq, final_batch))) return rc;
    src.release(j);  // the parse has synchronised the stream, which waited for the copy of chunk j
    size_t consumed = s->n;
    if (!final_batch) {
      if (s->h_pctr.cut == 0) return fail(FAUCET_E_ARG, "a single record does not fit in one batch");
      consumed = (size_t)s->h_pctr.cut;
      if (s->n - consumed > TAIL_MAX - 64) return fail(FAUCET_E_ARG, "a single record does not fit in one batch");
      uint64_t lines = s->h_pctr.total_newlines;
      *total_lines += lines - (lines % (fastq ? 4 : 2));
      tail_src = s->d_text + consumed;
      tail_len = s->n - consumed;
    } else {
      *total_lines += s->h_pctr.total_newlines + ((n > 0 && !src.ends_with_newline) ? 1 : 0);
    }
    if ((rc = per_batch(consumed, final_batch))) return rc;
    if (issued == j + 1 && issued < chunks.size() && (rc = issue())) return rc;  // now it may block on the disk
  }
  return 0;
}

static int load_pass(TextSource& src, int fastq, int k, int log2_tai, int n_hash, uint8_t* bloo2_out, uint8_t* bloo1_out,
                     faucet_load_stats* stats, faucet_session** s_out) {
  if (!bloo2_out) return fail(FAUCET_E_ARG, "bloo2_out is NULL");
  faucet_session* s;
  int rc = get_session(&s, k, log2_tai, n_hash, 0, 0);
  if (rc) return rc;
  *s_out = s;
  if ((rc = ensure_load_buffers(s)) || (rc = faucet_session_reset_filters(s))) return rc;
  uint64_t total_lines = 0;
  retained_clear(s);
  bool keep = g.retain_planes;
  const bool trace = getenv("FAUCET_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  rc = for_each_batch(s, src, fastq != 0, &total_lines, [&](size_t, bool) {
    const double t0 = now();
    int r = faucet_session_load(s);
    const double t1 = now();
    // pass 2 will run on these planes: sort its dependencies now, while the copy engine is the bottleneck
    if (!r && keep && g.epoch_mode == 0) r = faucet_session_flow_prepare(s);
    const double t2 = now();
    if (!r && keep && retained_push(s)) keep = false;
    if (trace) {
      cudaStreamSynchronize(s->stream);
      fprintf(stderr, "load_pass: batch %zu bytes %u recs at %.2f ms: load issue %.2f prep %.2f push+drain %.2f\n", s->n, s->n_recs,
              t0 - t_begin, t1 - t0, t2 - t1, now() - t2);
    }
    return r;
  });
  if (trace) fprintf(stderr, "load_pass: batches done at %.2f ms\n", now() - t_begin);
  if (rc) { retained_clear(s); return rc; }
  s->retained_valid = keep && g.retain_planes;
  if ((rc = faucet_session_get_bloom(s, bloo2_out, bloo1_out))) return rc;
  if (stats && (rc = faucet_session_load_stats(s, stats, total_lines))) return rc;
  drain_events(s);
  return 0;
}

static int scan_pass(TextSource& src, int fastq, int paired_ends, int no_cleaning, int k, int j, int max_spacer_dist,
                     const uint8_t* bloo2, int log2_tai, int n_hash, uint8_t* short_pf, int spf_log2_tai, int spf_n_hash,
                     uint8_t* long_pf, int lpf_log2_tai, int lpf_n_hash, faucet_junction_rec** recs_out,
                     uint64_t* n_recs_out, faucet_scan_stats* stats, faucet_session** s_out) {
  if (!bloo2) return fail(FAUCET_E_ARG, "bloo2 is NULL");
  if (j < 0 || j > MAX_J) return fail(FAUCET_E_ARG, "j must be in [0,4]");
  faucet_session* s;
  int rc = get_session(&s, k, log2_tai, n_hash, j, max_spacer_dist);
  if (rc) return rc;
  *s_out = s;
  if ((rc = ensure_scan_buffers(s)) || (rc = faucet_session_set_bloom(s, bloo2))) return rc;
  if ((rc = faucet_session_stitch_begin(s, paired_ends, no_cleaning, short_pf, spf_log2_tai, spf_n_hash, long_pf,
                                        lpf_log2_tai, lpf_n_hash)))
    return rc;
  uint64_t total_lines = 0;
  rc = for_each_batch(s, src, fastq != 0, &total_lines, [&](size_t, bool) {
    int r = faucet_session_scan_flags(s);
    if (r) return r;
    return faucet_session_stitch_batch(s);
  });
  if (rc) return rc;
  rc = faucet_session_get_junctions(s, recs_out, n_recs_out, stats);
  drain_events(s);
  return rc;
}

extern "C" {

int faucet_gpu_load_two_filters_mem(const char* text, size_t n, int fastq, int k, int log2_tai, int n_hash,
                                    uint8_t* bloo2_out, uint8_t* bloo1_out, faucet_load_stats* stats) {
  MemSource src(text, n);
  faucet_session* s = nullptr;
  return load_pass(src, fastq, k, log2_tai, n_hash, bloo2_out, bloo1_out, stats, &s);
}

int faucet_gpu_scan_mem(const char* text, size_t n, int fastq, int paired_ends, int no_cleaning, int k, int j,
                        int max_spacer_dist, const uint8_t* bloo2, int log2_tai, int n_hash, uint8_t* short_pf,
                        int spf_log2_tai, int spf_n_hash, uint8_t* long_pf, int lpf_log2_tai, int lpf_n_hash,
                        faucet_junction_rec** recs_out, uint64_t* n_recs_out, faucet_scan_stats* stats) {
  MemSource src(text, n);
  faucet_session* s = nullptr;
  return scan_pass(src, fastq, paired_ends, no_cleaning, k, j, max_spacer_dist, bloo2, log2_tai, n_hash, short_pf,
                   spf_log2_tai, spf_n_hash, long_pf, lpf_log2_tai, lpf_n_hash, recs_out, n_recs_out, stats, &s);
}

int faucet_gpu_scan_retained(int paired_ends, int no_cleaning, int k, int j, int max_spacer_dist, const uint8_t* bloo2,
                             int log2_tai, int n_hash, uint8_t* short_pf, int spf_log2_tai, int spf_n_hash,
                             uint8_t* long_pf, int lpf_log2_tai, int lpf_n_hash, faucet_junction_rec** recs_out,
                             uint64_t* n_recs_out, faucet_scan_stats* stats) {
  if (j < 0 || j > MAX_J) return fail(FAUCET_E_ARG, "j must be in [0,4]");
  faucet_session* s = g.cached;
  if (!s || !s->retained_valid || s->k != k || s->log2_tai != log2_tai || s->n_hash != n_hash)
    return fail(FAUCET_E_STATE, "no retained planes for this geometry: run faucet_gpu_load_two_filters_mem with the "
                                "\"retain_planes\" tuning set (and within \"retain_budget\"), or use faucet_gpu_scan_mem");
  if (s->j != j) s->memo_dirty = true;  // the depth of the j-check is part of what the memo holds
  s->j = j; s->max_spacer = max_spacer_dist;
  int rc;
  if ((rc = ensure_scan_buffers(s))) return rc;
  if (bloo2) { if ((rc = faucet_session_set_bloom(s, bloo2))) return rc; }  // NULL: the device copy pass 1 left behind
  if ((rc = faucet_session_stitch_begin(s, paired_ends, no_cleaning, short_pf, spf_log2_tai, spf_n_hash, long_pf,
                                        lpf_log2_tai, lpf_n_hash)))
    return rc;
  uint32_t *inval = s->d_inval, *packed = s->d_packed, *ss = s->d_seq_start, *se = s->d_seq_end;
  const bool trace = getenv("FAUCET_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  for (auto& b : s->retained) {  // the session's plane pointers visit the retained batches in stream order
    s->d_inval = b.inval; s->d_packed = b.packed; s->d_seq_start = b.seq_start; s->d_seq_end = b.seq_end;
    s->n = b.n; s->n_recs = b.n_recs; s->fastq = b.fastq; s->parsed = true;
    s->prep_valid = b.prep; s->prep_n = b.n_recs; s->prep_big = b.big; s->prep_rows = b.rows; s->prep_preds = b.preds;
    const double t0 = now();
    if ((rc = faucet_session_scan_flags(s)) || (rc = faucet_session_stitch_batch(s))) break;
    if (trace) fprintf(stderr, "scan_retained: batch of %zu bytes, %u records: %.2f ms\n", b.n, b.n_recs, now() - t0);
  }
  if (trace) fprintf(stderr, "scan_retained: %zu batches %.2f ms\n", s->retained.size(), now() - t_begin);
  s->d_inval = inval; s->d_packed = packed; s->d_seq_start = ss; s->d_seq_end = se;
  s->parsed = false; s->prep_valid = false;
  if (rc) return rc;
  const double t_collect = now();
  rc = faucet_session_get_junctions(s, recs_out, n_recs_out, stats);
  drain_events(s);
  if (trace) fprintf(stderr, "scan_retained: collect %.2f ms\n", now() - t_collect);
  return rc;
}

// Batched Bloom walks for the contig build: see ext_masks_kernel (scan.cuh).
int faucet_gpu_query_ext_masks(const uint64_t* kmers, uint64_t n, int k, int j, const uint8_t* bloo2, int log2_tai, int n_hash,
                               uint8_t* masks_out) {
  if (j < 0 || j > MAX_J) return fail(FAUCET_E_ARG, "j must be in [0,4]");
  if (n && (!kmers || !masks_out)) return fail(FAUCET_E_ARG, "NULL array");
  faucet_session* s = nullptr;
  int rc = get_session(&s, k, log2_tai, n_hash, j, 0);
  if (rc) return rc;
  if (bloo2) { if ((rc = faucet_session_set_bloom(s, bloo2))) return rc; }  // NULL: the device copy the last pass left behind
  else if (!s->d_bloom) return fail(FAUCET_E_STATE, "no bloo2 on the device: pass the filter");
  if (!n) return 0;
  unsigned long long* d_k = nullptr;
  uint8_t* d_m = nullptr;
  if ((rc = dmalloc(&d_k, n)) || (rc = dmalloc(&d_m, n))) { cudaFree(d_k); return rc; }
  ScanArgs a{};
  a.bloom = s->d_bloom; a.wmask = (uint32_t)((s->tai() - 1) >> 5); a.k = k; a.j = j; a.n_hash = n_hash;
  cudaMemcpyAsync(d_k, kmers, n * 8, cudaMemcpyHostToDevice, s->stream);
  const int grid = (int)std::min<uint64_t>((uint64_t)g.sm_count * 8, (n + 255) / 256);
  switch (n_hash) {
    case 1: ext_masks_kernel<1><<<grid, 256, 0, s->stream>>>(a, d_k, n, d_m); break;
    case 2: ext_masks_kernel<2><<<grid, 256, 0, s->stream>>>(a, d_k, n, d_m); break;
    case 3: ext_masks_kernel<3><<<grid, 256, 0, s->stream>>>(a, d_k, n, d_m); break;
    case 4: ext_masks_kernel<4><<<grid, 256, 0, s->stream>>>(a, d_k, n, d_m); break;
    default: ext_masks_kernel<0><<<grid, 256, 0, s->stream>>>(a, d_k, n, d_m); break;
  }
  s->launches++;
  cudaMemcpyAsync(masks_out, d_m, n, cudaMemcpyDeviceToHost, s->stream);
  cudaError_t e = cudaStreamSynchronize(s->stream);
  cudaFree(d_k); cudaFree(d_m);
  if (e != cudaSuccess) return fail(FAUCET_E_CUDA, std::string("ext_masks: ") + cudaGetErrorString(e));
  return check_launch("ext_masks");
}

// The path forms stream the file: a reader thread, pinned staging buffers, O(batch) host memory.  The session
// must exist before the FileSource (it owns the staging buffers), hence the get_session up front.
int faucet_gpu_load_two_filters(const char* reads_path, int fastq, int k, int log2_tai, int n_hash,
                                uint8_t* bloo2_out, uint8_t* bloo1_out, faucet_load_stats* stats) {
  faucet_session* s = nullptr;
  int rc = get_session(&s, k, log2_tai, n_hash, 0, 0);
  if (rc) return rc;
  FileSource src(s, reads_path);
  return load_pass(src, fastq, k, log2_tai, n_hash, bloo2_out, bloo1_out, stats, &s);
}

int faucet_gpu_scan(const char* reads_path, int fastq, int paired_ends, int no_cleaning, int k, int j,
                    int max_spacer_dist, const uint8_t* bloo2, int log2_tai, int n_hash, uint8_t* short_pf,
                    int spf_log2_tai, int spf_n_hash, uint8_t* long_pf, int lpf_log2_tai, int lpf_n_hash,
                    faucet_junction_rec** recs_out, uint64_t* n_recs_out, faucet_scan_stats* stats) {
  faucet_session* s = nullptr;
  int rc = get_session(&s, k, log2_tai, n_hash, j, max_spacer_dist);
  if (rc) return rc;
  FileSource src(s, reads_path);
  return scan_pass(src, fastq, paired_ends, no_cleaning, k, j, max_spacer_dist, bloo2, log2_tai, n_hash, short_pf,
                   spf_log2_tai, spf_n_hash, long_pf, lpf_log2_tai, lpf_n_hash, recs_out, n_recs_out, stats, &s);
}

}  // extern "C"
