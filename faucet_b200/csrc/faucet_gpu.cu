// libfaucet_gpu.so: sessions, batching and the C ABI declared in include/faucet_gpu.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/faucet_gpu.h"
#include "kmer.cuh"
#include "load.cuh"
#include "parse.cuh"
#include "scan.cuh"
#include "stitch_host.hpp"

using namespace faucet;

namespace {

struct Global {
  bool inited = false;
  int device = 0;
  int sm_count = 148;
  std::string err;
  size_t batch_bytes = (size_t)1 << 30;
  uint64_t epoch_limit = 0xfffffffeull;
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
enum { KT_PARSE = 0, KT_LOAD_A, KT_LOAD_B, KT_SCAN, KT_STITCH, KT_COUNT };

}  // namespace

struct faucet_session {
  int k = 0, log2_tai = 0, n_hash = 0, j = 0, max_spacer = 0;
  size_t cap = 0;  // bytes of text one batch may hold
  cudaStream_t stream = nullptr;
  // batch text: d_textbuf has TAIL_MAX bytes of head-room so that a carried tail can be prepended
  uint8_t* d_textbuf = nullptr;
  uint8_t* d_text = nullptr;  // start of the current batch inside d_textbuf (16-byte aligned)
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
  uint8_t* d_flags = nullptr;
  uint8_t* h_flags = nullptr;   // pinned
  uint8_t* h_text = nullptr;    // pinned copy of the batch text for the host stitch (device-resident runs)
  HostStitch* stitch = nullptr;
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
  if (!s->h_flags) CU(cudaHostAlloc((void**)&s->h_flags, s->cap + TEXT_PAD, cudaHostAllocDefault));
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
void faucet_gpu_free(void* p) { free(p); }

// ---- sessions ----------------------------------------------------------------------------------

int faucet_session_create(faucet_session** out, int k, int log2_tai, int n_hash, int j, int max_spacer_dist,
                          size_t max_text_bytes) {
  if (!g.inited) { int rc = faucet_gpu_init(0); if (rc) return rc; }
  if (k < 2 || k > 32) return fail(FAUCET_E_ARG, "k must be in [2,32]");
  if (log2_tai < 6 || log2_tai > 40) return fail(FAUCET_E_ARG, "log2_tai must be in [6,40]");
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
  if ((rc = dmalloc(&s->d_textbuf, s->cap + TEXT_PAD)) || (rc = dmalloc(&s->d_inval, words)) ||
      (rc = dmalloc(&s->d_packed, 2 * words)) || (rc = dmalloc(&s->d_skipA, words)) ||
      (rc = dmalloc(&s->d_chunk, s->cap / PARSE_CHUNK + 16)) || (rc = dmalloc(&s->d_pctr, 1)) ||
      (rc = dmalloc(&s->d_lctr, 1)) || (rc = dmalloc(&s->d_complex, s->complex_cap))) {
    faucet_session_destroy(s);
    return rc;
  }
  cudaMemsetAsync(s->d_inval, 0xff, words * 4, s->stream);
  cudaMemsetAsync(s->d_packed, 0, 2 * words * 4, s->stream);
  cudaMemsetAsync(s->d_lctr, 0, sizeof(LoadCounters), s->stream);
  cudaMemsetAsync(s->d_textbuf, '\n', s->cap + TEXT_PAD, s->stream);
  s->d_text = s->d_textbuf + TAIL_MAX;
  *out = s;
  return 0;
}

void faucet_session_destroy(faucet_session* s) {
  if (!s) return;
  if (s->stream) cudaStreamSynchronize(s->stream);
  drain_events(s);
  cudaFree(s->d_textbuf); cudaFree(s->d_inval); cudaFree(s->d_packed); cudaFree(s->d_skipA);
  cudaFree(s->d_pend); cudaFree(s->d_chunk); cudaFree(s->d_pctr); cudaFree(s->d_lctr);
  cudaFree(s->d_complex); cudaFree(s->d_fused); cudaFree(s->d_stamps); cudaFree(s->d_bloom);
  cudaFree(s->d_bloom1); cudaFree(s->d_flags);
  if (s->h_flags) cudaFreeHost(s->h_flags);
  if (s->h_text) cudaFreeHost(s->h_text);
  if (s->t0) cudaEventDestroy(s->t0);
  if (s->t1) cudaEventDestroy(s->t1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s->stitch;
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

// text placement: the batch starts `tail` bytes before d_textbuf+TAIL_MAX; the start is aligned down
// to 16 bytes and the gap is filled with '#' (it lands on a header line, where it is inert).
static int place_text(faucet_session* s, const void* text, size_t n, size_t tail, cudaMemcpyKind kind) {
  if (n + tail > s->cap - TAIL_MAX + tail || tail > TAIL_MAX) return fail(FAUCET_E_ARG, "text larger than the session batch capacity");
  uint8_t* start = s->d_textbuf + TAIL_MAX - tail;
  uint8_t* aligned = (uint8_t*)((uintptr_t)start & ~(uintptr_t)15);
  if (aligned != start) CU(cudaMemsetAsync(aligned, '#', start - aligned, s->stream));
  if (n) CU(cudaMemcpyAsync(s->d_textbuf + TAIL_MAX, text, n, kind, s->stream));
  s->d_text = aligned;
  s->n = (start - aligned) + tail + n;
  // bytes past the end must not look like bases of a previous, longer batch
  CU(cudaMemsetAsync(s->d_text + s->n, '\n', TEXT_PAD, s->stream));
  s->parsed = false;
  return 0;
}

int faucet_session_set_text(faucet_session* s, const void* text, size_t n, int src_is_device) {
  return place_text(s, text, n, 0, src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice);
}

int faucet_session_reset_filters(faucet_session* s) {
  if (!s->d_fused) return ensure_load_buffers(s);
  CU(cudaMemsetAsync(s->d_fused, 0, s->tai() / 32 * 8, s->stream));
  CU(cudaMemsetAsync(s->d_stamps, 0xff, s->tai() * 4, s->stream));
  CU(cudaMemsetAsync(s->d_lctr, 0, sizeof(LoadCounters), s->stream));
  s->stamp_base = 1;
  return 0;
}

static int parse_batch(faucet_session* s, bool fastq, bool final_batch) {
  s->fastq = fastq;
  s->final_batch = final_batch;
  uint32_t n_chunks = (uint32_t)((s->n + PARSE_CHUNK - 1) / PARSE_CHUNK);
  if (n_chunks == 0) n_chunks = 1;
  size_t words = (size_t)n_chunks * (PARSE_CHUNK / 32);
  CU(cudaMemsetAsync(s->d_pctr, 0, sizeof(ParseCounters), s->stream));
  CU(cudaMemsetAsync(s->d_skipA, 0, (words + 2) * 4, s->stream));
  {
    KTimer kt(s, KT_PARSE);
    parse_count_kernel<<<n_chunks, PARSE_THREADS, 0, s->stream>>>(s->d_text, s->n, s->d_chunk);
    parse_scan_kernel<<<1, 1024, 0, s->stream>>>(s->d_chunk, n_chunks, s->d_pctr);
    ParseArgs a;
    a.text = s->d_text; a.n = s->n; a.inval = s->d_inval; a.packed = s->d_packed; a.skipA = s->d_skipA;
    a.chunk_prefix = s->d_chunk; a.ctr = s->d_pctr; a.complex_list = s->d_complex; a.complex_cap = s->complex_cap;
    a.period_mask = fastq ? 3 : 1; a.final_batch = final_batch ? 1 : 0; a.k = s->k;
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
  const int grid = g.sm_count * 8;
  {
    KTimer kt(s, KT_LOAD_A);
    DISPATCH_NH(load_A_kernel, s->n_hash, grid, LOAD_THREADS, s->stream, a);
    s->launches++;
  }
  if (a.n_complex) {
    load_complex_kernel<0><<<std::min<uint32_t>(grid, (a.n_complex + 7) / 8), LOAD_THREADS, 0, s->stream>>>(a);
    s->launches++;
  }
  {
    KTimer kt(s, KT_LOAD_B);
    DISPATCH_NH(load_B_kernel, s->n_hash, grid, LOAD_THREADS, s->stream, a);
    s->launches++;
  }
  if (a.n_complex) {
    load_complex_kernel<1><<<std::min<uint32_t>(grid, (a.n_complex + 7) / 8), LOAD_THREADS, 0, s->stream>>>(a);
    s->launches++;
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
  s->launches++;
  if (bloo2_out) CU(cudaMemcpyAsync(bloo2_out, s->d_bloom, s->tai() / 8, cudaMemcpyDeviceToHost, s->stream));
  if (bloo1_out) CU(cudaMemcpyAsync(bloo1_out, s->d_bloom1, s->tai() / 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return check_launch("bloom_split");
}

int faucet_session_set_bloom(faucet_session* s, const uint8_t* bloo2) {
  int rc;
  if (!s->d_bloom && (rc = dmalloc(&s->d_bloom, s->tai() / 32))) return rc;
  CU(cudaMemcpyAsync(s->d_bloom, bloo2, s->tai() / 8, cudaMemcpyHostToDevice, s->stream));
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

int faucet_session_scan_flags(faucet_session* s) {
  if (!s->parsed) return fail(FAUCET_E_STATE, "faucet_session_scan_flags before faucet_session_parse");
  int rc = ensure_scan_buffers(s);
  if (rc) return rc;
  ScanArgs a;
  a.inval = s->d_inval; a.packed = s->d_packed; a.n_words = (uint32_t)((s->n + 31) / 32);
  a.bloom = s->d_bloom; a.tai_mask = s->tai() - 1; a.k = s->k; a.j = s->j; a.n_hash = s->n_hash; a.flags = s->d_flags;
  const int grid = g.sm_count * 8;
  {
    KTimer kt(s, KT_SCAN);
    DISPATCH_NH(scan_flags_kernel, s->n_hash, grid, SCAN_THREADS, s->stream, a);
    s->launches++;
  }
  return check_launch("scan_flags");
}

int faucet_session_stitch_begin(faucet_session* s, int paired_ends, int no_cleaning, uint8_t* short_pf,
                                int spf_log2_tai, int spf_n_hash, uint8_t* long_pf, int lpf_log2_tai,
                                int lpf_n_hash) {
  delete s->stitch;
  s->stitch = new HostStitch(s->k, s->j, s->max_spacer, paired_ends != 0, no_cleaning != 0);
  s->stitch->set_pair_filters(short_pf, spf_log2_tai, spf_n_hash, long_pf, lpf_log2_tai, lpf_n_hash);
  return 0;
}

// host_text: the same bytes as the device batch (NULL => copied back from the device)
int faucet_session_stitch_batch(faucet_session* s, const uint8_t* host_text, size_t valid_bytes) {
  if (!s->stitch) return fail(FAUCET_E_STATE, "stitch_begin not called");
  CU(cudaMemcpyAsync(s->h_flags, s->d_flags, s->n, cudaMemcpyDeviceToHost, s->stream));
  if (!host_text) {
    if (!s->h_text) CU(cudaHostAlloc((void**)&s->h_text, s->cap + TEXT_PAD, cudaHostAllocDefault));
    CU(cudaMemcpyAsync(s->h_text, s->d_text, s->n, cudaMemcpyDeviceToHost, s->stream));
    host_text = s->h_text;
  }
  CU(cudaStreamSynchronize(s->stream));
  s->stitch->process(host_text, valid_bytes, s->h_flags, s->fastq);
  return 0;
}

int faucet_session_stitch(faucet_session* s, int paired_ends, int no_cleaning, uint64_t* n_junctions_out) {
  int rc = faucet_session_stitch_begin(s, paired_ends, no_cleaning, nullptr, 0, 0, nullptr, 0, 0);
  if (rc) return rc;
  rc = faucet_session_stitch_batch(s, nullptr, s->n);
  if (rc) return rc;
  if (n_junctions_out) *n_junctions_out = s->stitch->records().size();
  return 0;
}

int faucet_session_get_junctions(faucet_session* s, faucet_junction_rec** recs_out, uint64_t* n_out,
                                 faucet_scan_stats* stats) {
  if (!s->stitch) return fail(FAUCET_E_STATE, "no stitch has run in this session");
  auto& r = s->stitch->records();
  if (recs_out) {
    *recs_out = (faucet_junction_rec*)malloc(std::max<size_t>(1, r.size()) * sizeof(faucet_junction_rec));
    if (!*recs_out) return fail(FAUCET_E_NOMEM, "malloc");
    if (!r.empty()) memcpy(*recs_out, r.data(), r.size() * sizeof(faucet_junction_rec));
  }
  if (n_out) *n_out = r.size();
  if (stats) *stats = s->stitch->stats();
  return 0;
}

}  // extern "C"

// ---- whole-pass entry points ---------------------------------------------------------------------

static int get_session(faucet_session** out, int k, int log2_tai, int n_hash, int j, int spacer) {
  faucet_session* c = g.cached;
  if (c && c->k == k && c->log2_tai == log2_tai && c->n_hash == n_hash && c->cap >= g.batch_bytes + TAIL_MAX) {
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

// Feeds `text` through the session in batches cut at record boundaries.  `per_batch` runs the pass
// on the parsed batch; consumed = bytes of the batch that belonged to complete records.
template <class F>
static int for_each_batch(faucet_session* s, const char* text, size_t n, bool fastq, uint64_t* total_lines, F per_batch) {
  size_t off = 0;
  *total_lines = 0;
  const size_t room = s->cap - TAIL_MAX;
  do {
    size_t len = std::min(room, n - off);
    bool final_batch = off + len == n;
    int rc = place_text(s, text + off, len, 0, cudaMemcpyHostToDevice);
    if (rc) return rc;
    if ((rc = parse_batch(s, fastq, final_batch))) return rc;
    size_t consumed = len;
    const size_t lead = s->n - len;  // '#' alignment bytes in front (0 here: no tail carried on the device)
    if (!final_batch) {
      if (s->h_pctr.cut == 0) return fail(FAUCET_E_ARG, "a single record does not fit in one batch");
      consumed = (size_t)s->h_pctr.cut - lead;
      uint64_t lines = s->h_pctr.total_newlines;
      *total_lines += lines - (lines % (fastq ? 4 : 2));
    } else {
      *total_lines += s->h_pctr.total_newlines + ((len > 0 && text[off + len - 1] != '\n') ? 1 : 0);
    }
    if ((rc = per_batch((const uint8_t*)text + off, lead, consumed, final_batch))) return rc;
    off += consumed;
  } while (off < n);
  return 0;
}

extern "C" {

int faucet_gpu_load_two_filters_mem(const char* text, size_t n, int fastq, int k, int log2_tai, int n_hash,
                                    uint8_t* bloo2_out, uint8_t* bloo1_out, faucet_load_stats* stats) {
  if (!bloo2_out) return fail(FAUCET_E_ARG, "bloo2_out is NULL");
  faucet_session* s;
  int rc = get_session(&s, k, log2_tai, n_hash, 0, 0);
  if (rc) return rc;
  if ((rc = ensure_load_buffers(s)) || (rc = faucet_session_reset_filters(s))) return rc;
  uint64_t total_lines = 0;
  rc = for_each_batch(s, text, n, fastq != 0, &total_lines,
                      [&](const uint8_t*, size_t, size_t, bool) { return faucet_session_load(s); });
  if (rc) return rc;
  if ((rc = faucet_session_get_bloom(s, bloo2_out, bloo1_out))) return rc;
  if (stats && (rc = faucet_session_load_stats(s, stats, total_lines))) return rc;
  drain_events(s);
  return 0;
}

int faucet_gpu_scan_mem(const char* text, size_t n, int fastq, int paired_ends, int no_cleaning, int k, int j,
                        int max_spacer_dist, const uint8_t* bloo2, int log2_tai, int n_hash, uint8_t* short_pf,
                        int spf_log2_tai, int spf_n_hash, uint8_t* long_pf, int lpf_log2_tai, int lpf_n_hash,
                        faucet_junction_rec** recs_out, uint64_t* n_recs_out, faucet_scan_stats* stats) {
  if (!bloo2) return fail(FAUCET_E_ARG, "bloo2 is NULL");
  if (j < 0 || j > MAX_J) return fail(FAUCET_E_ARG, "j must be in [0,4]");
  faucet_session* s;
  int rc = get_session(&s, k, log2_tai, n_hash, j, max_spacer_dist);
  if (rc) return rc;
  if ((rc = ensure_scan_buffers(s)) || (rc = faucet_session_set_bloom(s, bloo2))) return rc;
  if ((rc = faucet_session_stitch_begin(s, paired_ends, no_cleaning, short_pf, spf_log2_tai, spf_n_hash, long_pf,
                                        lpf_log2_tai, lpf_n_hash)))
    return rc;
  uint64_t total_lines = 0;
  rc = for_each_batch(s, text, n, fastq != 0, &total_lines,
                      [&](const uint8_t* host, size_t lead, size_t consumed, bool) {
                        int r = faucet_session_scan_flags(s);
                        if (r) return r;
                        // flags are indexed by device offsets = host offsets + lead
                        CU(cudaMemcpyAsync(s->h_flags, s->d_flags + lead, consumed, cudaMemcpyDeviceToHost, s->stream));
                        CU(cudaStreamSynchronize(s->stream));
                        s->stitch->process(host, consumed, s->h_flags, s->fastq);
                        return 0;
                      });
  if (rc) return rc;
  rc = faucet_session_get_junctions(s, recs_out, n_recs_out, stats);
  drain_events(s);
  return rc;
}

static int read_whole_file(const char* path, std::vector<char>* buf) {
  FILE* f = fopen(path, "rb");
  // the reference opens the ifstream unchecked and simply sees zero reads (utils/Bloom.cpp:268-269)
  if (!f) { buf->clear(); return 0; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (sz < 0) { fclose(f); return fail(FAUCET_E_IO, std::string("cannot size ") + path); }
  buf->resize((size_t)sz);
  size_t got = sz ? fread(buf->data(), 1, (size_t)sz, f) : 0;
  fclose(f);
  if (got != (size_t)sz) return fail(FAUCET_E_IO, std::string("short read on ") + path);
  return 0;
}

int faucet_gpu_load_two_filters(const char* reads_path, int fastq, int k, int log2_tai, int n_hash,
                                uint8_t* bloo2_out, uint8_t* bloo1_out, faucet_load_stats* stats) {
  std::vector<char> buf;
  int rc = read_whole_file(reads_path, &buf);
  if (rc) return rc;
  return faucet_gpu_load_two_filters_mem(buf.data(), buf.size(), fastq, k, log2_tai, n_hash, bloo2_out, bloo1_out, stats);
}

int faucet_gpu_scan(const char* reads_path, int fastq, int paired_ends, int no_cleaning, int k, int j,
                    int max_spacer_dist, const uint8_t* bloo2, int log2_tai, int n_hash, uint8_t* short_pf,
                    int spf_log2_tai, int spf_n_hash, uint8_t* long_pf, int lpf_log2_tai, int lpf_n_hash,
                    faucet_junction_rec** recs_out, uint64_t* n_recs_out, faucet_scan_stats* stats) {
  std::vector<char> buf;
  int rc = read_whole_file(reads_path, &buf);
  if (rc) return rc;
  return faucet_gpu_scan_mem(buf.data(), buf.size(), fastq, paired_ends, no_cleaning, k, j, max_spacer_dist, bloo2,
                             log2_tai, n_hash, short_pf, spf_log2_tai, spf_n_hash, long_pf, lpf_log2_tai, lpf_n_hash,
                             recs_out, n_recs_out, stats);
}

}  // extern "C"
