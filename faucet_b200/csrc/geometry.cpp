// Host-side Bloom geometry, bit-for-bit what the reference derives from the CLI flags.
// Double/float mix and integer truncations are the reference's (see each citation).
#include <cmath>
#include <utility>
#include <cstdint>

#include "../../include/faucet_gpu.h"

namespace {

// Bloom::Bloom(tai_bloom,k): hashSize = (int)log2(tai_bloom)+1, tai = 2^hashSize (utils/Bloom.cpp:165-189)
int ctor_log2_tai(uint64_t requested_bits) { return (int)std::log2((double)requested_bits) + 1; }

// set_number_of_hash_func (utils/Bloom.cpp:491-498) refuses values outside [1,10] and the
// constructor default of 4 stays
int effective_n_hash(int wanted) { return (wanted > 10 || wanted < 1) ? 4 : wanted; }

struct P1Fn {
  uint64_t est, sing;
  float fp;
  // my_func, src/Faucet.cpp:197-201.  log(fpRate) resolves to the float overload there.
  double operator()(double p1) const {
    double c = ((double)est - (1 - p1) * (double)sing) / (double)est;
    return std::log(2.0) * (double)std::log(fp) + std::log(p1) * std::log(1 - std::pow(2.0, -c));
  }
};

// brents_fun(f, lower, upper, tol, max_iter), utils/Bloom.cpp:33-124.  The first swap test there
// compares |f(a)| with |b| (not |f(b)|); that is kept because it decides the iteration path.
template <class F>
double brent_root(F f, double lower, double upper, double tol, unsigned max_iter, bool* ok) {
  double a = lower, b = upper, fa = f(a), fb = f(b), fs = 0;
  *ok = true;
  if (!(fa * fb < 0)) return -11;
  if (std::fabs(fa) < std::fabs(b)) { std::swap(a, b); std::swap(fa, fb); }
  double c = a, fc = fa, s = 0, d = 0;
  bool mflag = true;
  for (unsigned iter = 1; iter < max_iter; ++iter) {
    if (std::fabs(b - a) < tol) return s;
    if (fa != fc && fb != fc)
      s = (a * fb * fc / ((fa - fb) * (fa - fc))) + (b * fa * fc / ((fb - fa) * (fb - fc))) +
          (c * fa * fb / ((fc - fa) * (fc - fb)));
    else
      s = b - fb * (b - a) / (fb - fa);
    if (((s < (3 * a + b) * 0.25) || (s > b)) || (mflag && (std::fabs(s - b) >= (std::fabs(b - c) * 0.5))) ||
        (!mflag && (std::fabs(s - b) >= (std::fabs(c - d) * 0.5))) || (mflag && (std::fabs(b - c) < tol)) ||
        (!mflag && (std::fabs(c - d) < tol))) {
      s = (a + b) * 0.5;
      mflag = true;
    } else {
      mflag = false;
    }
    fs = f(s);
    d = c; c = b; fc = fb;
    if (fa * fs < 0) { b = s; fb = fs; } else { a = s; fa = fs; }
    if (std::fabs(fa) < std::fabs(fb)) { std::swap(a, b); std::swap(fa, fb); }
  }
  *ok = false;  // the reference runs off the end of a non-void function here
  return s;
}

}  // namespace

extern "C" int faucet_geometry_optimal(uint64_t estimated_items, float fp, int* log2_tai_out, int* n_hash_out) {
  // create_bloom_filter_optimal, utils/Bloom.cpp:229-247: int bits = -log(fpRate)/log(2)/log(2)
  int bits_per_item = (int)((double)(-std::log(fp)) / std::log(2.0) / std::log(2.0));
  if (bits_per_item < 1 || estimated_items == 0) return FAUCET_E_ARG;
  uint64_t size_bits = estimated_items * (uint64_t)bits_per_item;
  *log2_tai_out = ctor_log2_tai(size_bits);
  *n_hash_out = effective_n_hash((int)floorf((float)(0.7 * bits_per_item)));
  return 0;
}

extern "C" int faucet_geometry_2_hash(uint64_t estimated_items, float fp, int* log2_tai_out, int* n_hash_out) {
  // create_bloom_filter_2_hash, utils/Bloom.cpp:206-226
  int bits_per_item = 2 * (int)(1 / std::pow((double)fp, .5));
  if (bits_per_item < 1 || estimated_items == 0) return FAUCET_E_ARG;
  *log2_tai_out = ctor_log2_tai(estimated_items * (uint64_t)bits_per_item);
  *n_hash_out = 2;
  return 0;
}

extern "C" int faucet_geometry_from_reads(uint64_t estimated_kmers, uint64_t singletons, float fp, double* p1_out,
                                          int* log2_tai_out, int* n_hash_out) {
  // getBloomFilterFromReads, src/Faucet.cpp:204-219
  bool ok;
  P1Fn f{estimated_kmers, singletons, fp};
  double p1 = brent_root(f, (double)fp, 0.50, 0.0001, 1000, &ok);
  if (p1_out) *p1_out = p1;
  if (!ok || !(p1 > 0)) return FAUCET_E_ARG;
  return faucet_geometry_optimal(estimated_kmers, (float)p1, log2_tai_out, n_hash_out);
}

// ---- shard planning for the multi-GPU path (host only) ---------------------------------------------
// Cuts a FASTA/FASTQ text into n_shards contiguous ranges that start on record boundaries (a record =
// 2 or 4 lines, as the reference's getline loops see it: utils/Bloom.cpp:280-282,340) and are balanced
// by bytes.  offsets_out receives n_shards + 1 offsets, offsets_out[0] = 0, offsets_out[n_shards] = n.
#include <cstring>
extern "C" int faucet_host_plan_shards(const char* text, size_t n, int fastq, int n_shards, uint64_t* offsets_out) {
  if (n_shards < 1 || !offsets_out) return FAUCET_E_ARG;
  const uint64_t period = fastq ? 4 : 2;
  offsets_out[0] = 0;
  int next = 1;
  uint64_t lines = 0;
  size_t pos = 0;
  while (next < n_shards && pos < n) {
    const char* nl = (const char*)memchr(text + pos, '\n', n - pos);
    if (!nl) break;
    pos = (size_t)(nl - text) + 1;
    lines++;
    if (lines % period == 0)
      while (next < n_shards && pos >= (uint64_t)n * next / n_shards) offsets_out[next++] = pos;
  }
  while (next <= n_shards) offsets_out[next++] = n;
  return 0;
}
