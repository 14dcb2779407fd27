// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// extern "C" driver over the UNMODIFIED reference objects (compiled by oracle/Makefile straight from
// /root/reference into oracle/_ref/).  It lets the Python tests call the reference's own
// load_two_filters (utils/Bloom.cpp:267), ReadScanner::scanReads (src/ReadScanner.cpp:284),
// Bloom::oldHash (utils/Bloom.h:134), revcomp (utils/Kmer.cpp:238), brents_fun (utils/Bloom.cpp:33)
// and JunctionMap::writeToFile (utils/JunctionMap.cpp:579) so that oracle/faucet_oracle.c (our C
// restatement) and the CUDA path can be pinned against the real thing.
//
// This file is ours; it contains no reference code, only calls into it.
// every std header the reference pulls in is included first, so the access-specifier trick
// below only ever sees reference class bodies
#include <algorithm>
#include <cmath>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <queue>
#include <set>
#include <sstream>
#include <stack>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <inttypes.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>
#define private public      // ReadScanner counters are private (src/ReadScanner.h:46-48); read-only peek
#define protected public
#include "src/ReadScanner.h"
#undef private
#undef protected
#include <cstring>
#include <cstdio>
#include <string>
#include <set>
#include <unistd.h>
#include <fcntl.h>

namespace {
struct QuietStdout {  // the reference prints progress to stdout; silence it inside shim calls
  int saved;
  QuietStdout() { fflush(stdout); saved = dup(1); int nul = open("/dev/null", O_WRONLY); dup2(nul, 1); close(nul); }
  ~QuietStdout() { fflush(stdout); dup2(saved, 1); close(saved); }
};
Bloom* make_bloom(int log2_tai, int n_hash) {
  // Bloom::Bloom(tai_bloom,k) rounds up: hashSize=(int)log2(tai_bloom)+1  (utils/Bloom.cpp:165-189)
  Bloom* b = new Bloom((uint64_t)1 << (log2_tai - 1), sizeKmer);
  b->set_number_of_hash_func(n_hash);
  return b;
}
}  // namespace

struct ref_junction_rec {
  uint64_t kmer;
  uint8_t dist[5];
  uint8_t cov[4];
  uint8_t linked[5];
  uint8_t pad[2];
};

struct ref_scan_stats {
  uint64_t n_junctions, nb_jcheck_kmer, nb_no_juncs, nb_processed, nb_skipped, reads_no_errors,
      reads_processed, unambiguous_reads;
};

static double g_est, g_sing, g_fp;
static uint64_t g_est_u, g_sing_u;
static float g_fp_f;
// restatement of my_func (src/Faucet.cpp:197-201) over the same typed globals
static double shim_my_func(double p1) {
  double c = (g_est_u - (1 - p1) * g_sing_u) / g_est_u;
  return std::log(2) * std::log(g_fp_f) + std::log(p1) * std::log(1 - std::pow(2, -c));  // std::log(float): Faucet.o calls logf here
}

extern "C" {

void ref_set_k(int k) { setSizeKmer(k); }
int ref_get_k() { return sizeKmer; }
uint64_t ref_revcomp(uint64_t x) { return revcomp(x); }
uint64_t ref_get_canon(uint64_t x) { return get_canon(x); }
int ref_nt2int(char c) { return NT2int(c); }

uint64_t ref_old_hash(int log2_tai, uint64_t key, int i) {
  static Bloom* cache[64];  // one Bloom per geometry, kept (allocating 2^33 bits per call is slow)
  if (!cache[log2_tai]) cache[log2_tai] = make_bloom(log2_tai, 4);
  return cache[log2_tai]->oldHash(key, i);
}

// iteration order of a libstdc++ unordered_map<kmer_type, Junction> after inserting `keys` in the given order
// (what JunctionMap::writeToFile walks, utils/JunctionMap.cpp:579-596)
void ref_iteration_order(const uint64_t* keys, uint64_t n, uint64_t* out) {
  std::unordered_map<kmer_type, Junction> m;
  for (uint64_t i = 0; i < n; i++) m[keys[i]] = Junction();
  uint64_t j = 0;
  for (auto it = m.begin(); it != m.end(); ++it) out[j++] = it->first;
}

// JunctionMap::getValidJExtension (utils/JunctionMap.cpp:474-490) for each oriented k-mer, over the given bloo2
void ref_valid_j_extension(const uint64_t* kmers, uint64_t n, int j, const uint8_t* bloo2, int log2_tai, int n_hash, int* out) {
  Bloom* b = make_bloom(log2_tai, n_hash);
  memcpy(b->blooma, bloo2, ((size_t)1 << log2_tai) / 8);
  JChecker* jc = new JChecker(j, b);
  JunctionMap* jm = new JunctionMap(b, jc, 100);
  for (uint64_t i = 0; i < n; i++) out[i] = jm->getValidJExtension(DoubleKmer(kmers[i]));
  delete jm; delete jc; delete b;
}

uint64_t ref_seed(int i) {
  Bloom* b = make_bloom(10, 4);
  uint64_t s = b->seed_tab[i];
  delete b;
  return s;
}

// p1 as Faucet.cpp:208 computes it (reference brents_fun, restated my_func)
double ref_brent_p1(uint64_t est, uint64_t singletons, float fp) {
  QuietStdout q;
  g_est_u = est; g_sing_u = singletons; g_fp_f = fp;
  std::function<double(double)> f = shim_my_func;
  return brents_fun(f, fp, 0.50, 0.0001, 1000);
}

// geometry as create_bloom_filter_optimal derives it (utils/Bloom.cpp:229-247)
void ref_geometry_optimal(uint64_t est, float fp, int* log2_tai, int* n_hash, uint64_t* nchar) {
  QuietStdout q;
  Bloom dummy((uint64_t)8, sizeKmer);
  Bloom* b = dummy.create_bloom_filter_optimal(est, fp);
  *log2_tai = b->getHashSize();
  *n_hash = b->getNumHash();
  *nchar = b->tai / 8;
  delete b;
}
void ref_geometry_2_hash(uint64_t est, float fp, int* log2_tai, int* n_hash, uint64_t* nchar) {
  QuietStdout q;
  Bloom dummy((uint64_t)8, sizeKmer);
  Bloom* b = dummy.create_bloom_filter_2_hash(est, fp);
  *log2_tai = b->getHashSize();
  *n_hash = b->getNumHash();
  *nchar = b->tai / 8;
  delete b;
}

// the reference's pass 1; both bit arrays are copied out (nbytes each = 2^log2_tai/8)
int ref_load_two_filters(const char* path, int fastq, int log2_tai, int n_hash, uint8_t* bloo1_out,
                         uint8_t* bloo2_out) {
  QuietStdout q;
  Bloom* b1 = make_bloom(log2_tai, n_hash);
  Bloom* b2 = make_bloom(log2_tai, n_hash);
  load_two_filters(b1, b2, std::string(path), fastq != 0, false);
  uint64_t nb = b1->tai / 8;
  if (bloo1_out) memcpy(bloo1_out, b1->blooma, nb);
  if (bloo2_out) memcpy(bloo2_out, b2->blooma, nb);
  delete b1;
  delete b2;
  return 0;
}

// the reference's pass 2.  Junction records come back in std::unordered_map ITERATION order
// (what writeToFile emits).  Pair filters are in/out byte arrays (may be NULL with no_cleaning).
// If junctions_path != NULL the reference's own writeToFile is called on it as well.
int ref_scan(const char* path, int fastq, int paired_ends, int no_cleaning, int j, int max_spacer_dist,
             const uint8_t* bloo2, int log2_tai, int n_hash, uint8_t* short_pf, int spf_log2_tai,
             int spf_n_hash, uint8_t* long_pf, int lpf_log2_tai, int lpf_n_hash,
             ref_junction_rec* recs_out, uint64_t recs_cap, ref_scan_stats* stats,
             const char* junctions_path) {
  QuietStdout q;
  Bloom* bloom = make_bloom(log2_tai, n_hash);
  memcpy(bloom->blooma, bloo2, bloom->tai / 8);
  Bloom* spf = nullptr;
  Bloom* lpf = nullptr;
  if (short_pf) { spf = make_bloom(spf_log2_tai, spf_n_hash); memcpy(spf->blooma, short_pf, spf->tai / 8); }
  if (long_pf) { lpf = make_bloom(lpf_log2_tai, lpf_n_hash); memcpy(lpf->blooma, long_pf, lpf->tai / 8); }
  JChecker* jc = new JChecker(j, bloom);
  JunctionMap* jm = new JunctionMap(bloom, jc, 0);
  ReadScanner* sc = new ReadScanner(jm, std::string(path), bloom, spf, lpf, jc, max_spacer_dist);
  sc->scanReads(fastq != 0, paired_ends != 0, no_cleaning != 0);
  stats->n_junctions = jm->getNumJunctions();
  stats->nb_jcheck_kmer = sc->NbJCheckKmer;
  stats->nb_no_juncs = sc->NbNoJuncs;
  stats->nb_processed = sc->NbProcessed;
  stats->nb_skipped = sc->NbSkipped;
  stats->reads_no_errors = sc->readsNoErrors;
  stats->reads_processed = sc->readsProcessed;
  stats->unambiguous_reads = sc->unambiguousReads;
  uint64_t n = 0;
  for (auto it = jm->junctionMap.begin(); it != jm->junctionMap.end(); ++it, ++n) {
    if (n >= recs_cap) continue;
    ref_junction_rec& r = recs_out[n];
    memset(&r, 0, sizeof(r));
    r.kmer = it->first;
    for (int i = 0; i < 5; i++) { r.dist[i] = it->second.dist[i]; r.linked[i] = it->second.linked[i]; }
    for (int i = 0; i < 4; i++) r.cov[i] = (uint8_t)it->second.getCoverage(i);
  }
  if (junctions_path) jm->writeToFile(std::string(junctions_path));
  if (short_pf) memcpy(short_pf, spf->blooma, spf->tai / 8);
  if (long_pf) memcpy(long_pf, lpf->blooma, lpf->tai / 8);
  delete sc; delete jm; delete jc; delete bloom; delete spf; delete lpf;
  return 0;
}

// ReadscanTest-style run (src/newTests/ReadscanTest.cpp:60-100): fake Bloom = exact set of the
// canonical forms of `kmers`; reads are fed through scanInputRead one by one.
int ref_scan_fake(const char* const* reads, int n_reads, const uint64_t* valid_canon, int n_valid, int j,
                  int max_spacer_dist, ref_junction_rec* recs_out, uint64_t recs_cap,
                  ref_scan_stats* stats) {
  QuietStdout q;
  Bloom* bloom = new Bloom((uint64_t)10000, sizeKmer);
  bloom->fakify();
  std::set<bloom_elem> valids(valid_canon, valid_canon + n_valid);
  bloom->addFakeKmers(valids);
  Bloom* spf = new Bloom((uint64_t)10000, sizeKmer);
  Bloom* lpf = new Bloom((uint64_t)10000, sizeKmer);
  JChecker* jc = new JChecker(j, bloom);
  JunctionMap* jm = new JunctionMap(bloom, jc, 30);
  ReadScanner* sc = new ReadScanner(jm, "mockFileName", bloom, spf, lpf, jc, max_spacer_dist);
  sc->NbJCheckKmer = sc->NbNoJuncs = sc->NbSkipped = sc->NbProcessed = sc->readsNoErrors = 0;
  sc->readsProcessed = sc->unambiguousReads = 0;
  for (int i = 0; i < n_reads; i++) sc->scanInputRead(std::string(reads[i]), true);
  stats->n_junctions = jm->getNumJunctions();
  stats->nb_jcheck_kmer = sc->NbJCheckKmer;
  stats->nb_no_juncs = sc->NbNoJuncs;
  stats->nb_processed = sc->NbProcessed;
  stats->nb_skipped = sc->NbSkipped;
  stats->reads_no_errors = sc->readsNoErrors;
  stats->reads_processed = n_reads;
  stats->unambiguous_reads = sc->unambiguousReads;
  uint64_t n = 0;
  for (auto it = jm->junctionMap.begin(); it != jm->junctionMap.end(); ++it, ++n) {
    if (n >= recs_cap) continue;
    ref_junction_rec& r = recs_out[n];
    memset(&r, 0, sizeof(r));
    r.kmer = it->first;
    for (int i = 0; i < 5; i++) { r.dist[i] = it->second.dist[i]; r.linked[i] = it->second.linked[i]; }
    for (int i = 0; i < 4; i++) r.cov[i] = (uint8_t)it->second.getCoverage(i);
  }
  delete sc; delete jm; delete jc; delete bloom; delete spf; delete lpf;
  return 0;
}

}  // extern "C"
