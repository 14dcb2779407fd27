// The two reference functions of the hot path, re-implemented over libfaucet_gpu's C ABI with the
// reference's EXACT signatures, so that the reference's own main() (src/Faucet.cpp:248-330, compiled
// unmodified) runs its load and scan passes on the B200 without knowing:
//
//   void load_two_filters(Bloom*, Bloom*, std::string, bool, bool)     utils/Bloom.h:294
//   void ReadScanner::scanReads(bool, bool, bool)                       src/ReadScanner.h:66-67
//
// Link recipe (Makefile next to this file): every reference translation unit is compiled where it lies;
// utils/Bloom.cpp with -Dload_two_filters=load_two_filters_cpu and src/ReadScanner.cpp with
// -DscanReads=scanReads_cpu, which renames the two reference bodies out of the way (nothing calls them);
// this file supplies the names main() links against.  Everything downstream of the scan -- the
// std::unordered_map behind JunctionMap, writeToFile, buildContigGraph, cleaning, FASTG/FASTA output --
// is the reference's own code consuming what the GPU produced.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <string>

#include "src/ReadScanner.h"  // -I<reference root>: pulls in Bloom.h, JunctionMap.h, Junction.h, Kmer.h unmodified

#include "faucet_gpu.h"

namespace {
void die(const char* what, int rc) {
  fprintf(stderr, "faucet (GPU drop-in): %s failed (%d): %s\n", what, rc, faucet_gpu_last_error());
  exit(1);
}
}  // namespace

// utils/Bloom.cpp:267-350.  The caller allocated and zeroed both filters (utils/Bloom.cpp:184-186); the bit
// arrays are filled in place.  Prints what the reference prints.
void load_two_filters(Bloom* bloo1, Bloom* bloo2, string reads_filename, bool fastq, bool mercy) {
  if (mercy) {
    fprintf(stderr, "faucet (GPU drop-in): --mercy is not part of the accelerated path (utils/Bloom.cpp:300-333)\n");
    exit(1);
  }
  time_t start, stop;
  time(&start);
  printf("Weights before load: %f, %f \n", bloo1->weight(), bloo2->weight());
  faucet_load_stats st;
  int rc = faucet_gpu_load_two_filters(reads_filename.c_str(), fastq ? 1 : 0, sizeKmer, bloo2->getHashSize(), bloo2->getNumHash(),
                                       bloo2->blooma, bloo1->blooma, &st);
  if (rc) die("faucet_gpu_load_two_filters", rc);
  printf("\n");
  printf("Weights after load: %f, %f \n", bloo1->weight(), bloo2->weight());
  printf("Reads processed: %d\n", (int)st.reads_processed);
  printf("Unambiguous reads: %lli\n", (long long)st.unambiguous_reads);
  time(&stop);
  printf("Time to load: %f \n", difftime(stop, start));
}

// src/ReadScanner.cpp:284-359.  The junction records come back in the reference's creation order, so inserting
// them one by one rebuilds the unordered_map with the reference's iteration order: .junctions and
// everything built from the map come out byte-identical.
void ReadScanner::scanReads(bool fastq, bool paired_ends, bool no_cleaning) {
  NbCandKmer = 0, NbRawCandKmer = 0, NbJCheckKmer = 0, NbNoJuncs = 0, NbSkipped = 0, NbProcessed = 0, readsProcessed = 0,
  NbSolidKmer = 0, readsNoErrors = 0, NbJuncPairs = 0, unambiguousReads = 0;
  time_t start, stop;
  time(&start);
  printf("Weight before read scan: %f \n", bloom->weight());
  faucet_junction_rec* recs = nullptr;
  uint64_t n = 0;
  faucet_scan_stats st;
  Bloom* spf = no_cleaning ? nullptr : short_pair_filter;
  Bloom* lpf = (no_cleaning || !paired_ends) ? nullptr : long_pair_filter;
  int rc = faucet_gpu_scan(reads_file.c_str(), fastq ? 1 : 0, paired_ends ? 1 : 0, no_cleaning ? 1 : 0, sizeKmer, jchecker->j,
                           maxSpacerDist, bloom->blooma, bloom->getHashSize(), bloom->getNumHash(),
                           spf ? spf->blooma : nullptr, spf ? spf->getHashSize() : 0, spf ? spf->getNumHash() : 0,
                           lpf ? lpf->blooma : nullptr, lpf ? lpf->getHashSize() : 0, lpf ? lpf->getNumHash() : 0, &recs, &n, &st);
  if (rc) die("faucet_gpu_scan", rc);
  for (uint64_t i = 0; i < n; i++) {  // recs are sorted by creation_rank
    const faucet_junction_rec& r = recs[i];
    junctionMap->createJunction(r.kmer);
    Junction* jn = junctionMap->getJunction(r.kmer);
    for (int f = 0; f < 5; f++) { jn->dist[f] = r.dist[f]; jn->linked[f] = r.linked[f] != 0; }
    for (int f = 0; f < 4; f++) jn->setCoverage(f, r.cov[f]);
  }
  faucet_gpu_free(recs);
  NbJCheckKmer = st.nb_jcheck_kmer; NbNoJuncs = st.nb_no_juncs; NbProcessed = st.nb_processed; NbSkipped = st.nb_skipped;
  readsNoErrors = st.reads_no_errors; readsProcessed = st.reads_processed; unambiguousReads = st.unambiguous_reads;
  time(&stop);
  printf("Reads processed: %lli\n", (long long)readsProcessed);
  printf("Unambiguous reads: %lli\n", (long long)unambiguousReads);
  printf("Time in seconds for read scan: %f \n", difftime(stop, start));
}
