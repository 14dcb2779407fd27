// faucet (B200 edition) -- host driver for the two streaming passes, C++11, over the C ABI of
// include/faucet_gpu.h.  It keeps the reference's command line (src/Faucet.cpp:57-182) and its file
// formats, so that it is a drop-in for everything up to and including "<prefix>.junctions":
//
//   <prefix>.bloom               raw bit array of bloo2, no header           (utils/Bloom.cpp:571-578)
//   <prefix>.junctions           one text line per junction, in the iteration order of the
//                                reference's std::unordered_map                (utils/JunctionMap.cpp:579-596)
//   <prefix>.short_pair_filter   raw bit arrays, only without --no_cleaning   (src/Faucet.cpp:297-300)
//   <prefix>.long_pair_filter
//
// The contig-graph stage that follows in the reference (JunctionMap::buildContigGraph, cleaning, FASTG
// output) is host-side pointer chasing that consumes these files / structures unchanged; it is out of
// scope here.  Restart with the reference binary via -bloom_file / -junctions_file.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "faucet_gpu.h"

namespace {

struct Options {  // globals of src/Faucet.h:14-53
  float fpRate = .04f;
  int j = 1;
  std::string read_load_file, read_scan_file, bloom_input_file, junctions_prefix, file_prefix;
  int read_length = 0, size_kmer = 0;
  uint64_t estimated_kmers = 0, singletons = 0;
  bool load_file_flag = false, scan_file_flag = false, k_val_flag = false, max_len_flag = false, est_kmers_flag = false,
       est_sing_flag = false, pref_flag = false;
  bool two_hash = false, from_bloom = false, from_junctions = false, just_load = false, fastq = false, mercy = false,
       node_graph = false, paired_ends = false, no_cleaning = false, high_cov = false;
  int maxSpacerDist = 100;
  int device = 0;
};

void argumentError() {  // the reference's usage text, abridged (src/Faucet.cpp:26-55)
  fprintf(stderr,
          "Usage: faucet -read_load_file <file> -read_scan_file <file> -size_kmer <k> -max_read_length <len>\n"
          "  -estimated_kmers <n> -singletons <n> -file_prefix <prefix>\n"
          "  [-fp rate] [-j j] [-max_spacer_dist d] [--fastq] [--paired_ends] [--two_hash] [--no_cleaning]\n"
          "  [--just_load_bloom] [-bloom_file file] [-junctions_file prefix] [--mercy] [--high_cov] [--node_graph]\n"
          "  [-gpu device]\n");
}

int handle_arguments(int argc, char* argv[], Options& o) {
  if (argc == 1) { argumentError(); return 1; }
  for (int i = 1; i < argc; i++) {
    const char* a = argv[i];
    auto need = [&](const char* what) -> const char* {
      if (i + 1 >= argc) { fprintf(stderr, "Missing value after %s\n", what); exit(1); }
      return argv[++i];
    };
    if (!strcmp(a, "-read_load_file")) o.read_load_file = need(a), o.load_file_flag = true;
    else if (!strcmp(a, "-read_scan_file")) o.read_scan_file = need(a), o.scan_file_flag = true;
    else if (!strcmp(a, "-size_kmer")) o.size_kmer = atoi(need(a)), o.k_val_flag = true;
    else if (!strcmp(a, "-max_read_length")) o.read_length = atoi(need(a)), o.max_len_flag = true;
    else if (!strcmp(a, "-estimated_kmers")) o.estimated_kmers = (uint64_t)atoll(need(a)), o.est_kmers_flag = true;
    else if (!strcmp(a, "-singletons")) o.singletons = (uint64_t)atoll(need(a)), o.est_sing_flag = true;
    else if (!strcmp(a, "-fp")) o.fpRate = (float)atof(need(a));
    else if (!strcmp(a, "-j")) o.j = atoi(need(a));
    else if (!strcmp(a, "-file_prefix")) o.file_prefix = need(a), o.pref_flag = true;
    else if (!strcmp(a, "--two_hash")) o.two_hash = true;
    else if (!strcmp(a, "--just_load_bloom")) o.just_load = true;
    else if (!strcmp(a, "--no_cleaning")) o.no_cleaning = true;
    else if (!strcmp(a, "--fastq")) o.fastq = true;
    else if (!strcmp(a, "--mercy")) o.mercy = true;
    else if (!strcmp(a, "--high_cov")) o.high_cov = true;
    else if (!strcmp(a, "--node_graph")) o.node_graph = true;
    else if (!strcmp(a, "--paired_ends")) o.paired_ends = true;
    else if (!strcmp(a, "-bloom_file")) o.bloom_input_file = need(a), o.from_bloom = true;
    else if (!strcmp(a, "-max_spacer_dist")) o.maxSpacerDist = atoi(need(a));
    else if (!strcmp(a, "-junctions_file")) o.junctions_prefix = need(a), o.from_junctions = true;
    else if (!strcmp(a, "-gpu")) o.device = atoi(need(a));
    else if (!strcmp(a, "--help") || !strcmp(a, "-h")) { argumentError(); return 1; }
    else { fprintf(stderr, "Cannot parse tag %s\n", a); argumentError(); return 1; }
  }
  if (!(o.load_file_flag && o.scan_file_flag && o.k_val_flag && o.max_len_flag && o.est_kmers_flag && o.est_sing_flag && o.pref_flag)) {
    fprintf(stderr, "Some required argument is missing.\n");
    argumentError();
    return 1;
  }
  if (o.from_junctions && !o.from_bloom) {
    fprintf(stderr, "Cannot start from junctions without a bloom file.\n");
    argumentError();
    return 1;
  }
  if (o.size_kmer < 2 || o.size_kmer > 32) { fprintf(stderr, "-size_kmer must be in [2,32] (k-mers live in 64 bits)\n"); return 1; }
  if (o.from_junctions && o.from_bloom) printf("Starting from after read scan based on bloom and junction files.\n");
  else if (o.from_bloom) printf("Starting from after bloom load based on bloom file.\n");
  else printf("Starting at the beginning: will load bloom and find junctions from the read set.\n");
  if (o.just_load) printf("Only loading bloom, dumping and termination.\n");
  std::cout << "Read load file name: " << o.read_load_file << "\n";
  std::cout << "Read scan file name: " << o.read_scan_file << "\n";
  printf("k: %d \n", o.size_kmer);
  printf("Maximal read length: %d\n", o.read_length);
  printf("Estimated number of distinct kmers, for sizing bloom filter: %lli.\n", (long long)o.estimated_kmers);
  printf("False positive rate: %f\n", o.fpRate);
  printf("File prefix: %s\n", o.file_prefix.c_str());
  printf("Max spacer dist: %d\n", o.maxSpacerDist);
  printf(o.two_hash ? "Using 2 hash functions.\n" : "Using space-optimal hash settings.\n");
  std::cout << "Paired ends: " << o.paired_ends << "\n";
  return 0;
}

void die(const char* what) {
  fprintf(stderr, "faucet: %s: %s\n", what, faucet_gpu_last_error());
  exit(2);
}

bool write_file(const std::string& path, const void* p, size_t n) {  // Bloom::dump (utils/Bloom.cpp:571-578)
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { perror(path.c_str()); return false; }
  size_t w = n ? fwrite(p, 1, n, f) : 0;
  fclose(f);
  return w == n;
}

double weight(const std::vector<uint8_t>& bits) {  // Bloom::weight (utils/Bloom.cpp:191-203): float division
  uint64_t c = 0;
  for (uint8_t b : bits) c += (uint64_t)__builtin_popcount(b);
  return (double)((float)(long)c / (float)(bits.size() * 8));
}

// print_kmer (utils/Kmer.cpp:555-564): 2-bit codes back to letters, A0 C1 T2 G3
std::string kmer_string(uint64_t x, int k) {
  static const char NT[4] = {'A', 'C', 'T', 'G'};
  std::string s((size_t)k, 'A');
  for (int i = 0; i < k; i++) s[i] = NT[(x >> (2 * (k - 1 - i))) & 3];
  return s;
}

struct Junction { uint8_t dist[5], cov[4], linked[5]; };

// JunctionMap::writeToFile (utils/JunctionMap.cpp:579-596) + Junction::toString (utils/Junction.cpp:74-89).
// The records come sorted by creation rank; inserting them in that order into the same container type
// the reference uses reproduces its iteration order.
bool write_junctions(const std::string& path, const faucet_junction_rec* recs, uint64_t n, int k) {
  std::unordered_map<uint64_t, Junction> map;
  for (uint64_t i = 0; i < n; i++) {
    Junction& j = map[recs[i].kmer];
    memcpy(j.dist, recs[i].dist, 5); memcpy(j.cov, recs[i].cov, 4); memcpy(j.linked, recs[i].linked, 5);
  }
  int solid[5] = {0, 0, 0, 0, 0};  // getNumSolidJunctions(i) / Junction::isSolid(i): > 1 extension with cov >= i
  for (auto& kv : map)
    for (int i = 0; i < 5; i++) {
      int paths = 0;
      for (int e = 0; e < 4; e++) paths += kv.second.cov[e] >= i;
      solid[i] += paths > 1;
    }
  for (int i = 0; i < 5; i++) printf("There are %d junctions with solidity at least %d.\n", solid[i], i);
  printf("Writing to junction file\n");
  std::ofstream f(path);
  if (!f) { perror(path.c_str()); return false; }
  for (auto it = map.begin(); it != map.end(); ++it) {
    const Junction& j = it->second;
    f << kmer_string(it->first, k) << " ";
    for (int i = 0; i < 5; i++) f << (int)j.dist[i] << " ";
    f << " ";
    int sum = 0;
    for (int i = 0; i < 4; i++) { f << (int)j.cov[i] << " "; sum += j.cov[i]; }
    f << sum << " ";
    f << " ";
    for (int i = 0; i < 5; i++) f << (j.linked[i] ? 1 : 0) << " ";
    f << "\n";
  }
  printf("Done writing to junction file\n");
  return true;
}

// JunctionMap::buildFromFile (utils/JunctionMap.cpp:619-639) + Junction::Junction(string) (utils/Junction.cpp:102-118):
// "KMER d0..d4  c0..c3 csum  l0..l4 " per line; a later line with the same k-mer replaces the earlier one.
bool read_junctions(const std::string& path, int k, std::unordered_map<uint64_t, Junction>& map) {
  std::ifstream f(path);
  printf("Reading from Junction file to build junction map.\n");
  std::string line, word;
  while (std::getline(f, line)) {
    std::istringstream iss(line);
    iss >> word;
    if ((int)word.size() < k) continue;
    uint64_t kmer = 0;  // getFirstKmerFromRead: A=0 C=1 T=2 G=3 (utils/Kmer.cpp:82-88, 429-433)
    for (int i = 0; i < k; i++) kmer = (kmer << 2) | (uint64_t)((word[i] >> 1) & 3);
    Junction j;
    int v = 0;
    for (int i = 0; i < 5; i++) { iss >> v; j.dist[i] = (uint8_t)v; }
    for (int i = 0; i < 4; i++) { iss >> v; j.cov[i] = (uint8_t)v; }
    iss >> v;  // the coverage sum
    for (int i = 0; i < 5; i++) { iss >> v; j.linked[i] = v != 0; }
    map[kmer] = j;
  }
  return true;
}
// Bloom::load (utils/Bloom.cpp:580-587): raw bytes, unchecked there
void load_file(const std::string& path, std::vector<uint8_t>& bits) {
  FILE* f = fopen(path.c_str(), "rb");
  if (f) { size_t got = fread(bits.data(), 1, bits.size(), f); (void)got; fclose(f); }
}

}  // namespace

int main(int argc, char* argv[]) {
  Options o;
  if (handle_arguments(argc, argv, o) == 1) return 1;
  if (o.mercy) { fprintf(stderr, "faucet: --mercy is not offloaded (run the reference for that flag)\n"); return 1; }
  if (!(o.from_bloom && o.from_junctions)) {  // (that restart skips both passes: no device work)
    if (faucet_gpu_init(o.device)) die("init");
    printf("Device path: %s\n", faucet_gpu_version());
  }

  // ---- Bloom filter: from reads (pass 1) or from a file ------------------------------------------
  int log2_tai = 0, n_hash = 0;
  bool same_file = false;
  std::vector<uint8_t> bloom;
  if (o.from_bloom) {  // getBloomFilterFromFile (src/Faucet.cpp:185-195): geometry from -fp, --two_hash honoured HERE only
    if (o.two_hash ? faucet_geometry_2_hash(o.estimated_kmers, o.fpRate, &log2_tai, &n_hash)
                   : faucet_geometry_optimal(o.estimated_kmers, o.fpRate, &log2_tai, &n_hash)) die("geometry");
    bloom.assign(((size_t)1 << log2_tai) / 8, 0);
    FILE* f = fopen(o.bloom_input_file.c_str(), "rb");  // Bloom::load (utils/Bloom.cpp:580-587), unchecked there
    if (f) { size_t got = fread(bloom.data(), 1, bloom.size(), f); (void)got; fclose(f); }
    printf("Weight of bloom filter: %f\n", weight(bloom));
  } else {  // getBloomFilterFromReads (src/Faucet.cpp:204-223); --two_hash is a no-op on this path
    double p1 = 0;
    if (faucet_geometry_from_reads(o.estimated_kmers, o.singletons, o.fpRate, &p1, &log2_tai, &n_hash)) die("geometry");
    printf("Optimal fp rate for bloom 1: %f\n", p1);
    bloom.assign(((size_t)1 << log2_tai) / 8, 0);
    faucet_load_stats st;
    printf("Weights before load: %f, %f \n", 0.0, 0.0);
    // one reads file for both passes (the default): pass 1 leaves the parsed planes in HBM and pass 2 runs on them
    same_file = !o.just_load && o.read_load_file == o.read_scan_file;
    if (same_file && faucet_gpu_set_tuning("retain_planes", 1)) die("tuning");
    if (faucet_gpu_load_two_filters(o.read_load_file.c_str(), o.fastq, o.size_kmer, log2_tai, n_hash, bloom.data(), nullptr, &st))
      die("load_two_filters");
    printf("Weights after load: %f, %f \n", st.weight1, st.weight2);
    printf("Reads processed: %lli\n", (long long)st.reads_processed);
    printf("Unambiguous reads: %lli\n", (long long)st.unambiguous_reads);
    if (!write_file(o.file_prefix + ".bloom", bloom.data(), bloom.size())) return 2;
  }

  // ---- pair filters (src/Faucet.cpp:265-281) ------------------------------------------------------
  const uint64_t s_items = o.high_cov ? o.estimated_kmers / 2 : o.estimated_kmers / 20;
  const uint64_t l_items = o.high_cov ? o.estimated_kmers / 2 : o.estimated_kmers / 10;
  int s_log2 = 0, s_nh = 0, l_log2 = 0, l_nh = 0;
  if (faucet_geometry_optimal(s_items, 0.01f, &s_log2, &s_nh)) die("geometry");
  std::vector<uint8_t> spf(((size_t)1 << s_log2) / 8, 0), lpf;
  if (o.paired_ends) {
    if (faucet_geometry_optimal(l_items, 0.01f, &l_log2, &l_nh)) die("geometry");
    lpf.assign(((size_t)1 << l_log2) / 8, 0);
  }
  if (o.just_load) return 0;

  // ---- restart after both passes (src/Faucet.cpp:289-293): JunctionMap::buildFromFile + Bloom::load of the pair filters.
  // Both streaming passes are skipped, so there is no device work; what is left is the reference's graph stage.
  if (o.from_junctions) {
    std::unordered_map<uint64_t, Junction> map;
    if (!read_junctions(o.junctions_prefix + ".junctions", o.size_kmer, map)) return 2;
    load_file(o.junctions_prefix + ".short_pair_filter", spf);
    if (o.paired_ends) load_file(o.junctions_prefix + ".long_pair_filter", lpf);
    printf("Weight of short pair filter: %f\n", weight(spf));
    if (o.paired_ends) printf("Weight of long pair filter: %f\n", weight(lpf));
    printf("Number of junctions: %llu\n", (unsigned long long)map.size());
    printf("Junction map and pair filters read back; contig graph construction is the reference's host stage.\n");
    return 0;
  }

  // ---- pass 2 (buildJunctionMapFromReads, src/Faucet.cpp:240-246) ---------------------------------
  faucet_junction_rec* recs = nullptr;
  uint64_t n = 0;
  faucet_scan_stats st;
  printf("Weight before read scan: %f \n", weight(bloom));
  int rc = FAUCET_E_STATE;
  if (same_file)
    rc = faucet_gpu_scan_retained(o.paired_ends, o.no_cleaning, o.size_kmer, o.j, o.maxSpacerDist, nullptr, log2_tai, n_hash,
                                  spf.data(), s_log2, s_nh, o.paired_ends ? lpf.data() : nullptr, l_log2, l_nh, &recs, &n, &st);
  if (rc == FAUCET_E_STATE)  // other file, or more planes than the retention budget: read the scan file
    rc = faucet_gpu_scan(o.read_scan_file.c_str(), o.fastq, o.paired_ends, o.no_cleaning, o.size_kmer, o.j, o.maxSpacerDist,
                         bloom.data(), log2_tai, n_hash, spf.data(), s_log2, s_nh, o.paired_ends ? lpf.data() : nullptr, l_log2, l_nh,
                         &recs, &n, &st);
  if (rc) die("scan");
  printf("Reads processed: %lli\n", (long long)st.reads_processed);
  printf("Unambiguous reads: %lli\n", (long long)st.unambiguous_reads);
  // ReadScanner::printScanSummary (src/ReadScanner.cpp:19-27)
  printf("\nDistinct junctions: %lli \n", (long long)st.n_junctions);
  printf("Number of kmers that we j-checked: %lli \n", (long long)st.nb_jcheck_kmer);
  printf("Number of reads with no junctions: %lli \n", (long long)st.nb_no_juncs);
  printf("Number of processed kmers: %lli \n", (long long)st.nb_processed);
  printf("Number of skipped kmers: %lli \n", (long long)st.nb_skipped);
  printf("Reads without errors: %lli\n", (long long)st.reads_no_errors);
  if (!write_junctions(o.file_prefix + ".junctions", recs, n, o.size_kmer)) return 2;
  faucet_gpu_free(recs);
  if (!o.no_cleaning) {
    if (!write_file(o.file_prefix + ".short_pair_filter", spf.data(), spf.size())) return 2;
    if (o.paired_ends && !write_file(o.file_prefix + ".long_pair_filter", lpf.data(), lpf.size())) return 2;
  }
  printf("Weight of short pair filter: %f\n", weight(spf));
  if (o.paired_ends) printf("Weight of long pair filter: %f\n", weight(lpf));
  printf("Number of junctions: %llu\n", (unsigned long long)n);
  printf("Streaming passes done; contig graph construction is the reference's host stage "
         "(restart it with -bloom_file %s.bloom -junctions_file %s).\n", o.file_prefix.c_str(), o.file_prefix.c_str());
  faucet_gpu_shutdown();
  return 0;
}
