/* TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C99) of the reference's two-pass k-mer hot path.
 * Nothing under faucet_b200/ may include, link or call this; it exists to CHECK the CUDA path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here against the
 * unmodified reference compiled into oracle/_ref/ (hash/revcomp KATs, Bloom geometry, both Bloom bit
 * arrays, the junction map, pair filters, scan counters) and tests/test_golden.py checks it against
 * the reference's own src/newTests/ReadscanTest.cpp vectors and the committed tests/golden fixtures.
 */
#ifndef FAUCET_ORACLE_H
#define FAUCET_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint64_t kmer;     /* oriented k-mer (ReadKmer::getKmer, utils/ReadKmer.cpp:50-57) */
  uint8_t dist[5];   /* utils/Junction.h:18 */
  uint8_t cov[4];    /* utils/Junction.h:12 */
  uint8_t linked[5]; /* utils/Junction.h:19 */
  uint8_t pad[2];
} fo_junction_rec;   /* 24 bytes; returned in CREATION order */

typedef struct {
  uint64_t n_junctions, nb_jcheck_kmer, nb_no_juncs, nb_processed, nb_skipped, reads_no_errors,
      reads_processed, unambiguous_reads;
} fo_scan_stats;

typedef struct {
  uint64_t reads_processed, unambiguous_reads, kmers;
  double weight1, weight2;
} fo_load_stats;

int fo_nt2int(char c);
uint64_t fo_revcomp(uint64_t x, int k);
uint64_t fo_canon(uint64_t x, int k);
uint64_t fo_seed(int i);
uint64_t fo_old_hash(uint64_t key, int i, int log2_tai);
int fo_first_kmer(const char* s, int k, uint64_t* out);
void fo_kmer_string(uint64_t kmer, int k, char* out);

double fo_brent_p1(uint64_t est, uint64_t singletons, float fp);
void fo_geometry_optimal(uint64_t est, float fp, int* log2_tai, int* n_hash);
void fo_geometry_2_hash(uint64_t est, float fp, int* log2_tai, int* n_hash);
double fo_weight(const uint8_t* bits, int log2_tai);

/* pass 1 over an in-memory FASTA/FASTQ text; bloo1/bloo2 are tai/8-byte arrays, updated in place */
int fo_load_two_filters(const char* text, size_t n, int fastq, int k, int log2_tai, int n_hash,
                        uint8_t* bloo1, uint8_t* bloo2, fo_load_stats* stats);

/* pass 2.  fake_set != NULL switches the Bloom into the reference's "fake" mode (exact set of
 * canonical k-mers, sorted ascending).  recs_out is malloc'ed (free with fo_free). */
int fo_scan(const char* text, size_t n, int fastq, int paired_ends, int no_cleaning, int k, int j,
            int max_spacer_dist, const uint8_t* bloo2, int log2_tai, int n_hash, uint8_t* short_pf,
            int spf_log2_tai, int spf_n_hash, uint8_t* long_pf, int lpf_log2_tai, int lpf_n_hash,
            const uint64_t* fake_set, size_t n_fake, fo_junction_rec** recs_out, uint64_t* n_recs_out,
            fo_scan_stats* stats);

/* sequence-per-call form used with the ReadscanTest vectors (scanInputRead on each read) */
int fo_scan_reads(const char* const* reads, int n_reads, int k, int j, int max_spacer_dist,
                  const uint64_t* fake_set, size_t n_fake, fo_junction_rec** recs_out,
                  uint64_t* n_recs_out, fo_scan_stats* stats);

/* "KMER d0 d1 d2 d3 d4  c0 c1 c2 c3 csum  l0 l1 l2 l3 l4 " (utils/Junction.cpp:74-89) */
int fo_junction_line(const fo_junction_rec* r, int k, char* out, size_t cap);

char* fo_read_file(const char* path, size_t* n);
void fo_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
