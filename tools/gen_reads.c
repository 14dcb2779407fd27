/* Deterministic synthetic read generator for the BASELINE.json configs.
 *
 * uniform-random genome over ACGT from a fixed seed, optional planted repeat families (so that the
 * reference's graph cleaning survives), fragments sampled uniformly, random strand,
 * mate 2 = reverse complement of the fragment's far end, interleaved 4-line FASTQ (or 2-line FASTA)
 * with constant quality, optional i.i.d. substitution errors and optional 'N' bases.
 *
 *   gen_reads -o out.fq -g 1000000 -c 30 -l 100 -i 300 -s 1 [-e 0.005] [-r] [-n 0.001] [-a] [-p N] [-S stream]
 *
 * Uses its own splitmix64/xoshiro256** so the byte stream is identical on every platform.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[4];
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t splitmix(uint64_t* x) {
  uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
static void seed_rng(uint64_t seed) { for (int i = 0; i < 4; i++) s[i] = splitmix(&seed); }
static inline uint64_t next(void) {
  uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
  return r;
}
static inline uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
static inline double unif(void) { return (next() >> 11) * (1.0 / 9007199254740992.0); }

static const char NT[4] = {'A', 'C', 'G', 'T'};
static inline char comp(char c) {
  switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; }
}

int main(int argc, char** argv) {
  const char* out = NULL;
  uint64_t G = 1000000, seed = 1, max_pairs = 0, stream = 0;
  double cov = 30, err = 0, nrate = 0;
  int L = 100, insert = 300, repeats = 0, fasta = 0, lower = 0;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-o")) out = argv[++i];
    else if (!strcmp(argv[i], "-g")) G = strtoull(argv[++i], 0, 10);
    else if (!strcmp(argv[i], "-c")) cov = atof(argv[++i]);
    else if (!strcmp(argv[i], "-l")) L = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-i")) insert = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-s")) seed = strtoull(argv[++i], 0, 10);
    else if (!strcmp(argv[i], "-e")) err = atof(argv[++i]);
    else if (!strcmp(argv[i], "-n")) nrate = atof(argv[++i]);   /* per-base probability of an 'N' */
    else if (!strcmp(argv[i], "-r")) repeats = 1;               /* plant repeat families */
    else if (!strcmp(argv[i], "-a")) fasta = 1;                 /* 2-line FASTA instead of FASTQ */
    else if (!strcmp(argv[i], "-w")) lower = 1;                 /* sprinkle lowercase bases (rate = nrate) */
    else if (!strcmp(argv[i], "-p")) max_pairs = strtoull(argv[++i], 0, 10);
    else if (!strcmp(argv[i], "-S")) stream = strtoull(argv[++i], 0, 10);  /* same genome, another read sample */
    else { fprintf(stderr, "unknown flag %s\n", argv[i]); return 2; }
  }
  if (!out || insert < L || G < (uint64_t)insert) { fprintf(stderr, "usage: gen_reads -o FILE [-g G -c COV -l L -i INSERT -s SEED -e ERR -n NRATE -r -a -p PAIRS]\n"); return 2; }
  seed_rng(seed);
  char* g = (char*)malloc(G + 1);
  for (uint64_t i = 0; i < G; i++) g[i] = NT[next() & 3];
  if (repeats) { /* 8 families x 400 bp x 6 copies per Mbp */
    uint64_t fam = 8 * ((G + 999999) / 1000000);
    for (uint64_t f = 0; f < fam && G > 4000; f++) {
      uint64_t src = below(G - 400);
      char unit[400];
      memcpy(unit, g + src, 400);
      for (int c = 0; c < 5; c++) memcpy(g + below(G - 400), unit, 400);
    }
  }
  if (stream) seed_rng(seed ^ (0x9e3779b97f4a7c15ULL * stream));  /* the genome above depends on -s only */
  uint64_t pairs = (uint64_t)(G * cov / (2.0 * L));
  if (max_pairs && pairs > max_pairs) pairs = max_pairs;
  FILE* f = fopen(out, "wb");
  if (!f) { perror(out); return 1; }
  static char buf[1 << 22];
  setvbuf(f, buf, _IOFBF, sizeof buf);
  char* frag = (char*)malloc(insert + 1);
  char* r1 = (char*)malloc(L + 1);
  char* r2 = (char*)malloc(L + 1);
  char* q = (char*)malloc(L + 1);
  memset(q, 'I', L); q[L] = 0; r1[L] = r2[L] = 0;
  for (uint64_t p = 0; p < pairs; p++) {
    uint64_t st = below(G - insert + 1);
    if (next() & 1) memcpy(frag, g + st, insert);
    else for (int i = 0; i < insert; i++) frag[i] = comp(g[st + insert - 1 - i]);
    memcpy(r1, frag, L);
    for (int i = 0; i < L; i++) r2[i] = comp(frag[insert - 1 - i]);
    if (err > 0) for (int m = 0; m < 2; m++) { char* r = m ? r2 : r1;
      for (int i = 0; i < L; i++) if (unif() < err) { char c; do c = NT[next() & 3]; while (c == r[i]); r[i] = c; } }
    if (nrate > 0) for (int m = 0; m < 2; m++) { char* r = m ? r2 : r1;
      for (int i = 0; i < L; i++) if (unif() < nrate) r[i] = lower ? (char)(r[i] | 0x20) : 'N'; }
    if (fasta) fprintf(f, ">r%llu/1\n%s\n>r%llu/2\n%s\n", (unsigned long long)p, r1, (unsigned long long)p, r2);
    else fprintf(f, "@r%llu/1\n%s\n+\n%s\n@r%llu/2\n%s\n+\n%s\n", (unsigned long long)p, r1, q, (unsigned long long)p, r2, q);
  }
  fclose(f);
  return 0;
}
