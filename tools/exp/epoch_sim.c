/* Experiment (not product, not a test): how large is the "exact set" of the epoch stitch?
 *
 * Runs the oracle's scan with trace hooks, logs every WRITE (junction created / dist raised / link newly set)
 * with the record that caused it, then for a schedule of epochs computes
 *   writers  = records that wrote anything
 *   E        = writers + records that share a minimizer slot with an EARLIER write of the same epoch
 * which is what the epoch stitch has to run through the ordered kernel.
 *
 *   gcc -O2 -DFO_TRACE -o epoch_sim epoch_sim.c -lm && ./epoch_sim reads.fq k est sing e0 growth emax
 */
#define FO_TRACE 1
#include "../../oracle/faucet_oracle.c"

typedef struct { uint32_t rec; uint8_t kind; uint64_t key; } wr_t;
static wr_t* W; static size_t nW, capW;
static uint32_t cur_rec = (uint32_t)-1;
void fo_trace_record(void) { cur_rec++; }
void fo_trace_write(uint64_t key, int kind) {
  if (nW == capW) { capW = capW ? capW * 2 : 1 << 20; W = (wr_t*)realloc(W, capW * sizeof(wr_t)); }
  W[nW].rec = cur_rec; W[nW].kind = (uint8_t)kind; W[nW].key = key; nW++;
}

static uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16; return x; }
static uint32_t smer_h(uint32_t x, int s) {
  uint32_t r = 0, y = x;
  for (int i = 0; i < s; i++) { r = (r << 2) | ((y & 3) ^ 2); y >>= 2; }
  return mix32(x < r ? x : r);
}
#define RES_MASK ((1u << 24) - 1)
static uint32_t kmer_slot(uint64_t key, int k) {
  int s = k < 16 ? k : 16;
  uint32_t m = 0xffffffffu, smask = s == 16 ? 0xffffffffu : ((1u << (2 * s)) - 1);
  for (int i = 0; i + s <= k; i++) {
    uint32_t x = (uint32_t)(key >> (2 * (k - s - i))) & smask;
    uint32_t h = smer_h(x, s);
    if (h < m) m = h;
  }
  return m & RES_MASK;
}

int main(int argc, char** argv) {
  if (argc < 8) { fprintf(stderr, "usage: epoch_sim reads.fq k est sing e0 growth emax\n"); return 2; }
  size_t n; char* text = fo_read_file(argv[1], &n);
  int k = atoi(argv[2]);
  uint64_t est = strtoull(argv[3], 0, 10), sing = strtoull(argv[4], 0, 10);
  uint32_t e0 = atoi(argv[5]); double growth = atof(argv[6]); uint32_t emax = atoi(argv[7]);
  double p1 = fo_brent_p1(est, sing, 0.04f);
  int lt, nh; fo_geometry_optimal(est, (float)p1, &lt, &nh);
  uint8_t* b1 = calloc(1ull << lt >> 3, 1); uint8_t* b2 = calloc(1ull << lt >> 3, 1);
  fo_load_stats ls; fo_load_two_filters(text, n, 1, k, lt, nh, b1, b2, &ls);
  fo_junction_rec* recs; uint64_t nrec; fo_scan_stats st;
  fo_scan(text, n, 1, 1, 1, k, 1, 100, b2, lt, nh, 0, 0, 0, 0, 0, 0, 0, 0, &recs, &nrec, &st);
  uint32_t R = (uint32_t)st.reads_processed;
  fprintf(stderr, "records %u junctions %llu writes %zu  log2_tai %d n_hash %d w2 %.3f\n", R, (unsigned long long)nrec, nW, lt, nh, ls.weight2);
  /* per-record slots */
  uint32_t* rs = malloc((size_t)R * 32 * 4); uint8_t* rn = calloc(R, 1);
  { size_t pos = 0; uint32_t r = 0;
    while (pos < n && r < R) {
      while (pos < n && text[pos] != '\n') pos++; pos++;
      size_t a = pos; while (pos < n && text[pos] != '\n') pos++; size_t len = pos - a; pos++;
      for (int q = 0; q < 2; q++) { while (pos < n && text[pos] != '\n') pos++; pos++; }
      int s = k < 16 ? k : 16;
      uint32_t prev = 0xffffffffu; int cnt = 0;
      if ((int)len >= k) {
        uint32_t* sh = malloc(len * 4);
        for (size_t i = 0; i + s <= len; i++) { uint32_t x = 0; for (int t = 0; t < s; t++) x = (x << 2) | (uint32_t)fo_nt2int(text[a + i + t]); sh[i] = smer_h(x, s); }
        for (size_t p = 0; p + k <= len; p++) {
          uint32_t m = 0xffffffffu; for (int i = 0; i + s <= k; i++) if (sh[p + i] < m) m = sh[p + i];
          if (p == 0 || m != prev) { if (cnt < 32) rs[(size_t)r * 32 + cnt] = m & RES_MASK; cnt++; }
          prev = m;
        }
        free(sh);
      }
      rn[r] = (uint8_t)(cnt > 32 ? 32 : cnt); r++;
    }
  }
  { /* depth of the dependency DAG of the dataflow executor: a record runs after every earlier record sharing a slot */
    uint32_t* last = calloc(RES_MASK + 1ull, 4); uint32_t maxd = 0; double sum = 0;
    for (uint32_t r = 0; r < R; r++) {
      uint32_t d = 0;
      for (int i = 0; i < rn[r]; i++) if (last[rs[(size_t)r * 32 + i]] > d) d = last[rs[(size_t)r * 32 + i]];
      d++;
      for (int i = 0; i < rn[r]; i++) last[rs[(size_t)r * 32 + i]] = d;
      if (d > maxd) maxd = d; sum += d;
      if ((r + 1) % (R / 10) == 0) fprintf(stderr, "  after %u records: DAG depth %u\n", r + 1, maxd);
    }
    fprintf(stderr, "DAG depth %u for %u records (mean level %.1f)\n", maxd, R, sum / R);
    free(last);
  }
  uint8_t* writer = calloc(R, 1);
  for (size_t i = 0; i < nW; i++) writer[W[i].rec] = 1;
  uint32_t* dirty = malloc((RES_MASK + 1ull) * 4);
  size_t wi = 0; uint32_t a = 0, esz = e0;
  uint64_t totE = 0, totW = 0;
  printf("%10s %10s %9s %9s %7s %7s\n", "start", "size", "writers", "E", "w%", "E%");
  while (a < R) {
    uint32_t b = a + esz < R ? a + esz : R;
    memset(dirty, 0xff, (RES_MASK + 1ull) * 4);
    size_t w0 = wi;
    while (wi < nW && W[wi].rec < b) { if (W[wi].kind != 2) { uint32_t sl = kmer_slot(W[wi].key, k); if (W[wi].rec < dirty[sl]) dirty[sl] = W[wi].rec; } wi++; }
    (void)w0;
    uint32_t nw = 0, ne = 0;
    for (uint32_t r = a; r < b; r++) {
      int in = writer[r]; nw += writer[r];
      for (int i = 0; i < rn[r] && !in; i++) if (dirty[rs[(size_t)r * 32 + i]] < r) in = 1;
      ne += in;
    }
    printf("%10u %10u %9u %9u %6.2f%% %6.2f%%\n", a, b - a, nw, ne, 100.0 * nw / (b - a), 100.0 * ne / (b - a));
    totE += ne; totW += nw;
    a = b; esz = (uint32_t)(esz * growth); if (esz > emax) esz = emax;
  }
  printf("total: records %u writers %llu (%.2f%%) E %llu (%.2f%%)\n", R, (unsigned long long)totW, 100.0 * totW / R, (unsigned long long)totE, 100.0 * totE / R);
  return 0;
}
