/* TEST INFRASTRUCTURE ONLY -- see faucet_oracle.h.  Plain C99 restatement of the reference's
 * load pass (utils/Bloom.cpp:267-299) and scan pass (src/ReadScanner.cpp) written from the
 * specification in SURVEY.md Appendix A; every function cites the reference lines it follows.
 * Parity status: PINNED against oracle/_ref (the unmodified reference) by tests/test_oracle_vs_ref.py.
 */
#include "faucet_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* optional instrumentation for tools/exp/epoch_sim.c (never defined in the library build) */
#ifdef FO_TRACE
void fo_trace_write(uint64_t key, int kind);
void fo_trace_record(void);
#define TRACE_WRITE(key, kind) fo_trace_write(key, kind)
#define TRACE_RECORD() fo_trace_record()
#else
#define TRACE_WRITE(key, kind) ((void)0)
#define TRACE_RECORD() ((void)0)
#endif

/* ------------------------------------------------------------------ k-mer codec (utils/Kmer.cpp) */

/* NT2int, utils/Kmer.cpp:82-88: A=0 C=1 T=2 G=3 */
int fo_nt2int(char c) { return (c >> 1) & 3; }
/* revcomp_int, utils/Kmer.cpp:90-93 */
static int comp_nt(int nt) { return nt < 2 ? nt + 2 : nt - 2; }
/* isValidNuc, utils/Kmer.cpp:50-60 */
static int valid_nt(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
static uint64_t kmask(int k) { return k == 32 ? ~(uint64_t)0 : (((uint64_t)1 << (2 * k)) - 1); }

/* revcomp(uint64,size), utils/Kmer.cpp:238-252 (byte LUT there; per-nucleotide loop here) */
uint64_t fo_revcomp(uint64_t x, int k) {
  uint64_t r = 0;
  for (int i = 0; i < k; i++) {
    r = (r << 2) | (uint64_t)comp_nt((int)(x & 3));
    x >>= 2;
  }
  return r;
}
/* get_canon, utils/Kmer.cpp:531-533 */
uint64_t fo_canon(uint64_t x, int k) {
  uint64_t r = fo_revcomp(x, k);
  return x < r ? x : r;
}
/* shift_kmer FORWARD branch, utils/Kmer.cpp:419-423 */
static uint64_t ext_fwd(uint64_t x, int nt, int k) { return ((x << 2) + (uint64_t)nt) & kmask(k); }
/* shift_kmer BACKWARD branch, utils/Kmer.cpp:414-418 */
static uint64_t ext_bwd(uint64_t x, int nt, int k) {
  return ((x >> 2) + ((uint64_t)nt << (2 * k - 2))) & kmask(k);
}
/* getFirstKmerFromRead, utils/Kmer.cpp:429-433 */
int fo_first_kmer(const char* s, int k, uint64_t* out) {
  uint64_t x = 0;
  for (int i = 0; i < k; i++) x = ext_fwd(x, fo_nt2int(s[i]), k);
  *out = x;
  return 0;
}
/* code2seq, utils/Kmer.cpp:213-227 */
void fo_kmer_string(uint64_t kmer, int k, char* out) {
  static const char nt[4] = {'A', 'C', 'T', 'G'};
  for (int i = k - 1; i >= 0; i--) {
    out[i] = nt[kmer & 3];
    kmer >>= 2;
  }
  out[k] = 0;
}

/* ------------------------------------------------------------------ Bloom (utils/Bloom.h/.cpp) */

/* generate_hash_seed with user_seed = 0, utils/Bloom.cpp:500-511 (in-place update order kept) */
uint64_t fo_seed(int i) {
  static const uint64_t rbase[10] = {0xAAAAAAAA55555555ULL, 0x33333333CCCCCCCCULL, 0x6666666699999999ULL,
                                     0xB5B5B5B54B4B4B4BULL, 0xAA55AA5555335533ULL, 0x33CC33CCCC66CC66ULL,
                                     0x6699669999B599B5ULL, 0xB54BB54B4BAA4BAAULL, 0xAA33AA3355CC55CCULL,
                                     0x33663366CC99CC99ULL};
  uint64_t tab[10];
  for (int a = 0; a < 10; a++) tab[a] = rbase[a];
  for (int a = 0; a < 10; a++) tab[a] = tab[a] * tab[(a + 3) % 10];
  return tab[i];
}

/* Bloom::oldHash, utils/Bloom.h:134-145 */
uint64_t fo_old_hash(uint64_t key, int i, int log2_tai) {
  uint64_t h = fo_seed(i);
  h ^= (h << 7) ^ key * (h >> 3) ^ (~((h << 11) + (key ^ (h >> 5))));
  h = (~h) + (h << 21);
  h = h ^ (h >> 24);
  h = (h + (h << 3)) + (h << 8);
  h = h ^ (h >> 14);
  h = (h + (h << 2)) + (h << 4);
  h = h ^ (h >> 28);
  h = h + (h << 31);
  return h & (((uint64_t)1 << log2_tai) - 1);
}

typedef struct {
  uint8_t* bits;
  int log2_tai, n_hash;
  uint64_t mask;
  const uint64_t* fake; /* sorted canonical k-mers: Bloom::fakify/addFakeKmers, utils/Bloom.cpp:156-162 */
  size_t n_fake;
  uint64_t seed0, seed1;
} bloom_t;

static void bloom_init(bloom_t* b, uint8_t* bits, int log2_tai, int n_hash) {
  memset(b, 0, sizeof *b);
  b->bits = bits;
  b->log2_tai = log2_tai;
  b->n_hash = n_hash;
  b->mask = ((uint64_t)1 << log2_tai) - 1;
}
/* Bloom::add(h0,h1), utils/Bloom.h:217-226 */
static void bloom_add(bloom_t* b, uint64_t h0, uint64_t h1) {
  uint64_t h = h0;
  for (int i = 0; i < b->n_hash; i++, h += h1) {
    h &= b->mask;
    b->bits[h >> 3] |= (uint8_t)(1u << (h & 7));
  }
}
/* Bloom::contains(h0,h1), utils/Bloom.h:242-258 */
static int bloom_contains(const bloom_t* b, uint64_t h0, uint64_t h1) {
  uint64_t h = h0 & b->mask;
  for (int i = 0; i < b->n_hash; i++, h = (h + h1) & b->mask)
    if (!(b->bits[h >> 3] & (1u << (h & 7)))) return 0;
  return 1;
}
static int fake_has(const bloom_t* b, uint64_t x) {
  size_t lo = 0, hi = b->n_fake;
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    if (b->fake[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo < b->n_fake && b->fake[lo] == x;
}
/* Bloom::oldContains, utils/Bloom.h:162-173 */
static int bloom_old_contains(const bloom_t* b, uint64_t canon) {
  if (b->fake) return fake_has(b, canon);
  return bloom_contains(b, fo_old_hash(canon, 0, b->log2_tai), fo_old_hash(canon, 1, b->log2_tai));
}
/* Bloom::addPair / containsPair, utils/Bloom.cpp:127-154 */
static void pair_hashes(const bloom_t* b, uint64_t k1, uint64_t k2, int k, uint64_t* hA, uint64_t* hB) {
  uint64_t e1 = fo_canon(k1, k), e2 = fo_canon(k2, k);
  uint64_t lo = e1 < e2 ? e1 : e2, hi = e1 < e2 ? e2 : e1;
  *hA = fo_old_hash(lo, 0, b->log2_tai);
  *hB = fo_old_hash(hi, 1, b->log2_tai);
}
static void bloom_add_pair(bloom_t* b, uint64_t k1, uint64_t k2, int k) {
  uint64_t hA, hB;
  pair_hashes(b, k1, k2, k, &hA, &hB);
  bloom_add(b, hA, hB);
}
static int bloom_contains_pair(const bloom_t* b, uint64_t k1, uint64_t k2, int k) {
  uint64_t hA, hB;
  pair_hashes(b, k1, k2, k, &hA, &hB);
  return bloom_contains(b, hA, hB);
}

/* Bloom::weight, utils/Bloom.cpp:191-203 (float division there) */
double fo_weight(const uint8_t* bits, int log2_tai) {
  uint64_t nchar = ((uint64_t)1 << log2_tai) / 8, w = 0;
  for (uint64_t i = 0; i < nchar; i++) w += (uint64_t)__builtin_popcount(bits[i]);
  return (double)((float)(long)w / (float)((uint64_t)1 << log2_tai));
}

/* Bloom::Bloom(tai_bloom,k) geometry, utils/Bloom.cpp:165-189 */
static int ctor_log2_tai(uint64_t requested) { return (int)log2((double)requested) + 1; }
/* set_number_of_hash_func, utils/Bloom.cpp:491-498: out-of-range keeps the constructor default 4 */
static int clamp_n_hash(int n) { return (n > 10 || n < 1) ? 4 : n; }

/* create_bloom_filter_optimal, utils/Bloom.cpp:229-247.  `log(fpRate)` on a float picks the float
 * overload; the two divisions by log(2) are double. */
void fo_geometry_optimal(uint64_t est, float fp, int* log2_tai, int* n_hash) {
  int bits_per_item = (int)((double)(-logf(fp)) / log(2.0) / log(2.0));
  uint64_t size = (uint64_t)(est * (uint64_t)(int64_t)bits_per_item);
  *log2_tai = ctor_log2_tai(size);
  *n_hash = clamp_n_hash((int)floorf((float)(0.7 * bits_per_item)));
}
/* create_bloom_filter_2_hash, utils/Bloom.cpp:206-226 */
void fo_geometry_2_hash(uint64_t est, float fp, int* log2_tai, int* n_hash) {
  int bits_per_item = 2 * (int)(1 / pow((double)fp, .5));
  uint64_t size = (uint64_t)(est * (uint64_t)(int64_t)bits_per_item);
  *log2_tai = ctor_log2_tai(size);
  *n_hash = 2;
}

/* my_func, src/Faucet.cpp:197-201 (estimated_kmers, singletons are uint64; fpRate is float) */
typedef struct { uint64_t est, sing; float fp; } myfunc_t;
static double my_func(const myfunc_t* a, double p1) {
  double c = ((double)a->est - (1 - p1) * (double)a->sing) / (double)a->est;
  return log(2.0) * (double)logf(a->fp) + log(p1) * log(1 - pow(2.0, -c));
}
/* brents_fun(f, fpRate, 0.5, 1e-4, 1000), utils/Bloom.cpp:33-124 -- including its `abs(fa) < abs(b)`
 * comparison against b (not fb) on line 48 */
double fo_brent_p1(uint64_t est, uint64_t singletons, float fp) {
  myfunc_t arg = {est, singletons, fp};
  double a = fp, b = 0.50, tol = 0.0001;
  double fa = my_func(&arg, a), fb = my_func(&arg, b), fs = 0;
  if (!(fa * fb < 0)) return -11;
  if (fabs(fa) < fabs(b)) { double t = a; a = b; b = t; t = fa; fa = fb; fb = t; }
  double c = a, fc = fa, s = 0, d = 0;
  int mflag = 1;
  for (unsigned iter = 1; iter < 1000; ++iter) {
    if (fabs(b - a) < tol) return s;
    if (fa != fc && fb != fc)
      s = (a * fb * fc / ((fa - fb) * (fa - fc))) + (b * fa * fc / ((fb - fa) * (fb - fc))) +
          (c * fa * fb / ((fc - fa) * (fc - fb)));
    else
      s = b - fb * (b - a) / (fb - fa);
    if (((s < (3 * a + b) * 0.25) || (s > b)) || (mflag && (fabs(s - b) >= (fabs(b - c) * 0.5))) ||
        (!mflag && (fabs(s - b) >= (fabs(c - d) * 0.5))) || (mflag && (fabs(b - c) < tol)) ||
        (!mflag && (fabs(c - d) < tol))) {
      s = (a + b) * 0.5;
      mflag = 1;
    } else {
      mflag = 0;
    }
    fs = my_func(&arg, s);
    d = c; c = b; fc = fb;
    if (fa * fs < 0) { b = s; fb = fs; } else { a = s; fa = fs; }
    if (fabs(fa) < fabs(fb)) { double t = a; a = b; b = t; t = fa; fa = fb; fb = t; }
  }
  return NAN; /* reference falls off the end (UB) */
}

/* ------------------------------------------------------------------ record reader */

/* std::getline(ifstream&, string&) state machine as the two passes use it
 * (utils/Bloom.cpp:280-282,340; src/ReadScanner.cpp:306-308,349): a failed sentry leaves the
 * destination string untouched, which is how an unterminated trailing header line ends up being
 * processed as its own sequence. */
typedef struct { const char* t; size_t n, pos; int eof, fail; } reader_t;
typedef struct { const char* p; size_t len; } str_t;
static int rd_getline(reader_t* r, str_t* line) {
  if (r->eof || r->fail) { r->fail = 1; return 0; }
  const char* s = r->t + r->pos;
  const char* nl = (const char*)memchr(s, '\n', r->n - r->pos);
  if (nl) {
    line->p = s; line->len = (size_t)(nl - s);
    r->pos += line->len + 1;
    return 1;
  }
  line->p = s; line->len = r->n - r->pos;
  r->pos = r->n; r->eof = 1;
  if (line->len == 0) { r->fail = 1; return 0; }
  return 1;
}

/* getUnambiguousReads, utils/Kmer.cpp:64-80: maximal ACGT runs of length >= k, LAST run first */
typedef struct { size_t start, len; } seg_t;
static size_t unambiguous_segments(str_t read, int k, seg_t** out, size_t* cap) {
  size_t n = 0, i = 0;
  while (i < read.len) {
    while (i < read.len && !valid_nt(read.p[i])) i++;
    size_t s = i;
    while (i < read.len && valid_nt(read.p[i])) i++;
    if (i - s >= (size_t)k) {
      if (n == *cap) { *cap = *cap ? *cap * 2 : 8; *out = (seg_t*)realloc(*out, *cap * sizeof(seg_t)); }
      (*out)[n].start = s; (*out)[n].len = i - s; n++;
    }
  }
  for (size_t a = 0; a + 1 < n - a; a++) { seg_t t = (*out)[a]; (*out)[a] = (*out)[n - 1 - a]; (*out)[n - 1 - a] = t; }
  return n;
}

/* ------------------------------------------------------------------ pass 1 */

/* load_two_filters, utils/Bloom.cpp:267-350, non-mercy branch :288-299 */
int fo_load_two_filters(const char* text, size_t n, int fastq, int k, int log2_tai, int n_hash,
                        uint8_t* bloo1, uint8_t* bloo2, fo_load_stats* stats) {
  bloom_t b1, b2;
  bloom_init(&b1, bloo1, log2_tai, n_hash);
  bloom_init(&b2, bloo2, log2_tai, n_hash);
  reader_t rd = {text, n, 0, 0, 0};
  str_t read = {text, 0}, skip;
  seg_t* segs = NULL; size_t cap = 0;
  uint64_t reads = 0, unamb = 0, kmers = 0;
  while (rd_getline(&rd, &read)) {
    rd_getline(&rd, &read);
    size_t ns = unambiguous_segments(read, k, &segs, &cap);
    for (size_t si = 0; si < ns; si++) {
      const char* s = read.p + segs[si].start;
      size_t len = segs[si].len;
      unamb++;
      uint64_t fwd, rc;
      fo_first_kmer(s, k, &fwd);
      rc = fo_revcomp(fwd, k);
      for (size_t pos = 0;; pos++) {
        uint64_t canon = fwd < rc ? fwd : rc;
        uint64_t hA = fo_old_hash(canon, 0, log2_tai), hB = fo_old_hash(canon, 1, log2_tai);
        if (bloom_contains(&b1, hA, hB)) bloom_add(&b2, hA, hB); else bloom_add(&b1, hA, hB);
        kmers++;
        if (pos + (size_t)k >= len) break;
        int nt = fo_nt2int(s[pos + (size_t)k]);
        fwd = ext_fwd(fwd, nt, k);             /* DoubleKmer::forward, utils/DoubleKmer.cpp:5-8 */
        rc = ext_bwd(rc, comp_nt(nt), k);
      }
    }
    reads++;
    if (fastq) { skip = read; rd_getline(&rd, &skip); rd_getline(&rd, &skip); read = skip; }
  }
  free(segs);
  if (stats) {
    stats->reads_processed = reads; stats->unambiguous_reads = unamb; stats->kmers = kmers;
    stats->weight1 = fo_weight(bloo1, log2_tai); stats->weight2 = fo_weight(bloo2, log2_tai);
  }
  return 0;
}

/* ------------------------------------------------------------------ junction map */

/* stand-in for std::unordered_map<kmer_type,Junction> (utils/JunctionMap.h:61): open addressing over
 * an array of records kept in creation order */
typedef struct {
  fo_junction_rec* recs; uint64_t n, cap;
  int64_t* slots; uint64_t nslots;
} jmap_t;
static uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; return x ^ (x >> 33); }
static void jmap_init(jmap_t* m) {
  m->cap = 1024; m->n = 0; m->recs = (fo_junction_rec*)malloc(m->cap * sizeof(fo_junction_rec));
  m->nslots = 4096; m->slots = (int64_t*)malloc(m->nslots * sizeof(int64_t));
  for (uint64_t i = 0; i < m->nslots; i++) m->slots[i] = -1;
}
static int64_t jmap_find(const jmap_t* m, uint64_t key) {
  uint64_t s = mix64(key) & (m->nslots - 1);
  while (m->slots[s] >= 0) {
    if (m->recs[m->slots[s]].kmer == key) return m->slots[s];
    s = (s + 1) & (m->nslots - 1);
  }
  return -1;
}
static void jmap_place(jmap_t* m, int64_t idx) {
  uint64_t s = mix64(m->recs[idx].kmer) & (m->nslots - 1);
  while (m->slots[s] >= 0) s = (s + 1) & (m->nslots - 1);
  m->slots[s] = idx;
}
/* JunctionMap::createJunction, utils/JunctionMap.cpp:567-570 (insert: no-op when present) */
static int64_t jmap_create(jmap_t* m, uint64_t key) {
  int64_t f = jmap_find(m, key);
  if (f >= 0) return f;
  if (m->n == m->cap) { m->cap *= 2; m->recs = (fo_junction_rec*)realloc(m->recs, m->cap * sizeof(fo_junction_rec)); }
  memset(&m->recs[m->n], 0, sizeof(fo_junction_rec));
  m->recs[m->n].kmer = key;
  TRACE_WRITE(key, 0);
  if ((m->n + 1) * 2 > m->nslots) {
    m->nslots *= 2; m->slots = (int64_t*)realloc(m->slots, m->nslots * sizeof(int64_t));
    for (uint64_t i = 0; i < m->nslots; i++) m->slots[i] = -1;
    for (uint64_t i = 0; i < m->n; i++) jmap_place(m, (int64_t)i);
  }
  jmap_place(m, (int64_t)m->n);
  return (int64_t)m->n++;
}
/* Junction::update, utils/Junction.cpp:69-71; the int argument is narrowed to unsigned char at the call */
static void junc_update(fo_junction_rec* r, int idx, int length) {
  uint8_t l = (uint8_t)length;
  if (l > r->dist[idx]) { r->dist[idx] = l; TRACE_WRITE(r->kmer, 1); }
}
/* Junction::addCoverage, utils/Junction.cpp:59-67 */
static void junc_add_cov(fo_junction_rec* r, int nt) {
  r->cov[nt] = (uint8_t)(r->cov[nt] + 1);
  if (r->cov[nt] == 0) r->cov[nt] = 255;
}

/* ------------------------------------------------------------------ ReadKmer cursor (utils/ReadKmer.cpp) */

typedef struct {
  const char* read; int len; int k;
  uint64_t fwd, rc;
  int pos; int dir; /* 1 = FORWARD, 0 = BACKWARD */
} cursor_t;
static void cur_init(cursor_t* c, const char* read, int len, int k, int index, int dir) { /* :116-134 */
  c->read = read; c->len = len; c->k = k; c->pos = index; c->dir = dir;
  fo_first_kmer(read + index, k, &c->fwd);
  c->rc = fo_revcomp(c->fwd, k);
}
static int cur_total_pos(const cursor_t* c) { return 2 * c->pos + (c->dir ? 1 : 0); }             /* :33-35 */
static int cur_dist_to_end(const cursor_t* c) { return c->len * 2 - cur_total_pos(c) - 2 * c->k + 1; } /* :28-30 */
static uint64_t cur_kmer(const cursor_t* c) { return c->dir ? c->fwd : c->rc; }                   /* :50-57 */
static void cur_forward(cursor_t* c) {                                                            /* :72-83 */
  c->dir = !c->dir;
  if (c->dir) return;
  int nt = 0;
  if (c->pos + c->k < c->len) nt = fo_nt2int(c->read[c->pos + c->k]);
  c->fwd = ext_fwd(c->fwd, nt, c->k);
  c->rc = ext_bwd(c->rc, comp_nt(nt), c->k);
  c->pos++;
}
static int cur_real_nuc(const cursor_t* c) {                                                      /* :107-114 */
  if (c->dir) return fo_nt2int(c->pos + c->k < c->len ? c->read[c->k + c->pos] : '\0');
  return comp_nt(fo_nt2int(c->read[c->pos - 1]));
}
static uint64_t cur_extension(const cursor_t* c, int nt) { return ext_fwd(c->dir ? c->fwd : c->rc, nt, c->k); } /* :102-104, DoubleKmer.cpp:10-17 */
static uint64_t cur_real_extension(const cursor_t* c) { return cur_extension(c, cur_real_nuc(c)); }
static int cur_ext_index(const cursor_t* c, int dir) { return dir != c->dir ? 4 : cur_real_nuc(c); } /* :95-100 */

/* ------------------------------------------------------------------ pass 2 */

typedef struct {
  int k, j, max_spacer, no_cleaning;
  bloom_t bloom; bloom_t* spf; bloom_t* lpf;
  jmap_t map;
  fo_scan_stats st;
} scanner_t;

typedef struct { uint64_t* v; size_t n, cap; } klist_t;
static void kl_push(klist_t* l, uint64_t x) {
  if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 16; l->v = (uint64_t*)realloc(l->v, l->cap * sizeof(uint64_t)); }
  l->v[l->n++] = x;
}

/* JChecker::jcheck(kmer_type), utils/JChecker.cpp:51-80 */
static int jcheck(const scanner_t* s, uint64_t kmer) {
  uint64_t bufA[1024], bufB[1024];
  uint64_t *last = bufA, *next = bufB;
  int last_n = 1;
  last[0] = kmer;
  for (int lvl = 0; lvl < s->j; lvl++) {
    int next_n = 0;
    for (int a = 0; a < last_n; a++)
      for (int nt = 0; nt < 4; nt++) {
        uint64_t nk = ext_fwd(last[a], nt, s->k);
        if (bloom_old_contains(&s->bloom, fo_canon(nk, s->k))) next[next_n++] = nk;
      }
    if (next_n == 0) return 0;
    last_n = next_n;
    uint64_t* t = last; last = next; next = t;
  }
  return 1;
}

/* ReadScanner::testForJunction, src/ReadScanner.cpp:36-56 */
static int test_for_junction(scanner_t* s, const cursor_t* c) {
  uint64_t real_ext = cur_real_extension(c);
  for (int nt = 0; nt < 4; nt++) {
    uint64_t t = cur_extension(c, nt);
    if (t != real_ext && bloom_old_contains(&s->bloom, fo_canon(t, s->k))) {
      s->st.nb_jcheck_kmer++;
      if (jcheck(s, t)) return 1;
    }
  }
  return 0;
}

/* ReadScanner::find_next_junction, src/ReadScanner.cpp:61-86 */
static int find_next_junction(scanner_t* s, cursor_t* c, int last_junc_pos) {
  for (; cur_dist_to_end(c) > 2 * s->j; cur_forward(c)) {
    if (jmap_find(&s->map, cur_kmer(c)) >= 0) return 1;
    if (cur_total_pos(c) - last_junc_pos >= 2 * s->max_spacer - 1) return 1;
    if (test_for_junction(s, c)) return 1;
    s->st.nb_processed++;
  }
  return 0;
}

/* ReadScanner::add_fake_junction, src/ReadScanner.cpp:92-104 */
static uint64_t add_fake_junction(scanner_t* s, const char* read, int len) {
  cursor_t m;
  cur_init(&m, read, len, s->k, len / 2 - s->k / 2, 1);
  uint64_t ext = cur_real_extension(&m);
  const int64_t ji = jmap_create(&s->map, cur_kmer(&m)); /* may realloc map.recs: index first, pointer after */
  fo_junction_rec* r = &s->map.recs[ji];
  junc_add_cov(r, cur_real_nuc(&m));
  junc_update(r, cur_ext_index(&m, 0), cur_total_pos(&m) - 2 * s->j);
  junc_update(r, cur_ext_index(&m, 1), cur_dist_to_end(&m) - 2 * s->j);
  return ext;
}

/* ReadScanner::scan_forward, src/ReadScanner.cpp:112-231 */
static void scan_forward(scanner_t* s, const char* read, int len, klist_t* out) {
  size_t first = out->n;
  cursor_t c;
  cur_init(&c, read, len, s->k, 0, 0);
  for (int i = 0; i < 2 * s->j + 1; i++) cur_forward(&c);
  int have_last = 0, have_first_back = 0, have_last_fwd = 0;
  cursor_t last_kmer, first_back, last_fwd;
  int64_t last_junc = -1;
  int rev_pos = 0, for_pos = 0, last_junc_pos = 0;
  memset(&last_kmer, 0, sizeof last_kmer); memset(&first_back, 0, sizeof first_back); memset(&last_fwd, 0, sizeof last_fwd);
  while (find_next_junction(s, &c, last_junc_pos)) {
    int64_t ji = jmap_find(&s->map, cur_kmer(&c));
    last_junc_pos = cur_total_pos(&c);
    if (ji < 0) ji = jmap_create(&s->map, cur_kmer(&c));
    kl_push(out, cur_real_extension(&c));
    if (!c.dir) {
      if (!have_first_back) { have_first_back = 1; first_back = c; rev_pos = c.pos; }
    } else {
      if (!have_last_fwd) { have_last_fwd = 1; for_pos = c.pos; }
      last_fwd = c;
    }
    fo_junction_rec* junc = &s->map.recs[ji];
    junc_add_cov(junc, cur_real_nuc(&c));
    if (have_last) { /* directLinkJunctions, utils/JunctionMap.cpp:551-561 */
      fo_junction_rec* prev = &s->map.recs[last_junc];
      int e1 = cur_ext_index(&last_kmer, 1), e2 = cur_ext_index(&c, 0);
      int d = cur_total_pos(&c) - cur_total_pos(&last_kmer);
      junc_update(prev, e1, d); junc_update(junc, e2, d);
      if (!prev->linked[e1]) TRACE_WRITE(prev->kmer, 2);
      if (!junc->linked[e2]) TRACE_WRITE(junc->kmer, 2);
      prev->linked[e1] = 1; junc->linked[e2] = 1;
    } else {
      have_last = 1;
      junc_update(junc, cur_ext_index(&c, 0), cur_total_pos(&c) - 2 * s->j);
    }
    last_kmer = c; last_junc = ji;
    int dist = junc->dist[cur_ext_index(&c, 1)];
    if (dist < 1) dist = 1;
    for (int i = 0; i < dist; i++) cur_forward(&c);
    s->st.nb_processed++; s->st.nb_skipped += (uint64_t)(dist - 1);
  }
  if (!have_last) {
    s->st.nb_no_juncs++;
    kl_push(out, add_fake_junction(s, read, len));
  } else {
    junc_update(&s->map.recs[last_junc], cur_ext_index(&last_kmer, 1), cur_dist_to_end(&last_kmer) - 2 * s->j);
  }
  if (!s->no_cleaning && s->spf) {
    size_t cnt = out->n - first;
    uint64_t* v = out->v + first;
    if (cnt == 2) {
      if (have_first_back && have_last_fwd && !(rev_pos > for_pos))
        bloom_add_pair(s->spf, cur_real_extension(&first_back), cur_real_extension(&last_fwd), s->k);
      if ((have_first_back && !have_last_fwd) || (!have_first_back && have_last_fwd))
        bloom_add_pair(s->spf, v[0], v[1], s->k);
    } else if (cnt > 2) {
      for (size_t i = 0; i + 2 < cnt; i++) bloom_add_pair(s->spf, v[i], v[i + 2], s->k);
    }
  }
}

/* ReadScanner::getValidReads (src/ReadScanner.cpp:233-257) + scanInputRead (:260-282) */
static void scan_input_read(scanner_t* s, str_t read, klist_t* out, seg_t** segs, size_t* cap) {
  int k = s->k;
  size_t ns = unambiguous_segments(read, k, segs, cap);
  for (size_t si = 0; si < ns; si++) {
    const char* seg = read.p + (*segs)[si].start;
    int len = (int)(*segs)[si].len;
    if (len < k + 2 * s->j + 1) continue;
    s->st.unambiguous_reads++;
    int start = 0, end = 0;
    uint64_t fwd, rc;
    fo_first_kmer(seg, k, &fwd);
    rc = fo_revcomp(fwd, k);
    for (int pos = 0; pos + k <= len; pos++) {
      if (pos > 0) {
        int nt = fo_nt2int(seg[pos + k - 1]);
        fwd = ext_fwd(fwd, nt, k); rc = ext_bwd(rc, comp_nt(nt), k);
      }
      if (bloom_old_contains(&s->bloom, fwd < rc ? fwd : rc)) {
        end++;
      } else {
        if (end >= start + k) { scan_forward(s, seg + start, end - start + k - 1, out); s->st.reads_no_errors++; }
        start = pos + 1; end = pos + 1;
      }
    }
    if (end >= start + k) { scan_forward(s, seg + start, end - start + k - 1, out); s->st.reads_no_errors++; }
  }
}

static void scanner_init(scanner_t* s, int k, int j, int max_spacer, int no_cleaning) {
  memset(s, 0, sizeof *s);
  s->k = k; s->j = j; s->max_spacer = max_spacer; s->no_cleaning = no_cleaning;
  jmap_init(&s->map);
}
static void scanner_finish(scanner_t* s, fo_junction_rec** recs_out, uint64_t* n_recs_out, fo_scan_stats* stats) {
  s->st.n_junctions = s->map.n;
  if (stats) *stats = s->st;
  if (recs_out) { *recs_out = s->map.recs; *n_recs_out = s->map.n; } else free(s->map.recs);
  free(s->map.slots);
}

/* ReadScanner::scanReads, src/ReadScanner.cpp:284-359 */
int fo_scan(const char* text, size_t n, int fastq, int paired_ends, int no_cleaning, int k, int j,
            int max_spacer_dist, const uint8_t* bloo2, int log2_tai, int n_hash, uint8_t* short_pf,
            int spf_log2_tai, int spf_n_hash, uint8_t* long_pf, int lpf_log2_tai, int lpf_n_hash,
            const uint64_t* fake_set, size_t n_fake, fo_junction_rec** recs_out, uint64_t* n_recs_out,
            fo_scan_stats* stats) {
  scanner_t s;
  scanner_init(&s, k, j, max_spacer_dist, no_cleaning);
  bloom_init(&s.bloom, (uint8_t*)bloo2, log2_tai, n_hash);
  s.bloom.fake = fake_set; s.bloom.n_fake = n_fake;
  bloom_t spf, lpf;
  if (short_pf) { bloom_init(&spf, short_pf, spf_log2_tai, spf_n_hash); s.spf = &spf; }
  if (long_pf) { bloom_init(&lpf, long_pf, lpf_log2_tai, lpf_n_hash); s.lpf = &lpf; }
  reader_t rd = {text, n, 0, 0, 0};
  str_t read = {text, 0}, skip;
  seg_t* segs = NULL; size_t cap = 0;
  klist_t b1 = {0, 0, 0}, b2 = {0, 0, 0};
  int first_end = 1;
  while (rd_getline(&rd, &read)) {
    rd_getline(&rd, &read);
    klist_t* cur = first_end ? &b1 : &b2;
    cur->n = 0;
    TRACE_RECORD();
    scan_input_read(&s, read, cur, &segs, &cap);
    if (paired_ends && !first_end && b1.n && b2.n && !no_cleaning && s.lpf) {
      for (size_t a = 0; a < b1.n; a++) {
        int paired = 0;
        for (size_t b = 0; b < b2.n; b++)
          if (bloom_contains_pair(s.lpf, b1.v[a], b2.v[b], k)) { paired = 1; break; }
        if (!paired) bloom_add_pair(s.lpf, b1.v[a], b2.v[0], k);
      }
    }
    s.st.reads_processed++;
    if (fastq) { skip = read; rd_getline(&rd, &skip); rd_getline(&rd, &skip); read = skip; }
    first_end = !first_end;
  }
  free(segs); free(b1.v); free(b2.v);
  scanner_finish(&s, recs_out, n_recs_out, stats);
  return 0;
}

int fo_scan_reads(const char* const* reads, int n_reads, int k, int j, int max_spacer_dist,
                  const uint64_t* fake_set, size_t n_fake, fo_junction_rec** recs_out,
                  uint64_t* n_recs_out, fo_scan_stats* stats) {
  scanner_t s;
  scanner_init(&s, k, j, max_spacer_dist, 1);
  bloom_init(&s.bloom, NULL, 10, 4);
  s.bloom.fake = fake_set; s.bloom.n_fake = n_fake;
  seg_t* segs = NULL; size_t cap = 0;
  klist_t l = {0, 0, 0};
  for (int i = 0; i < n_reads; i++) {
    str_t r = {reads[i], strlen(reads[i])};
    l.n = 0;
    scan_input_read(&s, r, &l, &segs, &cap);
    s.st.reads_processed++;
  }
  free(segs); free(l.v);
  scanner_finish(&s, recs_out, n_recs_out, stats);
  return 0;
}

/* Junction::toString (utils/Junction.cpp:74-89) after print_kmer (utils/JunctionMap.cpp:589-592) */
int fo_junction_line(const fo_junction_rec* r, int k, char* out, size_t cap) {
  char km[40];
  fo_kmer_string(r->kmer, k, km);
  int csum = r->cov[0] + r->cov[1] + r->cov[2] + r->cov[3];
  return snprintf(out, cap, "%s %d %d %d %d %d  %d %d %d %d %d  %d %d %d %d %d \n", km, r->dist[0], r->dist[1],
                  r->dist[2], r->dist[3], r->dist[4], r->cov[0], r->cov[1], r->cov[2], r->cov[3], csum,
                  r->linked[0], r->linked[1], r->linked[2], r->linked[3], r->linked[4]);
}

char* fo_read_file(const char* path, size_t* n) {
  FILE* f = fopen(path, "rb");
  if (!f) { *n = 0; return NULL; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  char* buf = (char*)malloc((size_t)sz + 1);
  *n = fread(buf, 1, (size_t)sz, f);
  buf[*n] = 0;
  fclose(f);
  return buf;
}
void fo_free(void* p) { free(p); }
