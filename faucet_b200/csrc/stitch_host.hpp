// Pass 2, stream-order part, host edition ("stitch v1").
//
// Replays ReadScanner::scanReads / scanInputRead / scan_forward / find_next_junction
// (src/ReadScanner.cpp:61-359) over the flag bytes produced by scan_flags_kernel.  No Bloom query of
// bloo2 happens here: validity and testForJunction answers are table lookups; what remains is the
// strictly sequential junction-map bookkeeping (create / coverage / link / dist / skip) and the two
// pair filters.  Records keep their creation rank so the caller can rebuild the reference's
// std::unordered_map in insertion order.
#pragma once
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../../include/faucet_gpu.h"
#include "kmer.cuh"

namespace faucet {

// a small host Bloom filter: only used for the pair filters (src/Faucet.cpp:265-281)
struct HostBloom {
  uint8_t* bits = nullptr;
  uint64_t mask = 0;
  int n_hash = 0;
  void add(uint64_t h0, uint64_t h1) {  // Bloom::add, utils/Bloom.h:217-226
    uint64_t h = h0;
    for (int i = 0; i < n_hash; i++, h += h1) { h &= mask; bits[h >> 3] |= (uint8_t)(1u << (h & 7)); }
  }
  bool contains(uint64_t h0, uint64_t h1) const {  // Bloom::contains, utils/Bloom.h:242-258
    uint64_t h = h0 & mask;
    for (int i = 0; i < n_hash; i++, h = (h + h1) & mask)
      if (!(bits[h >> 3] & (1u << (h & 7)))) return false;
    return true;
  }
  // addPair / containsPair hash the smaller canonical k-mer with seed 0 and the larger with seed 1
  // (utils/Bloom.cpp:127-154)
  void pair_hashes(uint64_t k1, uint64_t k2, int k, uint64_t* hA, uint64_t* hB) const {
    uint64_t e1 = canon(k1, revcomp(k1, k)), e2 = canon(k2, revcomp(k2, k));
    *hA = hash0(e1 < e2 ? e1 : e2) & mask;
    *hB = hash1(e1 < e2 ? e2 : e1) & mask;
  }
  void add_pair(uint64_t k1, uint64_t k2, int k) { uint64_t a, b; pair_hashes(k1, k2, k, &a, &b); add(a, b); }
  bool contains_pair(uint64_t k1, uint64_t k2, int k) const { uint64_t a, b; pair_hashes(k1, k2, k, &a, &b); return contains(a, b); }
};

class HostStitch {
 public:
  HostStitch(int k, int j, int max_spacer, bool paired, bool no_cleaning)
      : k_(k), j_(j), spacer_(max_spacer), paired_(paired), no_cleaning_(no_cleaning), mask_(kmer_mask(k)) {
    std::memset(&st_, 0, sizeof st_);
  }
  void set_pair_filters(uint8_t* spf, int spf_log2, int spf_nh, uint8_t* lpf, int lpf_log2, int lpf_nh) {
    if (spf) { spf_.bits = spf; spf_.mask = (1ull << spf_log2) - 1; spf_.n_hash = spf_nh; }
    if (lpf) { lpf_.bits = lpf; lpf_.mask = (1ull << lpf_log2) - 1; lpf_.n_hash = lpf_nh; }
  }

  // text[0..n): a whole number of records (or the tail of the file when final); flags[p] as written
  // by scan_flags_kernel for the same byte offsets.
  void process(const uint8_t* text, size_t n, const uint8_t* flags, bool fastq) {
    size_t pos = 0;
    bool eof = false;
    const uint8_t* line;
    size_t len;
    // while(getline(header)) { getline(seq); ...; if fastq: 2 more getline }   (src/ReadScanner.cpp:306-350)
    while (!eof && next_line(text, n, &pos, &eof, &line, &len)) {
      const uint8_t* seq = line;  // failed sentry => the header text itself is scanned
      size_t seq_len = len;
      if (!eof) { if (!next_line(text, n, &pos, &eof, &seq, &seq_len)) { seq_len = 0; } }
      std::vector<uint64_t>& out = first_end_ ? back1_ : back2_;
      out.clear();
      scan_input_read(text, seq, seq_len, flags, out);
      if (paired_ && !first_end_ && !back1_.empty() && !back2_.empty() && !no_cleaning_ && lpf_.bits) {
        for (uint64_t p1 : back1_) {  // src/ReadScanner.cpp:317-343
          bool found = false;
          for (uint64_t p2 : back2_)
            if (lpf_.contains_pair(p1, p2, k_)) { found = true; break; }
          if (!found) lpf_.add_pair(p1, back2_.front(), k_);
        }
      }
      st_.reads_processed++;
      if (fastq) {
        const uint8_t* d; size_t dl;
        if (!eof) next_line(text, n, &pos, &eof, &d, &dl);
        if (!eof) next_line(text, n, &pos, &eof, &d, &dl);
      }
      first_end_ = !first_end_;
    }
  }

  std::vector<faucet_junction_rec>& records() { return recs_; }
  faucet_scan_stats stats() const { faucet_scan_stats s = st_; s.n_junctions = recs_.size(); return s; }

 private:
  // std::getline semantics: returns false (and sets eof) when nothing could be extracted
  static bool next_line(const uint8_t* t, size_t n, size_t* pos, bool* eof, const uint8_t** line, size_t* len) {
    if (*pos >= n) { *eof = true; return false; }
    const void* nl = std::memchr(t + *pos, '\n', n - *pos);
    *line = t + *pos;
    if (nl) { *len = (const uint8_t*)nl - (t + *pos); *pos += *len + 1; return true; }
    *len = n - *pos; *pos = n; *eof = true;
    return true;
  }

  uint32_t junction(uint64_t key) {  // getJunction-or-createJunction (utils/JunctionMap.cpp:533-570)
    auto it = index_.find(key);
    if (it != index_.end()) return it->second;
    faucet_junction_rec r;
    std::memset(&r, 0, sizeof r);
    r.kmer = key;
    r.creation_rank = recs_.size();
    recs_.push_back(r);
    index_.emplace(key, (uint32_t)(recs_.size() - 1));
    return (uint32_t)(recs_.size() - 1);
  }
  static void update(faucet_junction_rec& r, int idx, int length) {  // Junction::update (u8 narrowing at the call)
    uint8_t l = (uint8_t)length;
    if (l > r.dist[idx]) r.dist[idx] = l;
  }
  static void add_cov(faucet_junction_rec& r, int nt) {  // Junction::addCoverage, saturating
    if (r.cov[nt] != 255) r.cov[nt]++;
  }

  uint64_t fwd_at(const uint8_t* s, int pos) const {
    uint64_t x = 0;
    for (int i = 0; i < k_; i++) x = (x << 2) | nt_code(s[pos + i]);
    return x & mask_;
  }

  struct Cursor { int tp; uint64_t fwd; };  // half-step index and the forward k-mer at tp>>1

  // scan_forward (src/ReadScanner.cpp:112-231) on the valid sub-read s[0..len)
  void scan_forward(const uint8_t* s, int len, const uint8_t* fl, std::vector<uint64_t>& out) {
    const size_t first = out.size();
    const int tested_end = 2 * len - 2 * k_ + 1 - 2 * j_;  // distToEnd > 2j  <=>  tp < tested_end
    int tp = 2 * j_ + 1, cur_pos = -1, last_junc_pos = 0;
    uint64_t fwd = 0;
    bool have_last = false, have_fb = false, have_lf = false;
    int last_tp = 0, last_fwd_idx = 0, rev_pos = 0, for_pos = 0;
    uint32_t last_rec = 0;
    uint64_t fb_ext = 0, lf_ext = 0;
    while (true) {
      bool found = false;
      int pos = 0, dir = 0;
      uint64_t key = 0;
      for (; tp < tested_end; tp++) {  // find_next_junction, :61-86
        pos = tp >> 1; dir = tp & 1;
        if (pos != cur_pos) {  // roll or re-seed the forward k-mer
          if (cur_pos >= 0 && pos > cur_pos && pos - cur_pos < k_)
            for (int q = cur_pos + 1; q <= pos; q++) fwd = ext_fwd(fwd, nt_code(s[q + k_ - 1]), mask_);
          else
            fwd = fwd_at(s, pos);
          cur_pos = pos;
        }
        key = dir ? fwd : revcomp(fwd, k_);
        if (index_.find(key) != index_.end()) { found = true; break; }
        if (tp - last_junc_pos >= 2 * spacer_ - 1) { found = true; break; }
        uint8_t f = fl[pos];
        st_.nb_jcheck_kmer += dir ? ((f >> 3) & 3) : ((f >> 5) & 3);
        if (dir ? (f & 2) : (f & 4)) { found = true; break; }
        st_.nb_processed++;
      }
      if (!found) break;
      uint32_t ri = junction(key);
      last_junc_pos = tp;
      // real extension: next read base when facing forward, complement of the previous base when
      // facing backward (utils/ReadKmer.cpp:102-114)
      int real = dir ? (int)nt_code(s[pos + k_]) : (int)nt_comp(nt_code(s[pos - 1]));
      uint64_t real_ext = ext_fwd(key, (uint32_t)real, mask_);
      out.push_back(real_ext);
      if (!dir) { if (!have_fb) { have_fb = true; fb_ext = real_ext; rev_pos = pos; } }
      else { if (!have_lf) { have_lf = true; for_pos = pos; } lf_ext = real_ext; }
      add_cov(recs_[ri], real);
      const int fwd_idx = dir ? real : 4, back_idx = dir ? 4 : real;  // getExtensionIndex, :95-100
      if (have_last) {  // directLinkJunctions, utils/JunctionMap.cpp:551-561
        int d = tp - last_tp;
        update(recs_[last_rec], last_fwd_idx, d); update(recs_[ri], back_idx, d);
        recs_[last_rec].linked[last_fwd_idx] = 1; recs_[ri].linked[back_idx] = 1;
      } else {
        have_last = true;
        update(recs_[ri], back_idx, tp - 2 * j_);
      }
      last_tp = tp; last_rec = ri; last_fwd_idx = fwd_idx;
      int dist = recs_[ri].dist[fwd_idx];
      if (dist < 1) dist = 1;
      tp += dist;
      st_.nb_processed++; st_.nb_skipped += (uint64_t)(dist - 1);
    }
    if (!have_last) {  // add_fake_junction, :92-104
      st_.nb_no_juncs++;
      int pos = len / 2 - k_ / 2;
      uint64_t key = fwd_at(s, pos);
      int real = (int)nt_code(s[pos + k_]);
      uint32_t ri = junction(key);
      add_cov(recs_[ri], real);
      int mtp = 2 * pos + 1;
      update(recs_[ri], 4, mtp - 2 * j_);
      update(recs_[ri], real, (2 * len - mtp - 2 * k_ + 1) - 2 * j_);
      out.push_back(ext_fwd(key, (uint32_t)real, mask_));
    } else {
      update(recs_[last_rec], last_fwd_idx, (2 * len - last_tp - 2 * k_ + 1) - 2 * j_);
    }
    if (!no_cleaning_ && spf_.bits) {  // :207-225
      size_t cnt = out.size() - first;
      const uint64_t* v = out.data() + first;
      if (cnt == 2) {
        if (have_fb && have_lf && !(rev_pos > for_pos)) spf_.add_pair(fb_ext, lf_ext, k_);
        if (have_fb != have_lf) spf_.add_pair(v[0], v[1], k_);
      } else if (cnt > 2) {
        for (size_t i = 0; i + 2 < cnt; i++) spf_.add_pair(v[i], v[i + 2], k_);
      }
    }
  }

  // scanInputRead (:260-282) + getValidReads (:233-257) over one sequence line
  void scan_input_read(const uint8_t* text, const uint8_t* seq, size_t seq_len, const uint8_t* flags,
                       std::vector<uint64_t>& out) {
    segs_.clear();
    size_t i = 0;
    while (i < seq_len) {  // getUnambiguousReads, utils/Kmer.cpp:64-80
      while (i < seq_len && !nt_valid(seq[i])) i++;
      size_t s = i;
      while (i < seq_len && nt_valid(seq[i])) i++;
      if (i - s >= (size_t)k_) segs_.push_back({s, i - s});
    }
    for (size_t si = segs_.size(); si-- > 0;) {  // last segment first
      const uint8_t* s = seq + segs_[si].first;
      const int len = (int)segs_[si].second;
      if (len < k_ + 2 * j_ + 1) continue;
      st_.unambiguous_reads++;
      const uint8_t* fl = flags + (s - text);
      int start = 0, end = 0;
      for (int pos = 0; pos + k_ <= len; pos++) {
        if (fl[pos] & 1) { end++; continue; }
        if (end >= start + k_) { scan_forward(s + start, end - start + k_ - 1, fl + start, out); st_.reads_no_errors++; }
        start = end = pos + 1;
      }
      if (end >= start + k_) { scan_forward(s + start, end - start + k_ - 1, fl + start, out); st_.reads_no_errors++; }
    }
  }

  int k_, j_, spacer_;
  bool paired_, no_cleaning_;
  uint64_t mask_;
  bool first_end_ = true;
  HostBloom spf_, lpf_;
  std::unordered_map<uint64_t, uint32_t> index_;
  std::vector<faucet_junction_rec> recs_;
  std::vector<uint64_t> back1_, back2_;
  std::vector<std::pair<size_t, size_t>> segs_;
  faucet_scan_stats st_;
};

}  // namespace faucet
