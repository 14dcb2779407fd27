// Host side of pass 2's long pair filter (src/ReadScanner.cpp:317-343).
//
// The filter is read-modify-write per MATE PAIR in stream order (containsPair on what earlier pairs
// added, then addPair), so it is inherently sequential; it is also tiny: one step per read pair, over
// the handful of junction extensions each mate produced.  The GPU stitch emits those extension lists
// (chunks of the ext buffer, stitch.cuh); this class replays the filter logic over them in record
// order.  No Bloom query of bloo2 and no junction-map access happens here.
#pragma once
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "kmer.cuh"

namespace faucet {

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

class LongPairFilter {
 public:
  void init(uint8_t* bits, int log2_tai, int n_hash, int k) {
    lpf_.bits = bits; lpf_.mask = (1ull << log2_tai) - 1; lpf_.n_hash = n_hash; k_ = k;
    carry_.clear(); have_carry_ = false;
  }
  bool enabled() const { return lpf_.bits != nullptr; }

  // ext[0..n_words): chunks {header = record:32 | part:16 | count:16, count k-mers} in any order, for the
  // records [0, n_recs) of one batch whose first record has global index rec_base.
  void process_batch(const uint64_t* ext, uint64_t n_words, uint32_t n_recs, uint64_t rec_base) {
    first_.assign(n_recs, ~0ull);
    more_.clear();
    for (uint64_t off = 0; off < n_words;) {
      const uint64_t h = ext[off];
      const uint32_t rec = (uint32_t)(h >> 32), part = (uint32_t)(h >> 16) & 0xffffu, cnt = (uint32_t)h & 0xffffu;
      if (part == 0) first_[rec] = off; else more_[((uint64_t)rec << 16) | part] = off;
      off += 1 + cnt;
    }
    // Every extension takes part in up to |other list| pair tests; its canonical form and both hashes are computed once
    // (addPair / containsPair hash the smaller canonical k-mer with seed 0 and the larger with seed 1).
    std::vector<Ext> cur;
    for (uint32_t r = 0; r < n_recs; r++) {
      cur.clear();
      for (uint32_t part = 0;; part++) {
        uint64_t off;
        if (part == 0) { off = first_[r]; if (off == ~0ull) break; }
        else { auto it = more_.find(((uint64_t)r << 16) | part); if (it == more_.end()) break; off = it->second; }
        const uint32_t cnt = (uint32_t)ext[off] & 0xffffu;
        for (uint32_t i = 0; i < cnt; i++) {
          const uint64_t x = ext[off + 1 + i];
          Ext e;
          e.canon = canon(x, revcomp(x, k_));
          e.h0 = hash0(e.canon) & lpf_.mask;
          e.h1 = hash1(e.canon) & lpf_.mask;
          cur.push_back(e);
        }
        if (more_.empty()) break;
      }
      const bool first_end = ((rec_base + r) & 1ull) == 0;
      if (first_end) { carry_.swap(cur); have_carry_ = true; continue; }
      if (have_carry_ && !carry_.empty() && !cur.empty()) {
        for (const Ext& p1 : carry_) {  // src/ReadScanner.cpp:322-339
          bool found = false;
          for (const Ext& p2 : cur) {
            const bool lt = p1.canon < p2.canon;
            if (lpf_.contains(lt ? p1.h0 : p2.h0, lt ? p2.h1 : p1.h1)) { found = true; break; }
          }
          if (!found) {
            const Ext& p2 = cur.front();
            const bool lt = p1.canon < p2.canon;
            lpf_.add(lt ? p1.h0 : p2.h0, lt ? p2.h1 : p1.h1);
          }
        }
      }
      have_carry_ = false;
    }
  }

 private:
  HostBloom lpf_;
  int k_ = 0;
  struct Ext { uint64_t canon, h0, h1; };  // an extension k-mer: canonical form, seed-0 and seed-1 hash (masked)
  std::vector<Ext> carry_;       // mate-1 list waiting for its mate (may span a batch boundary)
  bool have_carry_ = false;
  std::vector<uint64_t> first_;
  std::unordered_map<uint64_t, uint64_t> more_;
};

}  // namespace faucet
