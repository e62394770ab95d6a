// Host-side restatement of CPython's `random.sample(range(n), k)` — the candidate draw of PyG 2.2.0's
// `negative_sampling` (`torch_geometric/utils/negative_sampling.py: sample()` -> `random.sample(range(population), k)`),
// reached from the reference at src/train_teacher_gnn.py:50-51 and src/main.py:81-82,206-207.  Bit-exact sampled
// indices need CPython's exact algorithm: MT19937 (`_randommodule.c`), `getrandbits`,
// `_randbelow_with_getrandbits` (rejection from the next power of two) and `Random.sample`'s two selection schemes
// (partial shuffle of a pool for small populations, rejection against a set otherwise).  The Python loop costs ~0.4 us
// per draw — 4-30 ms per training step for 10^4-10^5 candidates, more than the whole GPU step; this is the same stream
// in C++.  No CUDA here: the function runs on any host (tests/test_host_logic.py pins it against `random.sample`).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/llp_b200.h"

namespace {

constexpr int kN = 624, kM = 397;

struct MT {
  uint32_t s[kN];  // private copy of the 624 state words (the caller's block is written back once, at the end)
  uint32_t pos;    // index (0..624)
  void twist() {
    int kk;
    uint32_t y;
    for (kk = 0; kk < kN - kM; kk++) {
      y = (s[kk] & 0x80000000u) | (s[kk + 1] & 0x7fffffffu);
      s[kk] = s[kk + kM] ^ (y >> 1) ^ ((y & 0x1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < kN - 1; kk++) {
      y = (s[kk] & 0x80000000u) | (s[kk + 1] & 0x7fffffffu);
      s[kk] = s[kk + (kM - kN)] ^ (y >> 1) ^ ((y & 0x1u) ? 0x9908b0dfu : 0u);
    }
    y = (s[kN - 1] & 0x80000000u) | (s[0] & 0x7fffffffu);
    s[kN - 1] = s[kM - 1] ^ (y >> 1) ^ ((y & 0x1u) ? 0x9908b0dfu : 0u);
    pos = 0;
  }
  static uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  uint32_t next() {
    if (pos >= (uint32_t)kN) twist();
    return temper(s[pos++]);
  }
  // _random.Random.getrandbits(k), 1 <= k <= 64: words are filled little-endian, the LAST word keeps its top bits
  uint64_t getrandbits(int k) {
    if (k <= 32) return (uint64_t)(next() >> (32 - k));
    const uint64_t lo = next();
    const uint64_t hi = next() >> (32 - (k - 32));
    return lo | (hi << 32);
  }
  // Random._randbelow_with_getrandbits(n)
  uint64_t randbelow(uint64_t n) {
    int k = 0;
    for (uint64_t t = n; t != 0; t >>= 1) ++k;   // n.bit_length()
    uint64_t r = getrandbits(k);
    while (r >= n) r = getrandbits(k);
    return r;
  }
};

}  // namespace

extern "C" int llp_py_random_sample(uint32_t* mt_state, uint64_t n, int64_t k, int64_t* out) {
  if (mt_state == nullptr || out == nullptr || k < 0 || n == 0 || (uint64_t)k > n || n >= (1ull << 63)) return LLP_E_BADARG;
  if (mt_state[kN] > (uint32_t)kN) return LLP_E_BADARG;
  MT mt;
  memcpy(mt.s, mt_state, sizeof(mt.s));
  mt.pos = mt_state[kN];
  // setsize = 21; if k > 5: setsize += 4 ** _ceil(_log(k * 3, 4))    (math.log(x, 4) == log(x) / log(4) in C doubles)
  double setsize = 21.0;
  if (k > 5) setsize += pow(4.0, ceil(log((double)(k * 3)) / log(4.0)));
  if ((double)n <= setsize) {
    std::vector<int64_t> pool((size_t)n);
    for (uint64_t i = 0; i < n; ++i) pool[i] = (int64_t)i;
    for (int64_t i = 0; i < k; ++i) {
      const uint64_t j = mt.randbelow(n - (uint64_t)i);
      out[i] = pool[j];
      pool[j] = pool[n - (uint64_t)i - 1];   // move non-selected item into vacancy
    }
  } else {
    // `selected` set: open addressing over a power-of-two table (load <= 1/2), empty slot = all ones (never a valid
    // index).  32-bit slots when the population allows it: the table of a 72k-candidate draw is 1 MB instead of 4 MB
    // (it is filled at random, so its size against the host's cache is what a draw costs).
    size_t cap = 16;
    int shift = 60;   // 64 - log2(cap): the multiplicative hash keeps its HIGH bits
    while (cap < (size_t)k * 2) { cap <<= 1; --shift; }
    const size_t mask = cap - 1;
    auto run = [&](auto empty) {
      using Slot = decltype(empty);
      std::vector<Slot> table(cap, empty);
      auto insert = [&](uint64_t v) -> bool {   // false if v was already selected
        size_t h = (size_t)((v * 0x9E3779B97F4A7C15ull) >> shift) & mask;
        const Slot key = (Slot)v;
        while (table[h] != empty) {
          if (table[h] == key) return false;
          h = (h + 1) & mask;
        }
        table[h] = key;
        return true;
      };
      int bits = 0;
      for (uint64_t t = n; t != 0; t >>= 1) ++bits;   // n.bit_length(), hoisted out of randbelow
      // random.sample = "walk the stream of randbelow(n) values, keep first occurrences, stop at k".  The values do not
      // depend on the set, so they are drawn a batch ahead and their table slots prefetched (the table is hit at random:
      // a cache miss per draw otherwise); a batch never holds more values than outputs still missing, so the generator
      // stops exactly where CPython's loop stops.
      constexpr int kBatch = 32;
      uint64_t cand[kBatch];
      int64_t done = 0;
      while (done < k) {
        const int nb = (int)(k - done < kBatch ? k - done : kBatch);
        if (bits <= 32) {
          // _randbelow's rejection loop without a data-dependent branch (45 % of the draws are rejected when n sits just
          // above a power of two: a mispredicted branch per draw otherwise)
          const int sh = 32 - bits;
          int c = 0;
          while (c < nb) {
            const uint64_t j = (uint64_t)(mt.next() >> sh);
            cand[c] = j;
            c += j < n ? 1 : 0;
          }
        } else {
          for (int q = 0; q < nb; ++q) {
            uint64_t j;
            do { j = mt.getrandbits(bits); } while (j >= n);
            cand[q] = j;
          }
        }
        for (int q = 0; q < nb; ++q) __builtin_prefetch(&table[(size_t)((cand[q] * 0x9E3779B97F4A7C15ull) >> shift) & mask], 1, 1);
        for (int q = 0; q < nb; ++q)
          if (insert(cand[q])) out[done++] = (int64_t)cand[q];
      }
    };
    if (n < 0xffffffffull) run((uint32_t)0xffffffffu);
    else run(~0ull);
  }
  memcpy(mt_state, mt.s, sizeof(mt.s));
  mt_state[kN] = mt.pos;
  return 0;
}
