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

#include <vector>

#include "../../include/llp_b200.h"

namespace {

constexpr int kN = 624, kM = 397;

struct MT {
  uint32_t* s;   // 624 state words
  uint32_t* pos; // index (0..624)
  uint32_t next() {
    static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
    if (*pos >= (uint32_t)kN) {
      int kk;
      uint32_t y;
      for (kk = 0; kk < kN - kM; kk++) {
        y = (s[kk] & 0x80000000u) | (s[kk + 1] & 0x7fffffffu);
        s[kk] = s[kk + kM] ^ (y >> 1) ^ mag01[y & 0x1u];
      }
      for (; kk < kN - 1; kk++) {
        y = (s[kk] & 0x80000000u) | (s[kk + 1] & 0x7fffffffu);
        s[kk] = s[kk + (kM - kN)] ^ (y >> 1) ^ mag01[y & 0x1u];
      }
      y = (s[kN - 1] & 0x80000000u) | (s[0] & 0x7fffffffu);
      s[kN - 1] = s[kM - 1] ^ (y >> 1) ^ mag01[y & 0x1u];
      *pos = 0;
    }
    uint32_t y = s[(*pos)++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
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
  MT mt{mt_state, mt_state + kN};
  if (*mt.pos > (uint32_t)kN) return LLP_E_BADARG;
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
    // `selected` set: open addressing over a power-of-two table (load <= 1/4), empty slot = ~0 (never a valid index)
    size_t cap = 16;
    while (cap < (size_t)k * 4) cap <<= 1;
    std::vector<uint64_t> table(cap, ~0ull);
    const size_t mask = cap - 1;
    auto insert = [&](uint64_t v) -> bool {   // false if v was already selected
      size_t h = (size_t)((v * 0x9E3779B97F4A7C15ull) >> 17) & mask;
      while (table[h] != ~0ull) {
        if (table[h] == v) return false;
        h = (h + 1) & mask;
      }
      table[h] = v;
      return true;
    };
    for (int64_t i = 0; i < k; ++i) {
      uint64_t j = mt.randbelow(n);
      while (!insert(j)) j = mt.randbelow(n);
      out[i] = (int64_t)j;
    }
  }
  return 0;
}
