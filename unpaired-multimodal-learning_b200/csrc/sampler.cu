// Host-side epoch permutation, bit-exact with torch.randperm(n, generator=Generator().manual_seed(seed)) on
// the CPU - the call torch's RandomSampler makes once per epoch (reference: the DataLoader objects of
// vision_language/finetune.py:370-371; torch/utils/data/sampler.py RandomSampler.__iter__).
//
// torch's algorithm (ATen randperm_cpu, n < 2^32/20): r = 0..n-1, then for i in [0, n-1):
//   z = mt19937() % (n - i); swap(r[i], r[i + z])        with at::mt19937 seeded by the low 32 bits of `seed`.
// The swaps hit a 10 MB array at random for the ImageNet bank (1.28 M rows) and are latency bound (~45 ns
// each in torch: 57 ms per epoch, against 13 ms of GPU work per epoch at the throughput batch).  The random
// sequence does not depend on the data, so the swap targets are generated a window ahead and prefetched; the
// swaps themselves stay in program order, which keeps the result identical.
#include <cstdint>
#include <cstring>

#include "common.cuh"

namespace {

struct Mt19937 {  // MT19937, the parameters of std::mt19937 / at::mt19937
  uint32_t s[624];
  int next;
  explicit Mt19937(uint32_t seed) {
    s[0] = seed;
    for (uint32_t j = 1; j < 624; ++j) s[j] = 1812433253u * (s[j - 1] ^ (s[j - 1] >> 30)) + j;
    next = 624;
  }
  void twist() {
    for (int k = 0; k < 624; ++k) {
      const uint32_t y = (s[k] & 0x80000000u) | (s[(k + 1) % 624] & 0x7fffffffu);
      s[k] = s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    next = 0;
  }
  inline uint32_t operator()() {
    if (next >= 624) twist();
    uint32_t y = s[next++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};

constexpr int kAhead = 64;  // swap targets generated (and prefetched) this many iterations early

}  // namespace

extern "C" {

int uml_randperm_i64(uint64_t seed, int64_t n, int64_t* out) {
  UML_REQUIRE(out != nullptr || n == 0, "randperm: null output");
  UML_REQUIRE(n >= 0 && n < static_cast<int64_t>(UINT32_MAX / 20), "randperm: n=%lld outside the 32-bit sampler range",
              static_cast<long long>(n));
  for (int64_t i = 0; i < n; ++i) out[i] = i;
  if (n < 2) return 0;
  Mt19937 gen(static_cast<uint32_t>(seed & 0xffffffffu));
  const int64_t m = n - 1;  // number of swaps
  uint32_t ring[kAhead];
  int64_t made = 0;
  for (; made < kAhead && made < m; ++made) {
    const uint32_t z = gen() % static_cast<uint32_t>(n - made);
    ring[made % kAhead] = z;
    __builtin_prefetch(out + made + z, 1, 1);
  }
  for (int64_t i = 0; i < m; ++i) {
    const uint32_t z = ring[i % kAhead];
    if (made < m) {
      const uint32_t zn = gen() % static_cast<uint32_t>(n - made);
      ring[made % kAhead] = zn;
      __builtin_prefetch(out + made + zn, 1, 1);
      ++made;
    }
    const int64_t t = out[i];
    out[i] = out[i + z];
    out[i + z] = t;
  }
  return 0;
}

}  // extern "C"
