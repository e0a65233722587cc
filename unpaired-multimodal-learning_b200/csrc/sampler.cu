// Host-side epoch permutation, bit-exact with torch.randperm(n, generator=Generator().manual_seed(seed)) on
// the CPU - the call torch's RandomSampler makes once per epoch (reference: the DataLoader objects of
// vision_language/finetune.py:370-371; torch/utils/data/sampler.py RandomSampler.__iter__).
//
// torch's algorithm (ATen randperm_cpu, n < 2^32/20): r = 0..n-1, then for i in [0, n-1):
//   z = mt19937() % (n - i); swap(r[i], r[i + z])        with at::mt19937 seeded by the low 32 bits of `seed`.
//
// Two properties make this cheap enough for a GPU that consumes > 150 M rows/s:
//  * after iteration i the prefix r[0..i] is FINAL (later iterations only touch indices > i), so the permutation
//    can be produced incrementally - the first batch of an epoch needs B iterations, not n.  uml_randperm_begin /
//    uml_randperm_advance expose that; a host thread keeps the prefix ahead of the training loop;
//  * the random draws do not depend on the data: the swap targets are drawn 64 iterations ahead of the swaps and
//    prefetched, the swaps themselves stay in program order.
// torch: 57 ms for the ImageNet bank (1.28 M rows, latency-bound random swaps over 10 MB); here ~6.5 ms.
#include <sched.h>
#include <time.h>

#include <atomic>
#include <thread>
#include <vector>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int kAhead = 64;  // swap targets generated (and prefetched) this many iterations ahead of the swaps

struct PermState {
  uint32_t mt[624];
  int32_t next;          // next unread word of mt[]
  uint32_t ring[kAhead]; // z values of iterations [done, made)
  int64_t n;
  int64_t done;          // iterations performed = length of the final prefix (n - 1 iterations finish everything)
  int64_t made;          // iterations whose z has been drawn (done <= made <= min(done + kAhead, n - 1))
  int64_t* out;
  uint32_t* work;        // the permutation being built, 32-bit, in the UPPER half of out's memory (half the cache
                         // footprint of the random swaps: 5 MB instead of 10 MB for the ImageNet bank)
  int64_t converted;     // out[0..converted) holds the final int64 prefix
  int64_t published;     // length of the final prefix as seen by OTHER threads (release/acquire); -1 = failed
  int64_t next_filled;   // 1 once the NEXT epoch's buffer holds the identity (uml_randperm_run's last act)
};
static_assert(sizeof(PermState) <= UML_RANDPERM_STATE_BYTES, "UML_RANDPERM_STATE_BYTES too small");

inline void mt_seed(PermState& s, uint32_t seed) {
  s.mt[0] = seed;
  for (uint32_t j = 1; j < 624; ++j) s.mt[j] = 1812433253u * (s.mt[j - 1] ^ (s.mt[j - 1] >> 30)) + j;
  s.next = 624;
}

inline void mt_twist(uint32_t* m) {
  constexpr uint32_t kUpper = 0x80000000u, kLower = 0x7fffffffu, kMag = 0x9908b0dfu;
  int k = 0;
  for (; k < 624 - 397; ++k) {
    const uint32_t y = (m[k] & kUpper) | (m[k + 1] & kLower);
    m[k] = m[k + 397] ^ (y >> 1) ^ ((y & 1u) ? kMag : 0u);
  }
  for (; k < 623; ++k) {
    const uint32_t y = (m[k] & kUpper) | (m[k + 1] & kLower);
    m[k] = m[k + 397 - 624] ^ (y >> 1) ^ ((y & 1u) ? kMag : 0u);
  }
  const uint32_t y = (m[623] & kUpper) | (m[0] & kLower);
  m[623] = m[396] ^ (y >> 1) ^ ((y & 1u) ? kMag : 0u);
}

inline uint32_t mt_next(PermState& s) {
  if (s.next >= 624) {
    mt_twist(s.mt);
    s.next = 0;
  }
  uint32_t y = s.mt[s.next++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// ---- the draw stage in bulk (helper thread of uml_randperm_run): tempered MT words by the block, remainders without the
// integer divider.  x % d for 32-bit x, d through one double division: q = floor(x / d) is exact in double (the true
// quotient is at least 1 / d >= 2^-32 away from the next integer in relative terms 2^-32 >> 2^-53), r = x - q * d.  Four
// to eight quotients per divide instruction with AVX2 / AVX-512 against one 25-cycle `div` each - the draws were the
// slower stage of the two-thread pipeline.
inline void mt_fill(PermState& s, uint32_t* dst, int64_t cnt) {
  while (cnt > 0) {
    if (s.next >= 624) {
      mt_twist(s.mt);
      s.next = 0;
    }
    const int64_t take = cnt < 624 - s.next ? cnt : 624 - s.next;
    const uint32_t* src = s.mt + s.next;
    for (int64_t j = 0; j < take; ++j) {
      uint32_t y = src[j];
      y ^= y >> 11;
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= y >> 18;
      dst[j] = y;
    }
    s.next += static_cast<int32_t>(take);
    dst += take;
    cnt -= take;
  }
}

#define UML_MOD_BLOCK_BODY                                                                                       \
  for (int64_t k = 0; k < cnt; ++k) {                                                                            \
    const uint32_t x = y[k], d = d0 - static_cast<uint32_t>(k);                                                  \
    /* unsigned -> double through the signed conversion the vector units have */                                 \
    const double xd = static_cast<double>(static_cast<int32_t>(x ^ 0x80000000u)) + 2147483648.0;                 \
    const double dd = static_cast<double>(static_cast<int32_t>(d));  /* d < 2^31: n < 2^32 / 20 */               \
    const uint32_t q = static_cast<uint32_t>(static_cast<int64_t>(xd / dd));                                     \
    z[k] = x - q * d;                                                                                            \
  }

__attribute__((target("avx2"))) void mod_block_avx2(const uint32_t* __restrict__ y, uint32_t* __restrict__ z, int64_t cnt, uint32_t d0) {
  UML_MOD_BLOCK_BODY
}
void mod_block_generic(const uint32_t* __restrict__ y, uint32_t* __restrict__ z, int64_t cnt, uint32_t d0) {
  UML_MOD_BLOCK_BODY
}
#undef UML_MOD_BLOCK_BODY

inline void mod_block(const uint32_t* y, uint32_t* z, int64_t cnt, uint32_t d0) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) mod_block_avx2(y, z, cnt, d0);
  else mod_block_generic(y, z, cnt, d0);
}

// draw the z of iteration s.made (the draws happen strictly in iteration order, as in torch) and prefetch its target
inline void draw_one(PermState& s) {
  const uint32_t z = mt_next(s) % static_cast<uint32_t>(s.n - s.made);
  s.ring[s.made % kAhead] = z;
  __builtin_prefetch(s.work + s.made + z, 1, 1);
  ++s.made;
}

// iterations [s.done, upto) ; upto <= n - 1.  One fused loop: the out-of-order core overlaps the generator, the
// division and the (prefetched) swap of different iterations - measured faster than three blocked passes.
void advance(PermState& s, int64_t upto) {
  const int64_t m = s.n - 1;
  uint32_t* r = s.work;
  while (s.made < m && s.made < s.done + kAhead) draw_one(s);
  for (int64_t i = s.done; i < upto; ++i) {
    const uint32_t z = s.ring[i % kAhead];
    if (s.made < m) draw_one(s);  // refills the slot just read (made == i + kAhead)
    const uint32_t t = r[i];
    r[i] = r[i + z];
    r[i + z] = t;
  }
  s.done = upto;
}

// out[converted..upto) <- the final 32-bit entries.  The work array occupies bytes [4n, 8n) of out: writing out[j]
// (bytes 8j..8j+7) can only overwrite work entries 2j-n and 2j-n+1, both <= j, i.e. already converted.
void convert(PermState& s, int64_t upto) {
  for (int64_t j = s.converted; j < upto; ++j) s.out[j] = static_cast<int64_t>(s.work[j]);
  if (upto > s.converted) s.converted = upto;
}

}  // namespace

extern "C" {

static int randperm_begin(void* state, uint64_t seed, int64_t n, int64_t* out, bool prefilled);

int uml_randperm_begin(void* state, uint64_t seed, int64_t n, int64_t* out) {
  return randperm_begin(state, seed, n, out, false);
}

static int randperm_begin(void* state, uint64_t seed, int64_t n, int64_t* out, bool prefilled) {
  UML_REQUIRE(state && (out != nullptr || n == 0), "randperm: null pointer");
  UML_REQUIRE(n >= 0 && n < static_cast<int64_t>(UINT32_MAX / 20), "randperm: n=%lld outside the 32-bit sampler range",
              static_cast<long long>(n));
  PermState& s = *static_cast<PermState*>(state);
  mt_seed(s, static_cast<uint32_t>(seed & 0xffffffffu));
  s.n = n;
  s.done = 0;
  s.made = 0;
  s.out = out;
  s.work = reinterpret_cast<uint32_t*>(out) + n;  // upper half of the n x 8 bytes
  s.converted = 0;
  if (!prefilled)
    for (int64_t i = 0; i < n; ++i) s.work[i] = static_cast<uint32_t>(i);
  return 0;
}

int uml_randperm_advance(void* state, int64_t upto) {
  UML_REQUIRE(state, "randperm: null state");
  PermState& s = *static_cast<PermState*>(state);
  // a final prefix of length `upto` needs min(upto, n - 1) iterations (the last element falls into place)
  int64_t iters = upto < s.n - 1 ? upto : s.n - 1;
  if (iters < 0) iters = 0;
  if (iters > s.done) advance(s, iters);
  // iteration i makes element i final; after n - 1 iterations the last element is in place too
  const int64_t final_len = s.done >= s.n - 1 ? s.n : s.done;
  convert(s, upto < final_len ? upto : final_len);
  return 0;
}

// The whole permutation, chunk by chunk, publishing the length of the final prefix after every chunk.  Meant to
// be the body of a host thread (one C call, so a Python caller's interpreter lock is never needed in between).
// `prefilled`: `out` already holds the identity.  `next_out` (optional): once this permutation is complete the
// thread writes the identity into the NEXT epoch's buffer, so that epoch starts with the first swap instead of
// a 10 MB fill (the seed of the next epoch is not known yet, the identity is).
int uml_randperm_run(void* state, uint64_t seed, int64_t n, int64_t* out, int64_t chunk, int32_t prefilled,
                     int64_t* next_out) {
  UML_REQUIRE(state && chunk > 0, "randperm_run: bad arguments");
  PermState& s = *static_cast<PermState*>(state);
  __atomic_store_n(&s.published, static_cast<int64_t>(0), __ATOMIC_RELEASE);
  __atomic_store_n(&s.next_filled, static_cast<int64_t>(0), __ATOMIC_RELEASE);
  int rc = randperm_begin(state, seed, n, out, prefilled != 0);
  if (rc != 0) {
    __atomic_store_n(&s.published, static_cast<int64_t>(-1), __ATOMIC_RELEASE);
    return rc;
  }
  const int64_t m = n > 0 ? n - 1 : 0;  // swaps to perform
  static int pipelined = -1;
  if (pipelined < 0) {
    const char* e = getenv("UML_SAMPLER_PIPELINE");
    pipelined = (e && e[0] == '0') ? 0 : 1;
  }
  if (!pipelined) {  // single thread: draw-ahead ring + swaps fused in one loop
    int64_t upto = 0;
    while (upto < n) {
      upto = upto + chunk < n ? upto + chunk : n;
      rc = uml_randperm_advance(state, upto);
      if (rc != 0) {
        __atomic_store_n(&s.published, static_cast<int64_t>(-1), __ATOMIC_RELEASE);
        return rc;
      }
      __atomic_store_n(&s.published, upto, __ATOMIC_RELEASE);
    }
    if (next_out) {
      uint32_t* w = reinterpret_cast<uint32_t*>(next_out) + n;
      for (int64_t i = 0; i < n; ++i) w[i] = static_cast<uint32_t>(i);
      __atomic_store_n(&s.next_filled, static_cast<int64_t>(1), __ATOMIC_RELEASE);
    }
    return 0;
  }
  // Two-stage pipeline: a helper thread draws the swap targets (MT19937 + modulo, ~3 ns each), this thread performs
  // the swaps (~3 ns each, cache-miss bound) and publishes the final prefix - the two halves of the work overlap.
  constexpr int kBlk = 4096, kRing = 16;
  std::vector<uint32_t> ring(static_cast<size_t>(kBlk) * kRing);
  std::atomic<int64_t> produced{0}, consumed{0};  // in blocks
  const int64_t n_blocks = (m + kBlk - 1) / kBlk;
  std::thread producer([&] {
    for (int64_t b = 0; b < n_blocks; ++b) {
      while (b - consumed.load(std::memory_order_acquire) >= kRing) sched_yield();
      uint32_t* z = ring.data() + (b % kRing) * kBlk;
      const int64_t i0 = b * kBlk, cnt = m - i0 < kBlk ? m - i0 : kBlk;
      uint32_t y[kBlk];
      mt_fill(s, y, cnt);
      mod_block(y, z, cnt, static_cast<uint32_t>(n - i0));  // z[k] = y[k] % (n - (i0 + k))
      produced.store(b + 1, std::memory_order_release);
    }
  });
  uint32_t* r = s.work;
  int64_t next_pub = chunk < n ? chunk : n;
  for (int64_t b = 0; b < n_blocks; ++b) {
    while (produced.load(std::memory_order_acquire) <= b) sched_yield();
    const uint32_t* z = ring.data() + (b % kRing) * kBlk;
    const int64_t i0 = b * kBlk, cnt = m - i0 < kBlk ? m - i0 : kBlk;
    for (int64_t k = 0; k < cnt; ++k) {
      if (k + kAhead < cnt) __builtin_prefetch(r + i0 + k + kAhead + z[k + kAhead], 1, 1);
      const int64_t i = i0 + k;
      const uint32_t t = r[i];
      r[i] = r[i + z[k]];
      r[i + z[k]] = t;
    }
    consumed.store(b + 1, std::memory_order_release);
    s.done = i0 + cnt;
    while (next_pub <= s.done && next_pub < n) {  // publish whole chunks of the final prefix
      convert(s, next_pub);
      __atomic_store_n(&s.published, next_pub, __ATOMIC_RELEASE);
      next_pub = next_pub + chunk < n ? next_pub + chunk : n;
    }
  }
  producer.join();
  s.made = m;
  convert(s, n);
  __atomic_store_n(&s.published, n, __ATOMIC_RELEASE);
  if (next_out) {
    uint32_t* w = reinterpret_cast<uint32_t*>(next_out) + n;  // the work half of the next epoch's buffer
    for (int64_t i = 0; i < n; ++i) w[i] = static_cast<uint32_t>(i);
    __atomic_store_n(&s.next_filled, static_cast<int64_t>(1), __ATOMIC_RELEASE);
  }
  return 0;
}

// 1 when the thread running uml_randperm_run on `state` has finished writing the identity into next_out
int uml_randperm_next_filled(const void* state) {
  return static_cast<int>(__atomic_load_n(&static_cast<const PermState*>(state)->next_filled, __ATOMIC_ACQUIRE));
}

// Blocks (yielding the core) until uml_randperm_run on another thread has made out[0..upto) final.
int uml_randperm_wait(const void* state, int64_t upto) {
  UML_REQUIRE(state, "randperm_wait: null state");
  const PermState& s = *static_cast<const PermState*>(state);
  for (int spins = 0;; ++spins) {
    const int64_t p = __atomic_load_n(&s.published, __ATOMIC_ACQUIRE);
    UML_REQUIRE(p >= 0, "randperm_wait: the generating thread failed");
    if (p >= upto) return 0;
    if (spins < 64) {
      sched_yield();
    } else {  // leave the core to the producer (the process may be confined to very few CPUs)
      timespec ts = {0, 20000};
      nanosleep(&ts, nullptr);
    }
  }
}

int uml_randperm_i64(uint64_t seed, int64_t n, int64_t* out) {
  alignas(16) unsigned char buf[UML_RANDPERM_STATE_BYTES];
  int rc = uml_randperm_begin(buf, seed, n, out);
  if (rc) return rc;
  return uml_randperm_advance(buf, n);
}

}  // extern "C"
