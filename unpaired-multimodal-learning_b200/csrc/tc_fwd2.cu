// K2+K3 on the tensor cores, second generation: head forward, logit scale, softmax cross-entropy and the FINAL logit
// gradient G in ONE kernel and ONE pass over G (reference: engine/models/head.py:80-82,133-135 +
// finetune.py:186-188 + the autograd of F.cross_entropy).
//
//   logits[b,c] = scale_b * sum_d X[b,d] W[c,d]         bf16 x bf16 -> fp32 in TMEM
//   G[b,c]      = w_b * scale_b / n_b * (softmax(logits)[b,c] - [c == y_b])     written once, as bf16
//
// tc_fwd.cu walks the (up to four) 256-class chunks of a row tile one after the other inside one CTA pair; a row's
// final max / sum is then only known after its first chunks have left the SM, which cost a second pass over G
// (g_fixup_kernel: 155 MB of traffic + a launch gap, a quarter of the step) and re-streamed the X tile per chunk.
// Here the class chunks of a row tile are computed AT THE SAME TIME by different CTA pairs:
//
//   * CTA pair (cluster of 2, cta_group::2, M = 256) p owns class chunk p % n_chunks for good - and walks the 256-row
//     units of group p / n_chunks.  X tile and the pair's half of the W chunk come by TMA (4 stages, SWIZZLE_128B),
//     accumulators live in TMEM, double buffered (2 x 256 columns): the MMA of unit i+1 overlaps the epilogue of unit i.
//   * epilogue, 8 warps, one thread per (row, 128-column half): ONE sweep over the accumulator, 32 columns at a time
//     with the next tcgen05.ld in flight: running max (in the raw-logit domain, so that the maximal element's
//     exponential is exactly 1 and the cross entropy can never come out negative), exp(l - m_running) staged as
//     bf16 in shared memory (the whole 128 x 256 tile, 64 KB, TMA-store layout).  The TMEM buffer is handed back
//     to the MMA warp right after this sweep.
//   * the pairs of a group exchange one 16-byte record per row - {max, sum, sum p*raw, argmax} - through L2
//     (release/acquire flags keyed by a per-launch epoch: no memset, no cluster wider than the pair, every SM usable),
//   * then every thread rescales ITS OWN staged half row by  exp(m_group - M) * coef / S  (two bf16x2 FMAs per pair
//     with the factor split hi + lo, so the product carries fp32-level accuracy), patches the one-hot column and the
//     warp's 32 x 64 boxes leave as TMA stores.  G is written once and is final.
//   * the pair that owns a row's label column computes loss / hit / d(scale) for that row; per-(tile, chunk, warp)
//     partial sums go to tile_part and are reduced in a fixed order by the next kernel of the step (the dW GEMM's
//     idle warp) or by uml_reduce_tile_stats.
// All CTAs of the grid (<= one per SM) are co-resident, which the flag exchange relies on; a watchdog turns a
// missing peer into an error flag (uml_fwd_x_failed) instead of a hang.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace uml {

constexpr int kXStages = 4;
constexpr int kXABytes = 128 * 64 * 2;                 // X tile: 128 rows x 64 k
constexpr int kXBBytes = 128 * 64 * 2;                 // this CTA's half of the W chunk: 128 classes x 64 k
constexpr int kXStageBytes = kXABytes + kXBBytes;
constexpr int kXBoxBytes = 32 * 128;                   // staging / TMA-store box: 32 rows x 64 bf16 columns
constexpr int kXStagingBytes = 128 * 256 * 2;          // 16 boxes: [column group of 64][row quarter]
constexpr int kXHalfFloats = 8;                        // record the two column halves of a row exchange
constexpr int kXHxBytes = 2 * 2 * 128 * kXHalfFloats * 4;  // [tile parity][half][row]
constexpr int kXSmemBytes = kXStages * kXStageBytes + kXStagingBytes + kXHxBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kXWarpAlloc = 8, kXWarpMma = 10, kXWarpTma = 11;
constexpr int kXThreads = 384;
constexpr int kXMaxChunks = 4;

struct XSegs {
  int64_t n0;
  const float* scale_dev[2];
  float scale[2], dcoef[2];  // dcoef = w/n ; the logit-gradient coefficient is dcoef * scale
};

struct XWork {
  float* tile_part;      // [(tile * n_chunks + chunk) * 8 + q * 2 + seg] x 4 floats
  uint4* recs;           // [(row * n_chunks + chunk) * 2 + {0, 1}]: {max, epoch, sum, epoch}, {sum p*raw, epoch, argmax, epoch}
  unsigned* ctrl;        // [0] epoch of the last finished launch, [1] finished CTAs, [2] watchdog flag, [3] partial records
};

__device__ __forceinline__ float x_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// value + epoch word pairs: 8-byte units are delivered whole, so a matching epoch vouches for its value (the scheme of
// NCCL's LL protocol); volatile = every poll goes to L2
__device__ __forceinline__ void st_volatile_v4(uint4* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// explicit shared-space accesses (32-bit addresses): the generic-pointer forms compiled to LD.E / ST.E
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t a, unsigned short x) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(x) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned short lds_u16(uint32_t a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void tma_store_2d_s(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// makes the 32 registers of a tcgen05.ld "change" after the wait, so that no use can be scheduled above it
__device__ __forceinline__ void pin32(uint32_t (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 8)
    asm volatile("" : "+r"(v[i]), "+r"(v[i + 1]), "+r"(v[i + 2]), "+r"(v[i + 3]), "+r"(v[i + 4]), "+r"(v[i + 5]),
                      "+r"(v[i + 6]), "+r"(v[i + 7]));
}

#ifdef UML_FWD_TIMING
__device__ long long g_fwdx_dbg[148 * 16];
#define XDBG_DECL() long long _acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long _t0 = clock64()
#define XDBG_MARK() _t0 = clock64()
#define XDBG_ACC(slot) do { long long _t1 = clock64(); _acc[slot] += _t1 - _t0; _t0 = _t1; } while (0)
#define XDBG_FLUSH(base, n) do { for (int _i = 0; _i < (n); ++_i) g_fwdx_dbg[blockIdx.x * 16 + (base) + _i] = _acc[_i]; } while (0)
#else
#define XDBG_DECL()
#define XDBG_MARK()
#define XDBG_ACC(slot)
#define XDBG_FLUSH(base, n)
#endif

template <bool kPred, bool kDs>
__global__ void __launch_bounds__(kXThreads, 1)
    head_fwd_ce_x_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_g, int64_t n_rows, int dim, int n_classes, int n_groups,
                         const int32_t* __restrict__ labels, XSegs segs, int write_g, int64_t ldg,
                         float* __restrict__ row_loss, int32_t* __restrict__ row_pred, int32_t* __restrict__ row_correct,
                         float* __restrict__ row_dscale, XWork wk) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* staging = smem + kXStages * kXStageBytes;  // 1024-aligned
  float* hx = reinterpret_cast<float*>(staging + kXStagingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(hx) + kXHxBytes);
  uint64_t* empty_bar = full_bar + kXStages;
  uint64_t* tfull_bar = empty_bar + kXStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint32_t* epoch_slot = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_kb = (dim + 63) / 64;
  const int n_chunks = (n_classes + 255) / 256;
  const int pair = blockIdx.x >> 1;
  const int chunk = pair % n_chunks, group = pair / n_chunks;
  const int col0 = chunk * 256;
  const int n_valid = n_classes - col0 < 256 ? n_classes - col0 : 256;  // class columns of this chunk
  const int n_mma = ((n_valid + 15) / 16) * 16;                          // N of the pair's MMA
  const int64_t n_units = (n_rows + 255) / 256;

  if (warp == kXWarpTma && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    if (write_g) tma_prefetch_desc(&tmap_g);
  }
  if (warp == kXWarpMma && lane == 0) {
    for (int s = 0; s < kXStages; ++s) {
      mbar_init(&full_bar[s], 2);  // one arrival per producer of the pair; tx bytes are counted on the leader
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 16);  // one arrival per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == kXWarpAlloc) tmem_alloc_cg2(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();  // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) {
    *epoch_slot = *reinterpret_cast<volatile unsigned*>(wk.ctrl) + 1u;
    if (blockIdx.x == 0) wk.ctrl[3] = static_cast<unsigned>(n_units * 2 * n_chunks * 4);  // partial records per run
  }
  __syncthreads();
  const unsigned epoch = *epoch_slot;

  if (warp == kXWarpTma) {
    // ------------------------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      XDBG_DECL();
      const uint32_t lead_bar0 = mapa_cta(smem_u32(&full_bar[0]), 0);
      const int32_t wrow0 = col0 + static_cast<int32_t>(rank) * (n_mma / 2);
      for (int64_t unit = group; unit < n_units; unit += n_groups) {
        const int32_t row0 = static_cast<int32_t>((unit * 2 + rank) * 128);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kXStages, ph = (it / kXStages) & 1;
          XDBG_MARK();
          mbar_wait(&empty_bar[s], ph ^ 1);
          XDBG_ACC(0);
          unsigned char* a = smem + s * kXStageBytes;
          const uint32_t lead_bar = lead_bar0 + s * 8;
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * kXStageBytes);
          tma_load_2d_cg2(a, &tmap_x, lead_bar, kb * 64, row0);
          tma_load_2d_cg2(a + kXABytes, &tmap_w, lead_bar, kb * 64, wrow0);
          if (!leader) mbar_arrive_remote(lead_bar);
          XDBG_ACC(1);
        }
      }
      XDBG_FLUSH(12, 2);
    }
    __syncwarp();
  } else if (warp == kXWarpMma) {
    // ------------------------------------------------ MMA issuer --------------------------------
    if (lane == 0 && leader) {
      const uint32_t idesc = make_idesc_bf16(256, static_cast<uint32_t>(n_mma), 0, 0);
      uint32_t it = 0, acc_it = 0;
      XDBG_DECL();
      for (int64_t unit = group; unit < n_units; unit += n_groups, ++acc_it) {
        const uint32_t b = acc_it & 1, aph = (acc_it >> 1) & 1;
        XDBG_MARK();
        mbar_wait(&tempty_bar[b], aph ^ 1);
        XDBG_ACC(0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * 256;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kXStages, ph = (it / kXStages) & 1;
          mbar_wait(&full_bar[s], ph);
          XDBG_ACC(1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * kXStageBytes);
          const uint32_t b_addr = a_addr + kXABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // K-major, 128B swizzle: 8-row groups are 1024 B apart; a K step of 16 bf16 = 32 B
            const uint64_t da = make_smem_desc(a_addr + k * 32, 16, 1024, kLayoutSw128);
            const uint64_t db = make_smem_desc(b_addr + k * 32, 16, 1024, kLayoutSw128);
            umma_bf16_cg2(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit_cg2(&empty_bar[s]);  // frees the stage in both CTAs once these MMAs have read it
          XDBG_ACC(2);
        }
        umma_commit_cg2(&tfull_bar[b]);  // accumulator complete: both CTAs' epilogues wake
      }
      XDBG_FLUSH(9, 3);
    }
    __syncwarp();
  } else if (warp < 8) {
    // ------------------------------------------------ epilogue ----------------------------------
    const int q = warp & 3;   // TMEM lane quarter = row quarter of the tile
    const int h = warp >> 2;  // column half of the chunk: columns [h * 128, h * 128 + 128)
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kMasked = -1.0e30f;  // padded class column: never the maximum, exponential exactly 0, 0 * it finite
    const int rloc = q * 32 + lane;
    // this warp's two staging boxes (32 rows x 64 columns each): box j at + j * 4 * kXBoxBytes; this thread's row
    const uint32_t box0 = smem_u32(staging) + ((h * 2) * 4 + q) * kXBoxBytes;
    const uint32_t srow0 = box0 + lane * 128;
    const uint32_t swz = static_cast<uint32_t>(lane & 7) << 4;  // 16-byte chunk c of row r sits at chunk c ^ (r & 7)
    const uint32_t hx_base = smem_u32(hx);
    const uint32_t tempty_remote0 = mapa_cta(smem_u32(&tempty_bar[0]), 0);
    uint32_t tile_it = 0;
    XDBG_DECL();
    for (int64_t unit = group; unit < n_units; unit += n_groups, ++tile_it) {
      const int64_t tile = unit * 2 + rank;
      const int64_t row = tile * 128 + rloc;
      const bool valid = row < n_rows;
      const bool sg = valid && row >= segs.n0;
      const float* sdev = sg ? segs.scale_dev[1] : segs.scale_dev[0];
      const float scale = sdev ? __ldg(sdev) : (sg ? segs.scale[1] : segs.scale[0]);
      const float dcoef = sg ? segs.dcoef[1] : segs.dcoef[0];
      const float gcoef = dcoef * scale;
      const bool neg = scale < 0.f;                     // (a learnable temperature may in principle go negative)
      const float sgn = neg ? -1.f : 1.f;
      const float sabs = fabsf(scale);
      const float sl2 = sabs * kLog2e;                  // exponent per unit of (sign-adjusted) raw logit, in bits
      const bool neg_any = __any_sync(0xffffffffu, neg);
      const int label = valid ? labels[row] : -1;
      const int lcol = label - col0 - h * 128;          // label position inside this thread's 128 columns
      // running statistics of this thread's half row, in the raw (sign-adjusted) logit domain
      float run_max = -INFINITY, run_sum = 0.f, run_pr = 0.f, max_before = -INFINITY, lab_raw = -INFINITY;
      int arg = 0x7fffffff;
      float gm0 = -INFINITY, gm1 = -INFINITY, gm2 = -INFINITY, gm3 = -INFINITY;  // running max each group was written against

      const uint32_t b = tile_it & 1, aph = (tile_it >> 1) & 1;
      XDBG_MARK();
      mbar_wait(&tfull_bar[b], aph);
      XDBG_ACC(0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * 256 + h * 128;
      if (write_g) {
        if (lane == 0) bulk_wait_read<0>();  // the previous unit's TMA stores have read the staging boxes
        __syncwarp();
      }
      XDBG_ACC(1);
      uint32_t va[32], vb[32];
      tmem_ld32(taddr, va);
      tmem_ld_wait();
      pin32(va);

      // one 32-column group: running max, exponentials relative to it, bf16 staging.  Returns the max it was written against.
      auto run_group = [&](uint32_t (&v)[32], const int g) -> float {
        const int c0l = h * 128 + g * 32;  // first column of the group inside the chunk
        if (c0l >= n_valid) {              // nothing but padding: G stays zero there
          if (write_g && col0 + c0l < ldg) {
            const uint32_t srow = srow0 + (g >> 1) * (4 * kXBoxBytes);
#pragma unroll
            for (int j = 0; j < 4; ++j) sts128(srow + ((static_cast<uint32_t>((g & 1) * 4 + j) << 4) ^ swz), 0u, 0u, 0u, 0u);
          }
          return run_max;
        }
        if (neg_any) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * sgn);
        }
        if (c0l + 32 > n_valid) {  // padded class columns (or stale TMEM beyond the MMA's N)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0l + i >= n_valid) v[i] = __float_as_uint(kMasked);
        }
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(v[i]));
          m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
          m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
        }
        const float bm = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        // hit flag without an argmax index:  argmax == label  <=>  logit[label] == row max  and
        // logit[label] > max over the columns before it  (torch.argmax returns the FIRST maximal index)
        const int d = lcol - g * 32;
        if (d >= 32) {
          max_before = fmaxf(max_before, bm);
        } else if (d >= 0) {
          float b0 = -INFINITY, b1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float r0 = __uint_as_float(v[i]), r1 = __uint_as_float(v[i + 1]);
            b0 = fmaxf(b0, i < d ? r0 : -INFINITY);
            b1 = fmaxf(b1, i + 1 < d ? r1 : -INFINITY);
            if (i == d) lab_raw = r0;
            if (i + 1 == d) lab_raw = r1;
          }
          max_before = fmaxf(max_before, fmaxf(b0, b1));
        }
        if (kPred && bm > run_max) {  // first column holding the new maximum (columns are visited in order)
#pragma unroll
          for (int i = 31; i >= 0; --i)
            if (__uint_as_float(v[i]) == bm) arg = col0 + c0l + i;
        }
        const float new_max = fmaxf(run_max, bm);
        const float resc = x_exp2((run_max - new_max) * sl2);  // exp2(-inf) = 0 on the first group
        run_sum *= resc;
        if (kDs) run_pr *= resc;
        run_max = new_max;
        // exponentials relative to the running max: (raw - max) is exact for the maximal element -> p = 1
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        const uint32_t srow = srow0 + (g >> 1) * (4 * kXBoxBytes);
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 8 columns -> one 16-byte chunk of the staged row
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 8 * j + 2 * e;
            const float r0 = __uint_as_float(v[i]), r1 = __uint_as_float(v[i + 1]);
            const float p0 = x_exp2((r0 - new_max) * sl2);
            const float p1 = x_exp2((r1 - new_max) * sl2);
            if (e & 1) { s2 += p0; s3 += p1; if (kDs) { q2 = fmaf(p0, r0, q2); q3 = fmaf(p1, r1, q3); } }
            else       { s0 += p0; s1 += p1; if (kDs) { q0 = fmaf(p0, r0, q0); q1 = fmaf(p1, r1, q1); } }
            __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
            w[e] = *reinterpret_cast<uint32_t*>(&hh);
          }
          if (write_g) sts128(srow + ((static_cast<uint32_t>((g & 1) * 4 + j) << 4) ^ swz), w[0], w[1], w[2], w[3]);
        }
        run_sum += (s0 + s1) + (s2 + s3);
        if (kDs) run_pr += (q0 + q1) + (q2 + q3);
        return new_max;
      };
      // one sweep, 32 columns at a time, the next tcgen05.ld in flight while a group is processed
#pragma unroll 1
      for (int gp = 0; gp < 2; ++gp) {
        tmem_ld32(taddr + gp * 64 + 32, vb);
        const float ma = run_group(va, gp * 2);
        tmem_ld_wait();
        pin32(vb);
        if (gp == 0) tmem_ld32(taddr + 64, va);
        const float mb = run_group(vb, gp * 2 + 1);
        if (gp == 0) {
          tmem_ld_wait();
          pin32(va);
          gm0 = ma; gm1 = mb;
        } else {
          gm2 = ma; gm3 = mb;
        }
      }
      // accumulator buffer b may be overwritten by the (leader's) MMA warp now
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!leader) mbar_arrive_remote(tempty_remote0 + b * 8);
        else mbar_arrive(&tempty_bar[b]);
      }
      XDBG_ACC(2);

      // ---- the two column halves of a row meet (shared memory) ---------------------------------------
      const uint32_t mine = hx_base + ((((tile_it & 1) * 2 + h) * 128 + rloc) * kXHalfFloats) * 4;
      const uint32_t other = hx_base + ((((tile_it & 1) * 2 + (h ^ 1)) * 128 + rloc) * kXHalfFloats) * 4;
      sts128(mine, __float_as_uint(run_max), __float_as_uint(run_sum), __float_as_uint(run_pr), __float_as_uint(max_before));
      sts64(mine + 16, __float_as_uint(lab_raw), static_cast<uint32_t>(arg));
      named_bar_sync(1, 256);
      const uint4 o4 = lds128(other);
      const uint2 o2 = lds64(other + 16);
      const float o_max = __uint_as_float(o4.x), o_sum = __uint_as_float(o4.y), o_pr = __uint_as_float(o4.z),
                  o_before = __uint_as_float(o4.w), o_lab = __uint_as_float(o2.x);
      const float cm = fmaxf(run_max, o_max);  // chunk-level statistics
      const float e_me = x_exp2((run_max - cm) * sl2), e_ot = x_exp2((o_max - cm) * sl2);
      const float cs = run_sum * e_me + o_sum * e_ot;
      const float cpr = run_pr * e_me + o_pr * e_ot;
      // max over the columns of this chunk that precede the label (meaningful when the label is in this chunk):
      // label in half 0 -> half 0's max_before;  in half 1 -> max(all of half 0, half 1's max_before)
      const int lc_chunk = label - col0;
      const float m_h0 = h == 0 ? run_max : o_max, before_h0 = h == 0 ? max_before : o_before,
                  before_h1 = h == 0 ? o_before : max_before;
      const float before_loc = lc_chunk >= 128 ? fmaxf(m_h0, before_h1) : before_h0;
      const float lraw = fmaxf(lab_raw, o_lab);  // exactly one half saw the label column (the other holds -inf)
      int carg = arg;
      if (kPred) {
        const int oarg = static_cast<int>(o2.y);
        // larger maximum wins; on equal maxima the lower column index (torch.argmax)
        if (o_max > run_max || (o_max == run_max && oarg < arg)) carg = oarg;
      }
      XDBG_ACC(3);

      // ---- the class chunks of a row meet (L2) -------------------------------------------------------
      // Every value travels with the launch epoch in the same 8-byte word pair (which the memory system delivers whole):
      // the reader spins until the epoch matches - no fence, no separate flag, one L2 round trip.
      float M = cm, S = cs, PR = cpr, before_chunks = -INFINITY;
      int garg = carg;
      if (n_chunks > 1) {
        constexpr bool kSecond = kPred || kDs;
        if (h == 0 && valid) {
          uint4* rec = wk.recs + (row * n_chunks + chunk) * 2;
          st_volatile_v4(rec, __float_as_uint(cm), epoch, __float_as_uint(cs), epoch);
          if (kSecond) st_volatile_v4(rec + 1, __float_as_uint(cpr), epoch, static_cast<uint32_t>(carg), epoch);
        }
        XDBG_ACC(4);
        if (valid) {
          float om[kXMaxChunks], os[kXMaxChunks], opr[kXMaxChunks];
          int oa[kXMaxChunks];
#pragma unroll
          for (int c2 = 0; c2 < kXMaxChunks; ++c2) {
            om[c2] = -INFINITY; os[c2] = 0.f; opr[c2] = 0.f; oa[c2] = 0x7fffffff;
            if (c2 < n_chunks && c2 != chunk) {
              const uint4* rec = wk.recs + (row * n_chunks + c2) * 2;
              uint4 r0 = ld_volatile_v4(rec);
              long long t0 = 0;
              while (r0.y != epoch || r0.w != epoch) {
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > 2000000000ll) {  // ~1 s: never hang the GPU; the host checks the flag
                  atomicExch(wk.ctrl + 2, 1u);
                  break;
                }
                r0 = ld_volatile_v4(rec);
              }
              om[c2] = __uint_as_float(r0.x); os[c2] = __uint_as_float(r0.z);
              if (kSecond) {
                uint4 r1 = ld_volatile_v4(rec + 1);
                while (r1.y != epoch || r1.w != epoch) {
                  if (t0 == 0) t0 = clock64();
                  else if (clock64() - t0 > 2000000000ll) {
                    atomicExch(wk.ctrl + 2, 1u);
                    break;
                  }
                  r1 = ld_volatile_v4(rec + 1);
                }
                opr[c2] = __uint_as_float(r1.x); oa[c2] = static_cast<int>(r1.z);
              }
            }
          }
          XDBG_ACC(5);
#pragma unroll
          for (int c2 = 0; c2 < kXMaxChunks; ++c2) {
            M = fmaxf(M, om[c2]);
            if (c2 < chunk) before_chunks = fmaxf(before_chunks, om[c2]);
          }
          const float e_c = x_exp2((cm - M) * sl2);
          S = cs * e_c;
          PR = cpr * e_c;
#pragma unroll
          for (int c2 = 0; c2 < kXMaxChunks; ++c2) {
            const float e = x_exp2((om[c2] - M) * sl2);  // 0 for the absent chunks (-inf)
            S = fmaf(os[c2], e, S);
            if (kDs) PR = fmaf(opr[c2], e, PR);
          }
          if (kPred) {
            float bestm = cm;
            int bestc = chunk;
#pragma unroll
            for (int c2 = 0; c2 < kXMaxChunks; ++c2) {
              if (c2 < n_chunks && c2 != chunk && (om[c2] > bestm || (om[c2] == bestm && c2 < bestc))) {
                bestm = om[c2]; bestc = c2; garg = oa[c2];
              }
            }
          }
        }
        __syncwarp();
      }
      XDBG_ACC(6);

      // ---- rescale the staged half row, patch the one-hot column, store -------------------------------
      const float inv_sum = 1.f / S;
      if (write_g) {
        const float tc = inv_sum * gcoef;
        // the one-hot column, from the ORIGINAL staged element in fp32 (G = p * f - coef, one rounding): computed
        // before the sweep below overwrites it, written after
        const bool patch = lcol >= 0 && lcol < 128;
        uint32_t patch_addr = 0;
        __nv_bfloat16 patch_val = __float2bfloat16_rn(0.f);
        if (patch) {
          const int gl = lcol >> 5;
          const float gmx = gl == 0 ? gm0 : gl == 1 ? gm1 : gl == 2 ? gm2 : gm3;
          const float f = x_exp2((gmx - M) * sl2) * tc;
          patch_addr = srow0 + (lcol >> 6) * (4 * kXBoxBytes) + ((static_cast<uint32_t>((lcol & 63) >> 3) << 4) ^ swz) + (lcol & 7) * 2;
          const float pf = __uint_as_float(static_cast<uint32_t>(lds_u16(patch_addr)) << 16);
          patch_val = __float2bfloat16_rn(fmaf(pf, f, -gcoef));
        }
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          if (col0 + h * 128 + j * 64 >= ldg) break;
          const uint32_t srow = srow0 + j * (4 * kXBoxBytes);
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const float gmx = j == 0 ? (gg == 0 ? gm0 : gm1) : (gg == 0 ? gm2 : gm3);
            const float f = x_exp2((gmx - M) * sl2) * tc;
            const __nv_bfloat16 fh = __float2bfloat16_rn(f);
            const __nv_bfloat16 fl = __float2bfloat16_rn(f - __bfloat162float(fh));
            const __nv_bfloat162 fh2 = __halves2bfloat162(fh, fh), fl2 = __halves2bfloat162(fl, fl);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t addr = srow + ((static_cast<uint32_t>(gg * 4 + k) << 4) ^ swz);
              const uint4 in = lds128(addr);
              uint32_t wi[4] = {in.x, in.y, in.z, in.w}, wo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 pv = *reinterpret_cast<const __nv_bfloat162*>(&wi[e]);
                const __nv_bfloat162 r = __hfma2(pv, fh2, __hmul2(pv, fl2));
                wo[e] = *reinterpret_cast<const uint32_t*>(&r);
              }
              sts128(addr, wo[0], wo[1], wo[2], wo[3]);
            }
          }
          if (patch && (lcol >> 6) == j) sts_u16(patch_addr, __bfloat16_as_ushort(patch_val));
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {  // the warp's 32 x 64 box leaves as ONE coalesced TMA store
            tma_store_2d_s(&tmap_g, box0 + j * (4 * kXBoxBytes), col0 + h * 128 + j * 64,
                           static_cast<int32_t>(tile * 128 + q * 32));
            bulk_commit();
          }
        }
      }
      XDBG_ACC(7);

      // ---- per-row results: the (chunk, half 0) thread that owns the row's label column ----------------
      if (h == 0) {
        const bool own = valid && lc_chunk >= 0 && lc_chunk < 256;
        float loss = 0.f, dsc = 0.f;
        int hit = 0;
        if (own) {
          loss = logf(S) + (M - lraw) * sabs;  // lraw == M for a correctly classified row: loss = log S >= 0
          if (kDs) dsc = (PR * inv_sum - lraw) * sgn * dcoef;
          hit = (lraw == M && lraw > fmaxf(before_chunks, before_loc)) ? 1 : 0;
          if (row_loss) row_loss[row] = loss;
          if (kPred && row_pred) row_pred[row] = garg;
          if (row_correct) row_correct[row] = hit;
          if (kDs && row_dscale) row_dscale[row] = dsc;
        }
        if (wk.tile_part) {
          // deterministic per-(tile, chunk, warp, run) partial sums; the statistics kernel adds them in a fixed order
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const bool mn = own && (static_cast<int>(sg) == s2);
            const float a = warp_sum(mn ? loss : 0.f), dd = kDs ? warp_sum(mn ? dsc : 0.f) : 0.f;
            const int hh = warp_sum_i(mn ? hit : 0), cnt = warp_sum_i(mn ? 1 : 0);
            if (lane == 0)
              *reinterpret_cast<float4*>(wk.tile_part + (((tile * n_chunks + chunk) * 4 + q) * 2 + s2) * 4) =
                  make_float4(a, dd, static_cast<float>(hh), static_cast<float>(cnt));
          }
        }
      }
      XDBG_ACC(8);
    }
    if (write_g && lane == 0) bulk_wait<0>();  // this warp's TMA stores have landed before the kernel ends
#ifdef UML_FWD_TIMING
    if (warp == 0 && lane == 0) XDBG_FLUSH(0, 9);
#endif
  }

  tc_fence_before();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory until the last commit
  if (warp == kXWarpAlloc) tmem_dealloc_cg2(tmem_base, 512);
  if (threadIdx.x == 0) {
    // the last CTA of the grid closes the launch: flags written with `epoch` can never match a later launch
    __threadfence();
    const unsigned done = atomicAdd(wk.ctrl + 1, 1u);
    if (done == gridDim.x - 1) {
      wk.ctrl[1] = 0u;
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(wk.ctrl) = epoch;
    }
  }
}

// per-run statistics from the per-(tile, chunk, warp) partials, summed in a fixed order; one CTA per run
__global__ void __launch_bounds__(1024)
    x_tile_stats_kernel(const float* __restrict__ tile_part, int64_t n_entries, uml_seg_stats* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[4][32];
  const int s = blockIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t i = threadIdx.x; i < n_entries; i += blockDim.x) {
    const float* p = tile_part + (i * 2 + s) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += p[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 4; ++k)
      for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) tot[k] += sh[k][w];
    out[s].loss_mean = tot[3] > 0.f ? tot[0] / tot[3] : 0.f;
    out[s].dscale = tot[1];
    out[s].correct = static_cast<int32_t>(tot[2] + 0.5f);
    out[s].n = static_cast<int32_t>(tot[3] + 0.5f);
  }
}

// the same with the number of partial records read from the workspace (uml_reduce_tile_stats knows no class count)
__global__ void __launch_bounds__(1024)
    x_tile_stats_dev_kernel(const float* __restrict__ tile_part, const unsigned* __restrict__ ctrl, uml_seg_stats* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[4][32];
  const int s = blockIdx.x;
  const int64_t n_entries = ctrl[3];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t i = threadIdx.x; i < n_entries; i += blockDim.x) {
    const float* p = tile_part + (i * 2 + s) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += p[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 4; ++k)
      for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) tot[k] += sh[k][w];
    out[s].loss_mean = tot[3] > 0.f ? tot[0] / tot[3] : 0.f;
    out[s].dscale = tot[1];
    out[s].correct = static_cast<int32_t>(tot[2] + 0.5f);
    out[s].n = static_cast<int32_t>(tot[3] + 0.5f);
  }
}

// workspace carving (UML_TILE_WS_FLOATS): [ctrl 16 u32][(unused) units*8][tile_part units*2*4*32 f][recs n_rows*4*8 f]
struct XLayout {
  unsigned* ctrl;
  float* tile_part;
  uint4* recs;
  int64_t part_entries;  // (tile, chunk, warp) partial records per run
};
static XLayout x_layout(float* tile_ws, int64_t n_rows, int n_classes) {
  const int64_t units = (n_rows + 255) / 256;
  const int n_chunks = (n_classes + 255) / 256;
  XLayout l;
  l.ctrl = reinterpret_cast<unsigned*>(tile_ws);
  l.tile_part = tile_ws + 16 + units * 8;
  l.recs = reinterpret_cast<uint4*>(l.tile_part + units * 256);
  l.part_entries = units * 2 * n_chunks * 4;
  return l;
}

}  // namespace uml

// true when the exchange kernel serves this shape (else tc_fwd.cu's chunk-sequential kernel runs)
bool uml_fwd_x_eligible(int64_t n_rows, int32_t n_classes) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("UML_FWD_IMPL");
    mode = (e && e[0] == 'o') ? 0 : 1;  // UML_FWD_IMPL=old keeps the chunk-sequential kernel + fix-up pass
  }
  return mode == 1 && n_rows > 128 && n_classes <= uml::kXMaxChunks * 256;
}

// where the per-(tile, chunk, warp) partials of the exchange kernel live and how many there are per run
void uml_fwd_x_partials(float* tile_ws, int64_t n_rows, int32_t n_classes, const float** part, int64_t* n_entries) {
  const uml::XLayout l = uml::x_layout(tile_ws, n_rows, n_classes);
  *part = l.tile_part;
  *n_entries = l.part_entries;
}

int uml_head_fwd_ce_x_bf16(const uint16_t* X, int64_t n_rows, int32_t dim, const uint16_t* W, int32_t n_classes,
                           const int32_t* labels, const uml_tc_segments* segs, uint16_t* G, int64_t ldg, float* row_loss,
                           int32_t* row_pred, int32_t* row_correct, float* row_dscale, float* tile_ws, uml_seg_stats* stats,
                           void* stream) {
  using namespace uml;
  UML_REQUIRE(X && W && labels && segs && tile_ws && n_rows > 0 && dim > 0 && n_classes > 0, "head_fwd_ce_x: bad arguments");
  UML_REQUIRE(dim % 8 == 0, "head_fwd_ce_x: dim (%d) must be a multiple of 8 (16-byte bf16 rows for TMA)", dim);
  UML_REQUIRE(n_classes <= kXMaxChunks * 256, "head_fwd_ce_x: at most %d classes", kXMaxChunks * 256);
  UML_REQUIRE(!G || (ldg % 64 == 0 && ldg >= n_classes && ldg <= kXMaxChunks * 256),
              "head_fwd_ce_x: ldg must be a multiple of 64 and >= n_classes");
  UML_REQUIRE(segs->nseg >= 1 && segs->nseg <= UML_MAX_SEGMENTS, "head_fwd_ce_x: 1..2 segments");
  const int64_t n0 = segs->seg_rows[0], n1 = segs->nseg > 1 ? segs->seg_rows[1] : 0;
  UML_REQUIRE(n0 + n1 == n_rows, "head_fwd_ce_x: segment rows (%lld+%lld) != n_rows (%lld)", (long long)n0, (long long)n1,
              (long long)n_rows);
  CUtensorMap tx, tw, tg;
  if (make_tmap_2d(&tx, X, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dim, n_rows, static_cast<uint64_t>(dim) * 2, 64, 128,
                   CU_TENSOR_MAP_SWIZZLE_128B))
    return 1;
  if (make_tmap_2d(&tw, W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dim, n_classes, static_cast<uint64_t>(dim) * 2, 64, 128,
                   CU_TENSOR_MAP_SWIZZLE_128B))
    return 1;
  memset(&tg, 0, sizeof(tg));
  if (G) {
    if (make_tmap_2d(&tg, G, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, static_cast<uint64_t>(ldg), n_rows,
                     static_cast<uint64_t>(ldg) * 2, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B))
      return 1;
  }
  XSegs fs;
  fs.n0 = segs->nseg > 1 ? n0 : INT64_MAX;
  bool learnable = false;
  for (int i = 0; i < 2; ++i) {
    const int j = i < segs->nseg ? i : 0;
    const double n = static_cast<double>(segs->seg_rows[j] > 0 ? segs->seg_rows[j] : 1);
    fs.scale[i] = segs->scale[j];
    fs.scale_dev[i] = segs->scale_dev[j];
    fs.dcoef[i] = static_cast<float>(static_cast<double>(segs->loss_weight[j]) / n);
    learnable = learnable || segs->scale_dev[j] != nullptr;
  }
  const bool ds = learnable || row_dscale != nullptr;  // the sum p * raw is only needed for d loss / d scale
  using Kern = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, int64_t, int, int, int, const int32_t*, XSegs, int, int64_t,
                        float*, int32_t*, int32_t*, float*, XWork);
  const int slot = (row_pred ? 2 : 0) + (ds ? 1 : 0);
  const Kern kerns[4] = {head_fwd_ce_x_kernel<false, false>, head_fwd_ce_x_kernel<false, true>,
                         head_fwd_ce_x_kernel<true, false>, head_fwd_ce_x_kernel<true, true>};
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[slot]) {
    UML_CUDA(cudaFuncSetAttribute(kerns[slot], cudaFuncAttributeMaxDynamicSharedMemorySize, kXSmemBytes));
    attr_set[slot] = true;
  }
  const int n_chunks = (n_classes + 255) / 256;
  const int64_t units = (n_rows + 255) / 256;
  int64_t n_groups = (sm_count() / 2) / n_chunks;
  if (n_groups > units) n_groups = units;
  UML_REQUIRE(n_groups >= 1, "head_fwd_ce_x: the device has too few SMs for %d class chunks", n_chunks);
  const XLayout l = x_layout(tile_ws, n_rows, n_classes);
  XWork wk;
  wk.tile_part = l.tile_part;
  wk.recs = l.recs;
  wk.ctrl = l.ctrl;
  const dim3 grid(static_cast<unsigned>(n_groups * n_chunks * 2));
  UML_CUDA(launch_kernel(kerns[slot], grid, dim3(kXThreads), kXSmemBytes, as_stream(stream), 2, true, tx, tw, tg, n_rows,
                         static_cast<int>(dim), static_cast<int>(n_classes), static_cast<int>(n_groups), labels, fs,
                         G ? 1 : 0, ldg, row_loss, row_pred, row_correct, row_dscale, wk));
  if (stats)
    UML_CUDA(launch_kernel(x_tile_stats_kernel, dim3(segs->nseg), dim3(1024), 0, as_stream(stream), 1, true,
                           static_cast<const float*>(l.tile_part), l.part_entries, stats));
  return 0;
}

// per-run statistics from the partials of the last exchange-kernel launch on this workspace (record count read from it)
int uml_fwd_x_reduce_stats(float* tile_ws, int64_t n_rows, int32_t nseg, uml_seg_stats* stats, void* stream) {
  using namespace uml;
  const XLayout l = x_layout(tile_ws, n_rows, 1);  // (the offsets do not depend on the class count)
  UML_CUDA(launch_kernel(x_tile_stats_dev_kernel, dim3(nseg), dim3(1024), 0, as_stream(stream), 1, true,
                         static_cast<const float*>(l.tile_part), static_cast<const unsigned*>(l.ctrl), stats));
  return 0;
}

extern "C" {

#ifdef UML_FWD_TIMING
int uml_debug_fwdx_timing(long long* host_out /* [148*16] */, int reset) {
  if (reset) {
    static long long zeros[148 * 16];
    return cudaMemcpyToSymbol(uml::g_fwdx_dbg, zeros, sizeof(zeros)) != cudaSuccess;
  }
  return cudaMemcpyFromSymbol(host_out, uml::g_fwdx_dbg, sizeof(long long) * 148 * 16) != cudaSuccess;
}
#endif

// 1 when a launch of the exchange forward kernel ever timed out waiting for a peer CTA pair (results of that launch
// are undefined); the workspace it used is passed in.  Synchronises the device.
int uml_fwd_x_failed(const float* tile_ws) {
  unsigned v = 0;
  if (cudaMemcpy(&v, reinterpret_cast<const unsigned*>(tile_ws) + 2, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v != 0 ? 1 : 0;
}

}  // extern "C"
