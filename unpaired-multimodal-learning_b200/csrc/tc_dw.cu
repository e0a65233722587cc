// K4 on the tensor cores: dW = G^T X  (autograd of the shared head; finetune.py:190-193).
//
//   dW[c,d] = sum_b G[b,c] X[b,d]      M = classes, N = feature dim, K = batch rows
//
// Both operands are stored with K (the batch row) as the SLOW dimension - G[b, :] and X[b, :] are
// rows - so they are fed to tcgen05.mma as MN-major operands straight from TMA tiles; nothing is
// transposed in HBM.  The output is only C x D (1000 x 768), i.e. 24 tiles of 128 x 256, far fewer
// than 148 SMs, so the batch dimension is split across CTAs (split-K) and each CTA writes an fp32
// partial; the partials are summed inside the fused optimizer kernel (optim.cu), in a fixed order, so
// the result is deterministic and dW never makes a separate trip through HBM.
//
// One CTA = one (class tile, dim tile, K split).  warp 0 TMA producer (2 G boxes + 4 X boxes of
// 64 rows x 128 B per stage, SWIZZLE_128B), warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7
// epilogue (tcgen05.ld -> 128-byte fp32 row segments).
#include "common.cuh"

namespace uml {

constexpr int kDwBlockM = 128;   // classes per tile
constexpr int kDwBlockN = 256;   // feature dims per tile
constexpr int kDwBlockK = 64;    // batch rows per stage
constexpr int kDwStages = 4;
constexpr int kDwBoxBytes = 64 * kDwBlockK * 2;                 // 64 elements x 64 rows of bf16
constexpr int kDwABytes = (kDwBlockM / 64) * kDwBoxBytes;       // 16 KB
constexpr int kDwBBytes = (kDwBlockN / 64) * kDwBoxBytes;       // 32 KB
constexpr int kDwStageBytes = kDwABytes + kDwBBytes;
constexpr int kDwSmemBytes = kDwStages * kDwStageBytes + 1024 + 256;

__global__ void __launch_bounds__(256, 1)
    head_bwd_dw_tc_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                          int64_t n_rows, int dim, int n_classes, int n_splits, float* __restrict__ partials) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kDwStages * kDwStageBytes);
  uint64_t* empty_bar = full_bar + kDwStages;
  uint64_t* tfull_bar = empty_bar + kDwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, n_tile = blockIdx.y, split = blockIdx.z;
  const int num_kb = static_cast<int>((n_rows + kDwBlockK - 1) / kDwBlockK);
  // contiguous, balanced k-block ranges
  const int kb_lo = static_cast<int>((static_cast<int64_t>(num_kb) * split) / n_splits);
  const int kb_hi = static_cast<int>((static_cast<int64_t>(num_kb) * (split + 1)) / n_splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_g);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kDwBlockN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
        const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], kDwStageBytes);
        unsigned char* a = smem + s * kDwStageBytes;
#pragma unroll
        for (int j = 0; j < kDwBlockM / 64; ++j)
          tma_load_2d(a + j * kDwBoxBytes, &tmap_g, &full_bar[s], m_tile * kDwBlockM + j * 64, kb * kDwBlockK);
#pragma unroll
        for (int j = 0; j < kDwBlockN / 64; ++j)
          tma_load_2d(a + kDwABytes + j * kDwBoxBytes, &tmap_x, &full_bar[s], n_tile * kDwBlockN + j * 64,
                      kb * kDwBlockK);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kDwBlockM, kDwBlockN, 1, 1);  // both operands MN-major
      uint32_t it = 0;
      for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
        const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * kDwStageBytes);
        const uint32_t b_addr = a_addr + kDwABytes;
#pragma unroll
        for (int k = 0; k < kDwBlockK / 16; ++k) {
          // MN-major, 128B swizzle: a row (one k) holds 64 MN elements = 128 B; 8-row groups are
          // 1024 B apart (SBO); the next 64 MN elements live in the next TMA box (LBO = box bytes);
          // a K step of 16 rows advances the start address by 16 * 128 B.
          const uint64_t da = make_smem_desc(a_addr + k * 2048, kDwBoxBytes, 1024, kLayoutSw128);
          const uint64_t db = make_smem_desc(b_addr + k * 2048, kDwBoxBytes, 1024, kLayoutSw128);
          umma_bf16(tmem_base, da, db, idesc, (it | k) != 0);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tfull_bar);
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int c = m_tile * kDwBlockM + q * 32 + lane;
    float* out = partials + (static_cast<int64_t>(split) * n_classes + c) * dim;
    const bool have = kb_hi > kb_lo;
    if (have) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int cb = 0; cb < kDwBlockN / 32; ++cb) {
      uint32_t v[32];
      if (have) {
        tmem_ld32(taddr + cb * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      const int d0 = n_tile * kDwBlockN + cb * 32;
      if (c < n_classes) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          if (d0 + i + 4 <= dim) {
            *reinterpret_cast<uint4*>(out + d0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
            for (int j = 0; j < 4; ++j)
              if (d0 + i + j < dim) out[d0 + i + j] = __uint_as_float(v[i + j]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kDwBlockN);
}

static int dw_splits(int64_t n_rows, int32_t dim, int32_t n_classes) {
  const int64_t tiles = static_cast<int64_t>((n_classes + kDwBlockM - 1) / kDwBlockM) * ((dim + kDwBlockN - 1) / kDwBlockN);
  const int64_t num_kb = (n_rows + kDwBlockK - 1) / kDwBlockK;
  int64_t s = sm_count() / (tiles > 0 ? tiles : 1);
  if (s < 1) s = 1;
  if (s > num_kb) s = num_kb;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

}  // namespace uml

extern "C" {

int uml_tc_dw_splits(int64_t n_rows, int32_t dim, int32_t n_classes) { return uml::dw_splits(n_rows, dim, n_classes); }

int uml_head_bwd_dw_bf16(const uint16_t* G, int64_t ldg, const uint16_t* X, int64_t n_rows, int32_t dim,
                         int32_t n_classes, float* partials, int32_t n_splits, void* stream) {
  using namespace uml;
  UML_REQUIRE(G && X && partials && n_rows > 0 && dim > 0 && n_classes > 0 && n_splits >= 1,
              "head_bwd_dw_bf16: bad arguments");
  UML_REQUIRE(dim % 8 == 0 && ldg % 64 == 0 && ldg >= n_classes,
              "head_bwd_dw_bf16: dim must be a multiple of 8 and ldg a multiple of 64 >= n_classes");
  UML_REQUIRE((reinterpret_cast<uintptr_t>(partials) & 15u) == 0 && dim % 4 == 0, "head_bwd_dw_bf16: partials alignment");
  const int64_t num_kb = (n_rows + kDwBlockK - 1) / kDwBlockK;
  UML_REQUIRE(n_splits <= num_kb, "head_bwd_dw_bf16: n_splits (%d) exceeds the %lld k-blocks", n_splits, (long long)num_kb);
  CUtensorMap tg, tx;
  // inner dimension = class / feature index (contiguous), outer = batch row (K)
  if (make_tmap_2d(&tg, G, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, static_cast<uint64_t>(ldg), n_rows,
                   static_cast<uint64_t>(ldg) * 2, 64, kDwBlockK, CU_TENSOR_MAP_SWIZZLE_128B))
    return 1;
  if (make_tmap_2d(&tx, X, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, static_cast<uint64_t>(dim), n_rows,
                   static_cast<uint64_t>(dim) * 2, 64, kDwBlockK, CU_TENSOR_MAP_SWIZZLE_128B))
    return 1;
  static bool attr_set = false;
  if (!attr_set) {
    UML_CUDA(cudaFuncSetAttribute(head_bwd_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemBytes));
    attr_set = true;
  }
  dim3 grid((n_classes + kDwBlockM - 1) / kDwBlockM, (dim + kDwBlockN - 1) / kDwBlockN, n_splits);
  head_bwd_dw_tc_kernel<<<grid, 256, kDwSmemBytes, as_stream(stream)>>>(tg, tx, n_rows, dim, n_classes, n_splits,
                                                                       partials);
  UML_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
