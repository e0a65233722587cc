// K2+K3 on the tensor cores: head forward, logit scale, softmax cross-entropy and the logit
// gradient G in ONE kernel (reference: engine/models/head.py:80-82,133-135 + finetune.py:186-188
// + the autograd of F.cross_entropy).
//
//   logits[b,c] = scale_b * sum_d X[b,d] W[c,d]         bf16 x bf16 -> fp32 in TMEM
//   G[b,c]      = w_b * scale_b / n_b * (softmax(logits)[b,c] - [c == y_b])     written as bf16
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0   TMA producer : X tile 128x64 + W chunk 256x64 per stage, 4 stages, SWIZZLE_128B
//   warp 1   MMA issuer   : tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16, accumulators in
//                           TMEM; two 256-column accumulator buffers so the MMA of class-chunk j+1
//                           overlaps the epilogue of chunk j
//   warp 2   TMEM allocator
//   warps 4-7 epilogue    : one thread per row (TMEM lane).  Per chunk: tcgen05.ld, scale, running
//                           row max / sum (online softmax), argmax, label logit, and the
//                           unnormalised probabilities exp(l - m_running) go out as bf16.  After the
//                           last chunk the row's final max/sum are known; the thread re-reads its own
//                           512-byte row segments (still L2 resident - they were written microseconds
//                           ago), rescales them to w*s/n*(p - onehot) and writes them back.  Logits
//                           never exist in HBM in fp32 and nothing is recomputed.
// A row of 1000 classes needs 1000 fp32 TMEM columns, twice what an SM has, which is why the
// normalisation is deferred instead of holding the row in TMEM.
#include "common.cuh"

namespace uml {

constexpr int kFwdBlockM = 128;
constexpr int kFwdBlockN = 256;
constexpr int kFwdBlockK = 64;
constexpr int kFwdStages = 4;
constexpr int kFwdABytes = kFwdBlockM * kFwdBlockK * 2;
constexpr int kFwdBBytes = kFwdBlockN * kFwdBlockK * 2;
constexpr int kFwdStageBytes = kFwdABytes + kFwdBBytes;
constexpr int kFwdMaxChunks = 8;  // up to 2048 classes
constexpr int kFacStride = kFwdMaxChunks + 1;  // per row: one factor per class chunk + the onehot coefficient
constexpr int kFwdStoreBox = 32 * 128;          // 32 rows x 64 bf16 columns, SWIZZLE_128B
constexpr int kFwdStoreBytes = 4 * 2 * kFwdStoreBox;  // 4 epilogue warps x double buffer
constexpr int kFwdSmemBytes = kFwdStages * kFwdStageBytes + kFwdStoreBytes + 1024 /*align*/ + 256 /*barriers*/;

struct FwdSegs {
  int64_t n0;
  const float* scale_dev[2];
  float scale[2], dcoef[2];  // dcoef = w/n ; the logit-gradient coefficient is dcoef * scale
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256, 1)
    head_fwd_ce_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                          const __grid_constant__ CUtensorMap tmap_g, int64_t n_rows, int dim, int n_classes, const int32_t* __restrict__ labels, FwdSegs segs,
                          __nv_bfloat16* __restrict__ G, int64_t ldg, float* __restrict__ row_loss,
                          int32_t* __restrict__ row_pred, int32_t* __restrict__ row_correct,
                          float* __restrict__ row_dscale, float* __restrict__ fac) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* store_smem = smem + kFwdStages * kFwdStageBytes;  // 1024-aligned: stage sizes are multiples of 1024
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(store_smem + kFwdStoreBytes);
  uint64_t* empty_bar = full_bar + kFwdStages;
  uint64_t* tfull_bar = empty_bar + kFwdStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (dim + kFwdBlockK - 1) / kFwdBlockK;
  const int n_chunks = (n_classes + kFwdBlockN - 1) / kFwdBlockN;
  const int64_t n_tiles = (n_rows + kFwdBlockM - 1) / kFwdBlockM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kFwdStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int ch = 0; ch < n_chunks; ++ch) {
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const uint32_t s = it % kFwdStages, ph = (it / kFwdStages) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_arrive_expect_tx(&full_bar[s], kFwdStageBytes);
            unsigned char* a = smem + s * kFwdStageBytes;
            tma_load_2d(a, &tmap_x, &full_bar[s], kb * kFwdBlockK, static_cast<int32_t>(tile * kFwdBlockM));
            tma_load_2d(a + kFwdABytes, &tmap_w, &full_bar[s], kb * kFwdBlockK, ch * kFwdBlockN);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer --------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kFwdBlockM, kFwdBlockN, 0, 0);
      uint32_t it = 0, acc_it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int ch = 0; ch < n_chunks; ++ch, ++acc_it) {
          const uint32_t b = acc_it & 1, aph = (acc_it >> 1) & 1;
          mbar_wait(&tempty_bar[b], aph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + b * kFwdBlockN;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const uint32_t s = it % kFwdStages, ph = (it / kFwdStages) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + s * kFwdStageBytes);
            const uint32_t b_addr = a_addr + kFwdABytes;
#pragma unroll
            for (int k = 0; k < kFwdBlockK / 16; ++k) {
              // K-major, 128B swizzle: 8-row groups are 1024 B apart; a K step of 16 bf16 = 32 B
              const uint64_t da = make_smem_desc(a_addr + k * 32, 16, 1024, kLayoutSw128);
              const uint64_t db = make_smem_desc(b_addr + k * 32, 16, 1024, kLayoutSw128);
              umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
            }
            umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
          }
          umma_commit(&tfull_bar[b]);  // accumulator chunk complete
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue ----------------------------------
    const int q = warp - 4;  // TMEM lane quarter this warp may access
    constexpr float kLog2e = 1.4426950408889634f;
    uint32_t acc_it = 0, store_it = 0;
    if (lane == 0 && warp == 4) tma_prefetch_desc(&tmap_g);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row = tile * kFwdBlockM + q * 32 + lane;
      const bool valid = row < n_rows;
      const bool sg = valid && row >= segs.n0;
      const float* sdev = sg ? segs.scale_dev[1] : segs.scale_dev[0];
      const float scale = sdev ? __ldg(sdev) : (sg ? segs.scale[1] : segs.scale[0]);
      const float dcoef = sg ? segs.dcoef[1] : segs.dcoef[0];
      const float gcoef = dcoef * scale;
      const int label = valid ? labels[row] : -1;
      const bool grow = G != nullptr;
      float run_max = -INFINITY, run_sum = 0.f, run_pr = 0.f, lab_logit = 0.f, lab_raw = 0.f;
      int arg = 0;
      float chunk_max[kFwdMaxChunks];

      for (int ch = 0; ch < n_chunks; ++ch, ++acc_it) {
        const uint32_t b = acc_it & 1, aph = (acc_it >> 1) & 1;
        mbar_wait(&tfull_bar[b], aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * kFwdBlockN;
        const int col0 = ch * kFwdBlockN;
        // sub-pass A: running max / argmax (strict > keeps the first maximal index) and label logit
        const float old_max = run_max;
#pragma unroll 1
        for (int cb = 0; cb < kFwdBlockN / 32; ++cb) {
          uint32_t v[32];
          tmem_ld32(taddr + cb * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = col0 + cb * 32 + i;
            const float raw = __uint_as_float(v[i]);
            const float x = raw * scale;
            if (c < n_classes) {
              if (x > run_max) { run_max = x; arg = c; }
              if (c == label) { lab_logit = x; lab_raw = raw; }
            }
          }
        }
        // rescale the running sums to the new maximum (exp2(-inf) = 0 on the first chunk)
        const float new_max = run_max;
        const float resc = fast_exp2((old_max - new_max) * kLog2e);
        run_sum *= resc;
        run_pr *= resc;
        chunk_max[ch] = new_max;
        // sub-pass B: exp, sums, bf16 store of exp(x - m_running)
        const float mneg = -new_max * kLog2e;
#pragma unroll 1
        for (int cb = 0; cb < kFwdBlockN / 32; ++cb) {
          uint32_t v[32];
          tmem_ld32(taddr + cb * 32, v);
          tmem_ld_wait();
          uint32_t packed[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int c = col0 + cb * 32 + i;
            const float r0 = __uint_as_float(v[i]), r1 = __uint_as_float(v[i + 1]);
            float p0 = fast_exp2(fmaf(r0 * scale, kLog2e, mneg));
            float p1 = fast_exp2(fmaf(r1 * scale, kLog2e, mneg));
            if (c >= n_classes) p0 = 0.f;
            if (c + 1 >= n_classes) p1 = 0.f;
            run_sum += p0 + p1;
            run_pr = fmaf(p0, r0, fmaf(p1, r1, run_pr));
            __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
            packed[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
          }
          if (G) {
            // stage the 32-row x 32-column piece in shared memory in the TMA 128B-swizzled layout
            // (16-byte chunk j of row r lives at chunk j ^ (r & 7)); every second column block the
            // warp's 32 x 64 tile goes out as ONE coalesced TMA store.
            unsigned char* sbuf = store_smem + (q * 2 + (store_it & 1)) * kFwdStoreBox;
            if ((cb & 1) == 0) {
              if (lane == 0) bulk_wait_read<1>();  // the store that last read this buffer has drained it
              __syncwarp();
            }
            unsigned char* srow = sbuf + lane * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int chunk = ((cb & 1) * 4 + j) ^ (lane & 7);
              *reinterpret_cast<uint4*>(srow + chunk * 16) =
                  make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
            if (cb & 1) {
              fence_proxy_async();
              __syncwarp();
              const int c = col0 + (cb - 1) * 32;
              if (lane == 0 && c < ldg) {
                tma_store_2d(&tmap_g, sbuf, c, static_cast<int32_t>(tile * kFwdBlockM + q * 32));
                bulk_commit();
              } else if (lane == 0) {
                bulk_commit();  // keep one group per buffer use so wait_group.read<1> stays exact
              }
              ++store_it;
            }
          }
        }
        // accumulator buffer b may be overwritten by the MMA warp now
        tc_fence_before();
        mbar_arrive(&tempty_bar[b]);
      }

      if (valid) {
        const float inv_sum = 1.f / run_sum;
        row_loss[row] = logf(run_sum) - (lab_logit - run_max);
        if (row_pred) row_pred[row] = arg;
        if (row_correct) row_correct[row] = (arg == label) ? 1 : 0;
        if (row_dscale) row_dscale[row] = (run_pr * inv_sum - lab_raw) * dcoef;
        if (grow) {
          // deferred normalisation, finished by g_fixup_kernel:
          //   p = exp(x - m_j) * exp(m_j - m_final) / sum ;  G = gcoef * (p - onehot)
          float* f = fac + row * kFacStride;
          for (int ch = 0; ch < n_chunks; ++ch)
            f[ch] = fast_exp2((chunk_max[ch] - run_max) * kLog2e) * inv_sum * gcoef;
          f[kFwdMaxChunks] = gcoef;
        }
      }
    }
  }

  if (warp >= 4 && lane == 0) bulk_wait<0>();  // outstanding TMA stores read shared memory: drain before exit
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// Second half of the deferred softmax normalisation: one warp per row, 16-byte vectors, every load
// of a row in flight at once.  The rows were written moments ago by the forward kernel, so at
// training batch sizes this pass runs out of L2.
__global__ void __launch_bounds__(256)
    g_fixup_kernel(__nv_bfloat16* __restrict__ G, int64_t ldg, int64_t n_rows, const int32_t* __restrict__ labels,
                   const float* __restrict__ fac) {
  const int64_t row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const int label = labels[row];
  const float* f = fac + row * kFacStride;
  const float gcoef = f[kFwdMaxChunks];
  __nv_bfloat16* g = G + row * ldg;
  uint4 v[kFwdMaxChunks];
  const int iters = static_cast<int>((ldg + 255) / 256);
#pragma unroll
  for (int j = 0; j < kFwdMaxChunks; ++j) {
    const int c = j * 256 + lane * 8;
    if (j < iters && c < ldg) v[j] = *reinterpret_cast<const uint4*>(g + c);
  }
#pragma unroll
  for (int j = 0; j < kFwdMaxChunks; ++j) {
    const int c = j * 256 + lane * 8;
    if (j < iters && c < ldg) {
      const float fj = f[j];
      uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 p = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w[i]));
        p.x *= fj;
        p.y *= fj;
        if (c + 2 * i == label) p.x -= gcoef;
        if (c + 2 * i + 1 == label) p.y -= gcoef;
        __nv_bfloat162 h = __floats2bfloat162_rn(p.x, p.y);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(g + c) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

}  // namespace uml

extern "C" {

int uml_head_fwd_ce_bf16(const uint16_t* X, int64_t n_rows, int32_t dim, const uint16_t* W, int32_t n_classes,
                         const int32_t* labels, const uml_tc_segments* segs, uint16_t* G, int64_t ldg,
                         float* row_loss, int32_t* row_pred, int32_t* row_correct, float* row_dscale,
                         float* fac_ws, void* stream) {
  using namespace uml;
  UML_REQUIRE(X && W && labels && segs && row_loss && n_rows >= 0 && dim > 0 && n_classes > 0,
              "head_fwd_ce_bf16: bad arguments");
  UML_REQUIRE(dim % 8 == 0, "head_fwd_ce_bf16: dim (%d) must be a multiple of 8 (16-byte bf16 rows for TMA)", dim);
  UML_REQUIRE(n_classes <= kFwdMaxChunks * kFwdBlockN, "head_fwd_ce_bf16: at most %d classes", kFwdMaxChunks * kFwdBlockN);
  UML_REQUIRE(!G || (ldg % 64 == 0 && ldg >= n_classes), "head_fwd_ce_bf16: ldg must be a multiple of 64 and >= n_classes");
  UML_REQUIRE(!G || fac_ws, "head_fwd_ce_bf16: fac_ws (n_rows * UML_FAC_STRIDE floats) is required when G is written");
  UML_REQUIRE(segs->nseg >= 1 && segs->nseg <= UML_MAX_SEGMENTS, "head_fwd_ce_bf16: 1..2 segments");
  if (n_rows == 0) return 0;
  const int64_t n0 = segs->seg_rows[0], n1 = segs->nseg > 1 ? segs->seg_rows[1] : 0;
  UML_REQUIRE(n0 + n1 == n_rows, "head_fwd_ce_bf16: segment rows (%lld+%lld) != n_rows (%lld)", (long long)n0,
              (long long)n1, (long long)n_rows);
  CUtensorMap tx, tw;
  if (make_tmap_2d(&tx, X, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dim, n_rows, static_cast<uint64_t>(dim) * 2, kFwdBlockK,
                   kFwdBlockM, CU_TENSOR_MAP_SWIZZLE_128B))
    return 1;
  if (make_tmap_2d(&tw, W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dim, n_classes, static_cast<uint64_t>(dim) * 2,
                   kFwdBlockK, kFwdBlockN, CU_TENSOR_MAP_SWIZZLE_128B))
    return 1;
  CUtensorMap tg;
  memset(&tg, 0, sizeof(tg));
  if (G) {
    if (make_tmap_2d(&tg, G, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, static_cast<uint64_t>(ldg), n_rows,
                     static_cast<uint64_t>(ldg) * 2, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B))
      return 1;
  }
  FwdSegs fs;
  fs.n0 = segs->nseg > 1 ? n0 : INT64_MAX;
  for (int i = 0; i < 2; ++i) {
    const int j = i < segs->nseg ? i : 0;
    const double n = static_cast<double>(segs->seg_rows[j] > 0 ? segs->seg_rows[j] : 1);
    fs.scale[i] = segs->scale[j];
    fs.scale_dev[i] = segs->scale_dev[j];
    fs.dcoef[i] = static_cast<float>(static_cast<double>(segs->loss_weight[j]) / n);
  }
  static bool attr_set = false;
  if (!attr_set) {
    UML_CUDA(cudaFuncSetAttribute(head_fwd_ce_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    attr_set = true;
  }
  const int64_t tiles = (n_rows + kFwdBlockM - 1) / kFwdBlockM;
  const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  head_fwd_ce_tc_kernel<<<grid, 256, kFwdSmemBytes, as_stream(stream)>>>(
      tx, tw, tg, n_rows, dim, n_classes, labels, fs, reinterpret_cast<__nv_bfloat16*>(G), ldg, row_loss, row_pred,
      row_correct, row_dscale, fac_ws);
  UML_CUDA(cudaGetLastError());
  if (G) {
    g_fixup_kernel<<<static_cast<unsigned>((n_rows + 7) / 8), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<__nv_bfloat16*>(G), ldg, n_rows, labels, fac_ws);
    UML_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
