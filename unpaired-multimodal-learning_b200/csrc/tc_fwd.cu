// K2+K3 on the tensor cores: head forward, logit scale, softmax cross-entropy and the logit
// gradient G in ONE kernel (reference: engine/models/head.py:80-82,133-135 + finetune.py:186-188
// + the autograd of F.cross_entropy).
//
//   logits[b,c] = scale_b * sum_d X[b,d] W[c,d]         bf16 x bf16 -> fp32 in TMEM
//   G[b,c]      = w_b * scale_b / n_b * (softmax(logits)[b,c] - [c == y_b])     written as bf16
//
// One persistent CTA per SM, 512 threads, warp-specialised:
//   warp 0     TMA producer : X tile 128x64 + W chunk 256x64 per stage, 4 stages, SWIZZLE_128B
//   warp 1     MMA issuer   : tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16, accumulators in
//                             TMEM; two 256-column accumulator buffers so the MMA of class chunk j+1
//                             overlaps the epilogue of chunk j
//   warp 2     TMEM allocator
//   warps 4-7  epilogue     : one thread per row (TMEM lane).  Per chunk: tcgen05.ld, scale, running
//                             row max / sum (online softmax), argmax, label logit; the unnormalised
//                             probabilities exp(l - m_running) are staged in shared memory in the
//                             128B-swizzled layout and leave as coalesced TMA stores.
//   warps 8-15 normaliser   : a 1000-class fp32 row needs 1000 TMEM columns and an SM has 512, so the
//                             row cannot wait in TMEM for its final max/sum.  Instead, when a tile's
//                             last chunk is done the epilogue hands the per-row, per-chunk factors
//                             exp(m_chunk - m_final) / sum * coef to these warps through shared memory;
//                             they re-read the tile's 256 KB (written microseconds ago, L2 resident),
//                             apply the factor and the one-hot term and write the final G - one warp
//                             per row, 16 independent 16-byte loads in flight per lane - while the
//                             other warps are already working on the next tile.
// Logits never exist in HBM in fp32, nothing is recomputed, and HBM sees G once.
#include "common.cuh"

namespace uml {

constexpr int kFwdBlockM = 128;
constexpr int kFwdBlockN = 256;
constexpr int kFwdBlockK = 64;
constexpr int kFwdStages = 4;
constexpr int kFwdABytes = kFwdBlockM * kFwdBlockK * 2;
constexpr int kFwdBBytes = kFwdBlockN * kFwdBlockK * 2;
constexpr int kFwdStageBytes = kFwdABytes + kFwdBBytes;
constexpr int kFwdMaxChunks = 8;                      // up to 2048 classes
constexpr int kFwdStoreBox = 32 * 128;                // 32 rows x 64 bf16 columns, SWIZZLE_128B
constexpr int kFwdStoreBytes = 4 * kFwdStoreBox;      // one staging box per epilogue warp
constexpr int kFacFloats = kFwdMaxChunks + 2;         // per row: chunk factors, one-hot coefficient, label
constexpr int kFacBytes = 2 * kFwdBlockM * kFacFloats * 4;
constexpr int kFwdSmemBytes = kFwdStages * kFwdStageBytes + kFwdStoreBytes + kFacBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kFwdThreads = 512;
constexpr int kNormWarps = 8;  // warps 8-15

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

// kIters = 16-byte vectors per lane per G row (ceil(ldg / 256)) the normaliser is compiled for
template <int kIters>
__global__ void __launch_bounds__(kFwdThreads, 1)
    head_fwd_ce_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                          const __grid_constant__ CUtensorMap tmap_g, int64_t n_rows, int dim, int n_classes,
                          const int32_t* __restrict__ labels, FwdSegs segs, __nv_bfloat16* __restrict__ G, int64_t ldg,
                          float* __restrict__ row_loss, int32_t* __restrict__ row_pred,
                          int32_t* __restrict__ row_correct, float* __restrict__ row_dscale,
                          float* __restrict__ tile_part) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* store_smem = smem + kFwdStages * kFwdStageBytes;  // 1024-aligned (stage sizes are multiples of 1024)
  float* fac_smem = reinterpret_cast<float*>(store_smem + kFwdStoreBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(fac_smem) + kFacBytes);
  uint64_t* empty_bar = full_bar + kFwdStages;
  uint64_t* tfull_bar = empty_bar + kFwdStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* fix_full = tempty_bar + 2;
  uint64_t* fix_empty = fix_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fix_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (dim + kFwdBlockK - 1) / kFwdBlockK;
  const int n_chunks = (n_classes + kFwdBlockN - 1) / kFwdBlockN;
  const int64_t n_tiles = (n_rows + kFwdBlockM - 1) / kFwdBlockM;
  const bool write_g = G != nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    if (write_g) tma_prefetch_desc(&tmap_g);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kFwdStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 128);
      mbar_init(&fix_full[b], 4);   // one arrive per epilogue warp
      mbar_init(&fix_empty[b], kNormWarps);  // one arrive per normaliser warp
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
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------ epilogue ----------------------------------
    const int q = warp - 4;  // TMEM lane quarter this warp may access
    constexpr float kLog2e = 1.4426950408889634f;
    unsigned char* sbuf = store_smem + q * kFwdStoreBox;
    unsigned char* srow = sbuf + lane * 128;
    uint32_t acc_it = 0, tile_it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      const int64_t row = tile * kFwdBlockM + q * 32 + lane;
      const bool valid = row < n_rows;
      const bool sg = valid && row >= segs.n0;
      const float* sdev = sg ? segs.scale_dev[1] : segs.scale_dev[0];
      const float scale = sdev ? __ldg(sdev) : (sg ? segs.scale[1] : segs.scale[0]);
      const float dcoef = sg ? segs.dcoef[1] : segs.dcoef[0];
      const float gcoef = dcoef * scale;
      const int label = valid ? labels[row] : -1;
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
        // sub-pass B: exp, sums, bf16 staging of exp(x - m_running)
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
          if (write_g) {
            // 16-byte chunk j of row r sits at chunk j ^ (r & 7) (TMA SWIZZLE_128B); every second column
            // block the warp's 32 x 64 tile leaves as ONE coalesced TMA store
            if ((cb & 1) == 0) {
              if (lane == 0) bulk_wait_read<0>();  // the previous store has finished reading the box
              __syncwarp();
            }
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
              }
            }
          }
        }
        // accumulator buffer b may be overwritten by the MMA warp now
        tc_fence_before();
        mbar_arrive(&tempty_bar[b]);
      }

      // ---- row results --------------------------------------------------------------------------
      const float inv_sum = 1.f / run_sum;
      const float loss = valid ? logf(run_sum) - (lab_logit - run_max) : 0.f;
      const float dsc = valid ? (run_pr * inv_sum - lab_raw) * dcoef : 0.f;
      const int hit = (valid && arg == label) ? 1 : 0;
      if (valid) {
        if (row_loss) row_loss[row] = loss;
        if (row_pred) row_pred[row] = arg;
        if (row_correct) row_correct[row] = hit;
        if (row_dscale) row_dscale[row] = dsc;
      }
      if (tile_part) {
        // deterministic per-(tile, warp, run) partial sums; the stats kernel adds them in a fixed order
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const bool mine = valid && (static_cast<int>(sg) == s);
          const float a = warp_sum(mine ? loss : 0.f), d = warp_sum(mine ? dsc : 0.f);
          const int h = warp_sum_i(mine ? hit : 0), cnt = warp_sum_i(mine ? 1 : 0);
          if (lane == 0) {
            float* o = tile_part + (tile * 8 + q * 2 + s) * 4;
            o[0] = a; o[1] = d; o[2] = static_cast<float>(h); o[3] = static_cast<float>(cnt);
          }
        }
      }
      if (write_g) {
        // hand the row's normalisation record to the normaliser warps (double-buffered per tile)
        const uint32_t fb = tile_it & 1, fph = (tile_it >> 1) & 1;
        mbar_wait(&fix_empty[fb], fph ^ 1);
        float* f = fac_smem + (fb * kFwdBlockM + q * 32 + lane) * kFacFloats;
        for (int ch = 0; ch < n_chunks; ++ch)
          f[ch] = fast_exp2((chunk_max[ch] - run_max) * kLog2e) * inv_sum * gcoef;
        f[kFwdMaxChunks] = gcoef;
        f[kFwdMaxChunks + 1] = __int_as_float(label);
        if (lane == 0) bulk_wait<0>();  // this warp's TMA stores of the tile are complete (written, not just read)
        __threadfence();
        __syncwarp();
        if (lane == 0) mbar_arrive(&fix_full[fb]);
      }
    }
  } else if (warp >= 8 && write_g) {
    // ------------------------------------------------ normaliser --------------------------------
    const int w = warp - 8;
    uint32_t tile_it = 0;
    const int iters = static_cast<int>((ldg + 255) / 256);  // 16-byte vectors per lane per row (<= 8)
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      const uint32_t fb = tile_it & 1, fph = (tile_it >> 1) & 1;
      mbar_wait(&fix_full[fb], fph);
      const float* fbase = fac_smem + fb * kFwdBlockM * kFacFloats;
      // rows w, w+8, ...; four rows per pass keep 4 * iters (16 for 1000 classes) independent 16-byte
      // loads in flight per lane, which is what hides the L2 round trip
      constexpr int kRowsPerPass = 16 / kIters;
#pragma unroll 1
      for (int r = w; r < kFwdBlockM; r += kRowsPerPass * kNormWarps) {
        uint4 v[kRowsPerPass][kIters];
#pragma unroll
        for (int h = 0; h < kRowsPerPass; ++h) {
          const int64_t row = tile * kFwdBlockM + r + h * kNormWarps;
          const __nv_bfloat16* g = G + row * ldg;
#pragma unroll
          for (int j = 0; j < kIters; ++j) {
            const int c = j * 256 + lane * 8;
            if (j < iters && c < ldg && row < n_rows) v[h][j] = __ldcg(reinterpret_cast<const uint4*>(g + c));
          }
        }
#pragma unroll
        for (int h = 0; h < kRowsPerPass; ++h) {
          const int64_t row = tile * kFwdBlockM + r + h * kNormWarps;
          if (row >= n_rows) continue;
          const float* f = fbase + (r + h * kNormWarps) * kFacFloats;
          const float gcoef = f[kFwdMaxChunks];
          const int label = __float_as_int(f[kFwdMaxChunks + 1]);
          __nv_bfloat16* g = G + row * ldg;
#pragma unroll
          for (int j = 0; j < kIters; ++j) {
            const int c = j * 256 + lane * 8;
            if (j < iters && c < ldg) {
              const float fj = f[j];
              uint32_t wds[4] = {v[h][j].x, v[h][j].y, v[h][j].z, v[h][j].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float2 p = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&wds[i]));
                p.x *= fj;
                p.y *= fj;
                if (c + 2 * i == label) p.x -= gcoef;
                if (c + 2 * i + 1 == label) p.y -= gcoef;
                __nv_bfloat162 hh = __floats2bfloat162_rn(p.x, p.y);
                wds[i] = *reinterpret_cast<uint32_t*>(&hh);
              }
              *reinterpret_cast<uint4*>(g + c) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&fix_empty[fb]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// per-run statistics from the per-tile partials, summed in a fixed order (deterministic)
__global__ void __launch_bounds__(1024)
    tile_stats_kernel(const float* __restrict__ tile_part, int64_t n_tiles, int nseg, uml_seg_stats* __restrict__ out) {
  __shared__ float sh[4][32];
  const int s = blockIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int64_t n = n_tiles * 4;  // (tile, warp) partials of this run
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
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
  (void)nseg;
}

}  // namespace uml

extern "C" {

int uml_head_fwd_ce_bf16(const uint16_t* X, int64_t n_rows, int32_t dim, const uint16_t* W, int32_t n_classes,
                         const int32_t* labels, const uml_tc_segments* segs, uint16_t* G, int64_t ldg,
                         float* row_loss, int32_t* row_pred, int32_t* row_correct, float* row_dscale,
                         float* tile_ws, void* stream) {
  using namespace uml;
  UML_REQUIRE(X && W && labels && segs && n_rows >= 0 && dim > 0 && n_classes > 0, "head_fwd_ce_bf16: bad arguments");
  UML_REQUIRE(dim % 8 == 0, "head_fwd_ce_bf16: dim (%d) must be a multiple of 8 (16-byte bf16 rows for TMA)", dim);
  UML_REQUIRE(n_classes <= kFwdMaxChunks * kFwdBlockN, "head_fwd_ce_bf16: at most %d classes", kFwdMaxChunks * kFwdBlockN);
  UML_REQUIRE(!G || (ldg % 64 == 0 && ldg >= n_classes && ldg <= kFwdMaxChunks * kFwdBlockN),
              "head_fwd_ce_bf16: ldg must be a multiple of 64 and >= n_classes");
  UML_REQUIRE(segs->nseg >= 1 && segs->nseg <= UML_MAX_SEGMENTS, "head_fwd_ce_bf16: 1..2 segments");
  if (n_rows == 0) return 0;
  const int64_t n0 = segs->seg_rows[0], n1 = segs->nseg > 1 ? segs->seg_rows[1] : 0;
  UML_REQUIRE(n0 + n1 == n_rows, "head_fwd_ce_bf16: segment rows (%lld+%lld) != n_rows (%lld)", (long long)n0,
              (long long)n1, (long long)n_rows);
  CUtensorMap tx, tw, tg;
  if (make_tmap_2d(&tx, X, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dim, n_rows, static_cast<uint64_t>(dim) * 2, kFwdBlockK,
                   kFwdBlockM, CU_TENSOR_MAP_SWIZZLE_128B))
    return 1;
  if (make_tmap_2d(&tw, W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dim, n_classes, static_cast<uint64_t>(dim) * 2,
                   kFwdBlockK, kFwdBlockN, CU_TENSOR_MAP_SWIZZLE_128B))
    return 1;
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
    UML_CUDA(cudaFuncSetAttribute(head_fwd_ce_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    UML_CUDA(cudaFuncSetAttribute(head_fwd_ce_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    attr_set = true;
  }
  const int64_t tiles = (n_rows + kFwdBlockM - 1) / kFwdBlockM;
  const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  auto kern = (!G || ldg <= 4 * 256) ? head_fwd_ce_tc_kernel<4> : head_fwd_ce_tc_kernel<8>;
  kern<<<grid, kFwdThreads, kFwdSmemBytes, as_stream(stream)>>>(tx, tw, tg, n_rows, dim, n_classes, labels, fs,
                                                               reinterpret_cast<__nv_bfloat16*>(G), ldg, row_loss,
                                                               row_pred, row_correct, row_dscale, tile_ws);
  UML_CUDA(cudaGetLastError());
  return 0;
}

int uml_reduce_tile_stats(const float* tile_ws, int64_t n_rows, int32_t nseg, uml_seg_stats* stats, void* stream) {
  using namespace uml;
  UML_REQUIRE(tile_ws && stats && nseg >= 1 && nseg <= UML_MAX_SEGMENTS && n_rows >= 0, "reduce_tile_stats: bad arguments");
  const int64_t tiles = (n_rows + kFwdBlockM - 1) / kFwdBlockM;
  tile_stats_kernel<<<nseg, 1024, 0, as_stream(stream)>>>(tile_ws, tiles, nseg, stats);
  UML_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
