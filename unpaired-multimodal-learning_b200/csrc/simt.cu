// fp32 exact path (SIMT FFMA): the head step at the reference's own batch sizes (8..64 rows per
// modality, engine/optimizer/default.py:8,24,39), where tensor cores cannot help and the work is
// launch/latency bound.  Three launches per step:
//   1. raw logits  L = [X_img ; X_txt] W^T          (rows gathered from the banks by index)
//   2. per-row softmax / CE / argmax, L <- G = w*s/n (softmax - onehot)
//   3. dW = G^T [X_img ; X_txt] with the AdamW/Adam/SGD update applied in the GEMM epilogue
// plus the adapter GEMMs (K5) and the streaming eval kernel (K7).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace uml {

// Rows of a logical matrix that may live in two banks and be index-gathered (image run, text run).
struct RowSrc {
  const float* base[2];
  const int64_t* idx[2];
  int64_t n0;  // rows in run 0; rows >= n0 belong to run 1
  int64_t ld[2];
  __device__ __forceinline__ const float* row(int64_t r) const {
    const bool s = r >= n0;  // ternaries (not array indexing) keep the struct in registers
    const int64_t l = s ? r - n0 : r;
    const int64_t* ix = s ? idx[1] : idx[0];
    const int64_t src = ix ? ix[l] : l;
    return (s ? base[1] : base[0]) + src * (s ? ld[1] : ld[0]);
  }
};

static RowSrc dense_src(const float* p, int64_t ld, const int64_t* idx = nullptr) {
  RowSrc r;
  r.base[0] = r.base[1] = p;
  r.idx[0] = idx;
  r.idx[1] = nullptr;
  r.n0 = INT64_MAX;
  r.ld[0] = r.ld[1] = ld;
  return r;
}

struct DeviceUpdate {
  int kind;  // 0 store, 1 adamw, 2 adam(L2), 3 sgd
  float lr, beta1, beta2, eps, wd, momentum;
  float step_size, bc2_sqrt_inv, decay;  // derived on the host in double
  int first_step;
  float* m;
  float* v;
};

static DeviceUpdate make_update(const uml_update* u) {
  DeviceUpdate d;
  memset(&d, 0, sizeof(d));
  if (!u || u->kind == 0) return d;
  d.kind = u->kind;
  d.lr = u->lr;
  d.beta1 = u->beta1;
  d.beta2 = u->beta2;
  d.eps = u->eps;
  d.wd = u->weight_decay;
  d.momentum = u->momentum;
  const double t = static_cast<double>(u->step);
  const double bc1 = 1.0 - pow(static_cast<double>(u->beta1), t);
  const double bc2 = 1.0 - pow(static_cast<double>(u->beta2), t);
  d.step_size = static_cast<float>(static_cast<double>(u->lr) / bc1);
  d.bc2_sqrt_inv = static_cast<float>(1.0 / sqrt(bc2));
  d.decay = static_cast<float>(1.0 - static_cast<double>(u->lr) * static_cast<double>(u->weight_decay));
  d.first_step = u->step <= 1;
  d.m = u->m;
  d.v = u->v;
  return d;
}

// One parameter element; the same arithmetic order as torch.optim's single-tensor rules.
__device__ __forceinline__ void update_value(const DeviceUpdate& u, float& w, float& m, float& v, float g) {
  if (u.kind == 3) {  // SGD momentum, L2 decay folded into the gradient
    g = fmaf(u.wd, w, g);
    const float b = u.first_step ? g : fmaf(u.momentum, m, g);
    m = b;
    w = w - u.lr * b;
    return;
  }
  if (u.kind == 1) w *= u.decay;            // AdamW: decoupled decay
  else if (u.wd != 0.f) g = fmaf(u.wd, w, g);  // Adam: L2
  m = m + (g - m) * (1.f - u.beta1);
  v = v * u.beta2 + (1.f - u.beta2) * g * g;
  const float denom = sqrtf(v) * u.bc2_sqrt_inv + u.eps;
  w = w - u.step_size * (m / denom);
}
__device__ __forceinline__ void apply_update(const DeviceUpdate& u, float* p, int64_t i, float g) {
  float w = p[i], m = u.m[i], v = u.kind == 3 ? 0.f : u.v[i];
  update_value(u, w, m, v, g);
  p[i] = w;
  u.m[i] = m;
  if (u.kind != 3) u.v[i] = v;
}

// -------------------------------------------------------------------------------------------------
// generic tiled SGEMM  C[m,n] = alpha * sum_k A(m,k) B(k,n)
//   A_K: A(m,k) = rowA(m)[k]   else A(m,k) = rowA(k)[m]
//   B_K: B(k,n) = rowB(n)[k]   else B(k,n) = rowB(k)[n]
// -------------------------------------------------------------------------------------------------
// kBK: depth of a staged k-tile.  The small-tile instantiation (reference batch sizes: a 64 x 1000 x 512 forward
// is 64 CTAs of 32 x 32) is bound by the latency of one dependent global-load -> shared -> sync round per k-tile,
// not by FMAs, so it stages 64 deep (8 rounds instead of 32 for D = 512).
template <int BM, int BN, bool A_K, bool B_K, int kBK = 16>
__global__ void __launch_bounds__(256)
    sgemm_kernel(RowSrc A, RowSrc B, float* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int64_t K, float alpha,
                 float* __restrict__ P, DeviceUpdate upd, int64_t plane) {
  constexpr int TM = BM / 16, TN = BN / 16;
  // split-K over gridDim.z: split z sums k in [k_lo, k_hi) into its own output plane (C + z * plane); the consumer
  // adds the planes in a fixed order.  Used by the small-batch forward, where 64 CTAs would leave half the SMs idle.
  const int64_t k_tiles = (K + kBK - 1) / kBK;
  const int64_t k_lo = (k_tiles * blockIdx.z / gridDim.z) * kBK;
  const int64_t k_hi_raw = (k_tiles * (blockIdx.z + 1) / gridDim.z) * kBK;
  const int64_t k_hi = k_hi_raw < K ? k_hi_raw : K;
  if (C) C += blockIdx.z * plane;
  __shared__ float As[kBK][BM + 4];
  __shared__ float Bs[kBK][BN + 4];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * BM, n0 = static_cast<int64_t>(blockIdx.x) * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = k_lo; k0 < k_hi; k0 += kBK) {
    // ---- stage A tile: BM x kBK ----
    if (A_K) {
      for (int e = t; e < BM * kBK; e += 256) {
        const int m = e / kBK, k = e % kBK;
        float x = 0.f;
        if (m0 + m < M && k0 + k < k_hi) x = A.row(m0 + m)[k0 + k];
        As[k][m] = x;
      }
    } else {
      for (int e = t; e < BM * kBK; e += 256) {
        const int k = e / BM, m = e % BM;
        float x = 0.f;
        if (m0 + m < M && k0 + k < k_hi) x = A.row(k0 + k)[m0 + m];
        As[k][m] = x;
      }
    }
    if (B_K) {
      for (int e = t; e < BN * kBK; e += 256) {
        const int n = e / kBK, k = e % kBK;
        float x = 0.f;
        if (n0 + n < N && k0 + k < k_hi) x = B.row(n0 + n)[k0 + k];
        Bs[k][n] = x;
      }
    } else {
      for (int e = t; e < BN * kBK; e += 256) {
        const int k = e / BN, n = e % BN;
        float x = 0.f;
        if (n0 + n < N && k0 + k < k_hi) x = B.row(k0 + k)[n0 + n];
        Bs[k][n] = x;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx * TN + j;
      if (n >= N) continue;
      const float g = alpha * acc[i][j];
      if (C) C[m * ldc + n] = g;
      if (upd.kind) apply_update(upd, P, m * ldc + n, g);
    }
  }
}

template <bool A_K, bool B_K>
static int launch_sgemm(const RowSrc& A, const RowSrc& B, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                        float alpha, float* P, const DeviceUpdate& upd, cudaStream_t st, int splits = 1,
                        int64_t plane = 0) {
  if (M <= 0 || N <= 0) return 0;
  UML_REQUIRE(splits >= 1 && (splits == 1 || (upd.kind == 0 && C && plane >= M * ldc)), "sgemm: bad split-K arguments");
  const int64_t ctas64 = ((M + 63) / 64) * ((N + 63) / 64);
  if (ctas64 >= 2 * sm_count()) {
    dim3 grid(static_cast<unsigned>((N + 63) / 64), static_cast<unsigned>((M + 63) / 64), static_cast<unsigned>(splits));
    sgemm_kernel<64, 64, A_K, B_K><<<grid, 256, 0, st>>>(A, B, C, ldc, M, N, K, alpha, P, upd, plane);
  } else {
    dim3 grid(static_cast<unsigned>((N + 31) / 32), static_cast<unsigned>((M + 31) / 32), static_cast<unsigned>(splits));
    sgemm_kernel<32, 32, A_K, B_K, 64><<<grid, 256, 0, st>>>(A, B, C, ldc, M, N, K, alpha, P, upd, plane);
  }
  UML_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// per-row softmax / cross entropy / argmax; rewrites the raw logits row into G
// -------------------------------------------------------------------------------------------------
struct SegInfo {
  int64_t n0, n1;
  const int64_t* idx[2];
  const int64_t* labels[2];
  const float* scale_dev[2];
  float scale[2], weight[2];
};

__device__ __forceinline__ float block_reduce_max(float v, float* sh) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = fmaxf(r, sh[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) r += sh[w];
  __syncthreads();
  return r;
}

// one row by one CTA of 256 threads: raw logits row -> G row, row loss / hit / d loss / d scale
__device__ __forceinline__ void softmax_ce_grad_row(int64_t r, float* __restrict__ L, int64_t ldl, int C, const SegInfo& seg,
                                                    float* __restrict__ row_loss, int32_t* __restrict__ row_correct,
                                                    float* __restrict__ row_dscale, float* sh, int& sh_arg) {
  const bool s = r >= seg.n0;
  const int64_t l = s ? r - seg.n0 : r;
  const int64_t n_seg = s ? seg.n1 : seg.n0;
  const int64_t* ix = s ? seg.idx[1] : seg.idx[0];
  const int64_t src = ix ? ix[l] : l;
  const int label = static_cast<int>((s ? seg.labels[1] : seg.labels[0])[src]);
  const float* sdev = s ? seg.scale_dev[1] : seg.scale_dev[0];
  const float scale = sdev ? *sdev : (s ? seg.scale[1] : seg.scale[0]);
  const float weight = s ? seg.weight[1] : seg.weight[0];
  float* row = L + r * ldl;
  const float label_raw = row[label];  // read before the row is overwritten with G

  float mx = -INFINITY;
  int arg = INT_MAX;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float x = row[c] * scale;
    if (x > mx) { mx = x; arg = c; }
  }
  if (threadIdx.x == 0) sh_arg = INT_MAX;
  const float bmax = block_reduce_max(mx, sh);  // contains the barriers that publish sh_arg
  // first maximal index, like torch.argmax on a contiguous row
  if (mx == bmax) atomicMin(&sh_arg, arg);
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += expf(row[c] * scale - bmax);
  const float sum = block_reduce_sum(se, sh);
  const float inv = 1.f / sum;
  const float gcoef = weight * scale / static_cast<float>(n_seg);
  float ds = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float raw = row[c];
    float p = expf(raw * scale - bmax) * inv;
    if (c == label) p -= 1.f;
    ds = fmaf(p, raw, ds);
    row[c] = p * gcoef;
  }
  const float dsum = block_reduce_sum(ds, sh);
  if (threadIdx.x == 0) {
    row_loss[r] = logf(sum) - (label_raw * scale - bmax);  // log_softmax form: exact 0 for a dominant label
    row_correct[r] = (sh_arg == label) ? 1 : 0;
    row_dscale[r] = dsum * weight / static_cast<float>(n_seg);
  }
}

__global__ void __launch_bounds__(256)
    softmax_ce_grad_kernel(float* __restrict__ L, int64_t ldl, int C, SegInfo seg, float* __restrict__ row_loss,
                           int32_t* __restrict__ row_correct, float* __restrict__ row_dscale, int planes, int64_t plane) {
  __shared__ float sh[8];
  __shared__ int sh_arg;
  const int64_t r = blockIdx.x;
  if (planes > 1) {  // split-K forward: raw logits = sum of the planes, in plane order
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float x = L[r * ldl + c];
      for (int z = 1; z < planes; ++z) x += L[z * plane + r * ldl + c];
      L[r * ldl + c] = x;
    }
    __syncthreads();
  }
  softmax_ce_grad_row(r, L, ldl, C, seg, row_loss, row_correct, row_dscale, sh, sh_arg);
}

// -------------------------------------------------------------------------------------------------
// The whole step of the reference's own batch sizes in ONE cooperative launch (one CTA per SM):
// at 32 + 32 rows the four launches above spend their time starting, draining and re-reading
// (56 us per step); here the step's rows stay in shared memory from the logits to the gradient,
// every CTA owns a contiguous range of classes for both contractions, and the three phases are
// separated by two grid-wide barriers:
//   1. raw logits of the CTA's classes for all rows            (weight rows streamed once)
//   2. softmax / CE / argmax, one row per CTA, G written back  (the row kernel's arithmetic)
//   3. dW of the CTA's classes (rows summed in order, like the GEMM) + optimizer update, and the
//      per-run statistics by CTA 0
// Needs 16-byte aligned rows and R x D floats of shared memory; anything else takes the launches above.
// -------------------------------------------------------------------------------------------------
constexpr int kFusedMaxRows = 128;
// 16-byte asynchronous global -> shared copy (LDGSTS)
__device__ __forceinline__ void ldgsts16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
struct FusedStepArgs {
  RowSrc X;
  SegInfo seg;
  float* W;
  DeviceUpdate upd;
  float* G;
  int64_t ldg;
  int C, D, R, nseg;
  float* row_loss;
  int32_t* row_correct;
  float* row_dscale;
  uml_seg_stats* stats;
};

__global__ void __launch_bounds__(256, 1) head_step_fused_kernel(const __grid_constant__ FusedStepArgs a) {
  extern __shared__ __align__(16) float fused_smem[];
  __shared__ float sh[8];
  __shared__ int sh_arg;
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int C = a.C, D = a.D, R = a.R, D4 = D >> 2;
  const int c_lo = static_cast<int>(static_cast<int64_t>(C) * blockIdx.x / gridDim.x);
  const int c_hi = static_cast<int>(static_cast<int64_t>(C) * (blockIdx.x + 1) / gridDim.x);
  const int nc = c_hi - c_lo;
  // rows of the step, gathered from the banks by index; row stride D4 + 4 float4: the eight lanes of a quarter warp in
  // phase 1 - two rows x four interleaved k-lanes - then cover all 32 banks exactly once
  const int XS = D4 + 4;
  float4* xs = reinterpret_cast<float4*>(fused_smem);  // [R][XS]
  __shared__ const float4* rowp[kFusedMaxRows];
  for (int r = t; r < R; r += 256) rowp[r] = reinterpret_cast<const float4*>(a.X.row(r));
  __syncthreads();
  // the CTA's weight rows (contiguous in HBM) come along: read from HBM exactly once, by all threads at once - a thread
  // walking its row 16 bytes at a time straight from HBM made phase 1 a 30 us pointer chase
  float4* wsm = xs + static_cast<size_t>(R) * XS;  // [nc][D4]
  {
    // the optimizer state of the CTA's classes (contiguous, first touched in phase 3) is asked into L2 now
    if (t < 2 && nc > 0) {
      const float* st = t == 0 ? a.upd.m : a.upd.v;
      if (st != nullptr)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(st + static_cast<int64_t>(c_lo) * D),
                     "r"(static_cast<uint32_t>(nc) * static_cast<uint32_t>(D) * 4u)
                     : "memory");
    }
    // both tiles go global -> shared without passing through registers (cp.async): every 16-byte piece of the CTA's share is
    // in flight at once instead of eight per thread
    const float4* wsrc = reinterpret_cast<const float4*>(a.W + static_cast<int64_t>(c_lo) * D);
    const int nw = nc * D4;
    for (int e = t; e < nw; e += 256) ldgsts16(wsm + e, wsrc + e);
    // every CTA needs the same R rows: each starts somewhere else in them, or all SMs would ask the same L2 lines at the
    // same moment (measured: the in-step walk took 15 us for 128 KB per CTA, 10 us staggered)
    const int n = R * D4;
    const int rot = static_cast<int>(static_cast<int64_t>(n) * blockIdx.x / gridDim.x);
    for (int e0 = t; e0 < n; e0 += 256) {
      int e = e0 + rot;
      if (e >= n) e -= n;
      const int r = e / D4, q = e - r * D4;
      ldgsts16(xs + r * XS + q, rowp[r] + q);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  // ---- 1. raw logits: thread = (row, k-lane g); lane g walks the 16-byte chunks q = g (mod 4) of the row for up to eight
  //         classes at a time (x from shared memory once per eight classes), the four lanes of a row meet by shuffle ----
  {
    const int g = t & 3;
    for (int r = t >> 2; r < ((R + 63) & ~63); r += 64) {  // (whole warps stay in the loop: the shuffles below need them)
      const float4* xrow = xs + min(r, R - 1) * XS;
      for (int j0 = 0; j0 < nc; j0 += 8) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int q = g; q < D4; q += 4) {
          const float4 x = xrow[q];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 w = wsm[min(j0 + k, nc - 1) * D4 + q];  // (a class past the CTA's range repeats the last one, unused)
            float c = acc[k];
            c = fmaf(x.x, w.x, c);
            c = fmaf(x.y, w.y, c);
            c = fmaf(x.z, w.z, c);
            c = fmaf(x.w, w.w, c);
            acc[k] = c;
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
          acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
        }
        if (r < R) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if ((k & 3) == g && j0 + k < nc) a.G[static_cast<int64_t>(r) * a.ldg + c_lo + j0 + k] = acc[k];
        }
      }
    }
  }
  grid.sync();
  // ---- 2. rows -------------------------------------------------------------------------------------
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    softmax_ce_grad_row(r, a.G, a.ldg, C, a.seg, a.row_loss, a.row_correct, a.row_dscale, sh, sh_arg);
    __syncthreads();
  }
  grid.sync();
  // ---- 3. dW of this CTA's classes + update: thread = (four classes, one 16-byte chunk of the dim); the rows are summed
  //         in order, one FFMA per term, like the GEMM -----------------------------------------------------
  const int NCP = (nc + 3) & ~3;
  float* gs = reinterpret_cast<float*>(wsm + static_cast<size_t>(nc) * D4);  // [R][NCP] the CTA's columns of G
  for (int e = t; e < R * NCP; e += 256) {
    const int r = e / NCP, j = e - r * NCP;
    // (written by other CTAs in this launch: not through the read-only path)
    gs[e] = j < nc ? __ldcg(a.G + static_cast<int64_t>(r) * a.ldg + c_lo + j) : 0.f;
  }
  __syncthreads();
  for (int it = t; it < (NCP >> 2) * D4; it += 256) {
    const int jb = (it / D4) << 2, q = it % D4;
    float4 w[4], m[4], v[4], acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[k] = w[k] = m[k] = v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (jb + k < nc) {
        const int64_t off = (static_cast<int64_t>(c_lo + jb + k) * D) + 4 * q;
        w[k] = wsm[(jb + k) * D4 + q];  // (= W[c_lo + j][4 q ..]: nobody has written it since phase 1)
        m[k] = *reinterpret_cast<const float4*>(a.upd.m + off);
        if (a.upd.kind != 3) v[k] = *reinterpret_cast<const float4*>(a.upd.v + off);
      }
    }
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
      const float4 g4 = *reinterpret_cast<const float4*>(gs + r * NCP + jb);
      const float4 x = xs[r * XS + q];
      const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[k].x = fmaf(g[k], x.x, acc[k].x);
        acc[k].y = fmaf(g[k], x.y, acc[k].y);
        acc[k].z = fmaf(g[k], x.z, acc[k].z);
        acc[k].w = fmaf(g[k], x.w, acc[k].w);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (jb + k >= nc) continue;
      const int64_t off = (static_cast<int64_t>(c_lo + jb + k) * D) + 4 * q;
      update_value(a.upd, w[k].x, m[k].x, v[k].x, acc[k].x);
      update_value(a.upd, w[k].y, m[k].y, v[k].y, acc[k].y);
      update_value(a.upd, w[k].z, m[k].z, v[k].z, acc[k].z);
      update_value(a.upd, w[k].w, m[k].w, v[k].w, acc[k].w);
      *reinterpret_cast<float4*>(a.W + off) = w[k];
      *reinterpret_cast<float4*>(a.upd.m + off) = m[k];
      if (a.upd.kind != 3) *reinterpret_cast<float4*>(a.upd.v + off) = v[k];
    }
  }
  // per-run statistics, fixed order: warp s of CTA 0 sums run s (lanes stride over the rows, shuffle tree)
  if (blockIdx.x == 0 && warp < a.nseg) {
    const int64_t beg = warp ? a.seg.n0 : 0, n = warp ? a.seg.n1 : a.seg.n0;
    float ls = 0.f, ds = 0.f;
    int hits = 0;
    for (int64_t i = lane; i < n; i += 32) {
      ls += __ldcg(a.row_loss + beg + i);
      ds += __ldcg(a.row_dscale + beg + i);
      hits += __ldcg(a.row_correct + beg + i);
    }
    ls = warp_sum(ls);
    ds = warp_sum(ds);
    hits = warp_sum_i(hits);
    if (lane == 0) {
      a.stats[warp].loss_mean = n > 0 ? ls / static_cast<float>(n) : 0.f;
      a.stats[warp].dscale = ds;
      a.stats[warp].correct = hits;
      a.stats[warp].n = static_cast<int32_t>(n);
    }
  }
}

// deterministic per-run reduction of the per-row results (fixed summation order)
__global__ void __launch_bounds__(1024)
    seg_stats_kernel(const float* __restrict__ row_loss, const int32_t* __restrict__ row_correct,
                     const float* __restrict__ row_dscale, int64_t n0, int64_t n1, uml_seg_stats* __restrict__ out) {
  __shared__ float sh[32];
  const int s = blockIdx.x;
  const int64_t beg = s ? n0 : 0, n = s ? n1 : n0;
  float ls = 0.f, ds = 0.f;
  int hits = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    ls += row_loss[beg + i];
    if (row_dscale) ds += row_dscale[beg + i];
    hits += row_correct[beg + i];
  }
  const float lsum = block_reduce_sum(ls, sh);
  const float dsum = block_reduce_sum(ds, sh);
  const float hsum = block_reduce_sum(static_cast<float>(hits), sh);
  if (threadIdx.x == 0) {
    out[s].loss_mean = n > 0 ? lsum / static_cast<float>(n) : 0.f;
    out[s].dscale = dsum;
    out[s].correct = static_cast<int32_t>(hsum + 0.5f);
    out[s].n = static_cast<int32_t>(n);
  }
}

// -------------------------------------------------------------------------------------------------
// K7 streaming eval: 32 rows per CTA, classes swept in tiles of 64 with an online softmax/argmax,
// so logits never leave the SM (the reference ships them to the CPU, finetune.py:301-304).
// -------------------------------------------------------------------------------------------------
struct RowStat {
  float m, l, lab;
  int arg;
};
__device__ __forceinline__ void merge(RowStat& a, const RowStat& b) {
  const float M = fmaxf(a.m, b.m);
  const float la = (a.m == -INFINITY) ? 0.f : a.l * expf(a.m - M);
  const float lb = (b.m == -INFINITY) ? 0.f : b.l * expf(b.m - M);
  if (b.m > a.m || (b.m == a.m && b.arg < a.arg)) a.arg = b.arg;
  a.m = M;
  a.l = la + lb;
  a.lab += b.lab;
}

// Heads of a sweep group evaluated over the same bank in one launch (blockIdx.y): head h reads its weights at
// W + id[h] * w_stride and writes rows [h * n_rows, (h + 1) * n_rows) of the outputs.  A single head is {n = 1, id = {0}}.
struct EvalHeads {
  int n;
  int id[UML_SWEEP_MAX_HEADS];
  float scale[UML_SWEEP_MAX_HEADS];
};

__global__ void __launch_bounds__(256)
    eval_kernel(const float* __restrict__ X, int64_t ldx, const int64_t* __restrict__ labels, int64_t n_rows, int D,
                const float* __restrict__ W0, int64_t w_stride, int C, EvalHeads heads, float* __restrict__ row_loss0,
                int32_t* __restrict__ row_pred0) {
  constexpr int BM = 32, BN = 64, TM = 2, TN = 4, kBK = 16;
  const float* __restrict__ W = W0 + heads.id[blockIdx.y] * w_stride;
  const float scale = heads.scale[blockIdx.y];
  float* __restrict__ row_loss = row_loss0 + blockIdx.y * n_rows;
  int32_t* __restrict__ row_pred = row_pred0 + blockIdx.y * n_rows;
  __shared__ float Xs[kBK][BM + 4];
  __shared__ float Ws[kBK][BN + 4];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * BM;
  RowStat st[TM];
  int lab[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    st[i].m = -INFINITY; st[i].l = 0.f; st[i].lab = 0.f; st[i].arg = INT_MAX;
    const int64_t m = m0 + ty * TM + i;
    lab[i] = m < n_rows ? static_cast<int>(labels[m]) : -1;
  }
  for (int c0 = 0; c0 < C; c0 += BN) {
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += kBK) {
      for (int e = t; e < BM * kBK; e += 256) {
        const int m = e / kBK, k = e % kBK;
        float x = 0.f;
        if (m0 + m < n_rows && k0 + k < D) x = X[(m0 + m) * ldx + k0 + k];
        Xs[k][m] = x;
      }
      for (int e = t; e < BN * kBK; e += 256) {
        const int n = e / kBK, k = e % kBK;
        float x = 0.f;
        if (c0 + n < C && k0 + k < D) x = W[static_cast<int64_t>(c0 + n) * D + k0 + k];
        Ws[k][n] = x;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kBK; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = Xs[k][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Ws[k][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int c = c0 + tx * TN + j;
        if (c >= C) continue;
        const float x = acc[i][j] * scale;
        if (c == lab[i]) st[i].lab = x;
        if (x > st[i].m) {
          st[i].l = st[i].l * expf(st[i].m - x) + 1.f;  // exp(-inf) = 0 on the first hit
          st[i].m = x;
          st[i].arg = c;
        } else {
          st[i].l += expf(x - st[i].m);
        }
      }
    }
  }
  // merge the 16 threads (tx) that share a row: they sit in one half-warp
#pragma unroll
  for (int i = 0; i < TM; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      RowStat b;
      b.m = __shfl_xor_sync(0xffffffffu, st[i].m, o);
      b.l = __shfl_xor_sync(0xffffffffu, st[i].l, o);
      b.lab = __shfl_xor_sync(0xffffffffu, st[i].lab, o);
      b.arg = __shfl_xor_sync(0xffffffffu, st[i].arg, o);
      merge(st[i], b);
    }
    const int64_t m = m0 + ty * TM + i;
    if (tx == 0 && m < n_rows) {
      row_loss[m] = logf(st[i].l) - (st[i].lab - st[i].m);
      row_pred[m] = st[i].arg;
    }
  }
}

// mean over reference batches of the batch-mean loss, and total hits; one CTA per head, fixed order: a warp per
// reference batch (lanes stride over its rows, shuffle tree), the batch means added up in batch order by the warps'
// partial sums in warp order - the result does not depend on timing
__global__ void __launch_bounds__(1024)
    eval_reduce_kernel(const float* __restrict__ row_loss, const int32_t* __restrict__ row_pred,
                       const int64_t* __restrict__ labels, int64_t n_rows, int64_t bs, float* __restrict__ out_loss,
                       int32_t* __restrict__ out_correct) {
  __shared__ float sh_loss[32];
  __shared__ int sh_hits[32];
  row_loss += blockIdx.x * n_rows;  // (one CTA per head of a group)
  row_pred += blockIdx.x * n_rows;
  out_loss += blockIdx.x;
  out_correct += blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const int64_t n_batches = (n_rows + bs - 1) / bs;
  float acc = 0.f;
  int hits = 0;
  for (int64_t b = warp; b < n_batches; b += n_warps) {
    const int64_t beg = b * bs, end = min(n_rows, beg + bs);
    float s = 0.f;
    for (int64_t i = beg + lane; i < end; i += 32) {
      s += row_loss[i];
      hits += labels ? (row_pred[i] == static_cast<int32_t>(labels[i])) : (row_pred[i] != 0);
    }
    s = uml::warp_sum(s);
    acc += s / static_cast<float>(end - beg);
  }
  hits = uml::warp_sum_i(hits);
  if (lane == 0) {
    sh_loss[warp] = acc;
    sh_hits[warp] = hits;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    int h = 0;
    for (int w = 0; w < n_warps; ++w) {
      tot += sh_loss[w];
      h += sh_hits[w];
    }
    out_loss[0] = tot / static_cast<float>(n_batches);
    out_correct[0] = h;
  }
}

// K8 gradient diagnostics: dot, |a|^2, |b|^2, #(sign(a)==sign(b)); partials then a fixed-order finish
__global__ void __launch_bounds__(256)
    grad_diag_partial(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ part) {
  __shared__ float sh[8];
  float d = 0.f, aa = 0.f, bb = 0.f, ag = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x = a[i], y = b[i];
    d = fmaf(x, y, d);
    aa = fmaf(x, x, aa);
    bb = fmaf(y, y, bb);
    const int sx = (x > 0.f) - (x < 0.f), sy = (y > 0.f) - (y < 0.f);
    ag += (sx == sy) ? 1.f : 0.f;
  }
  const float r0 = block_reduce_sum(d, sh), r1 = block_reduce_sum(aa, sh), r2 = block_reduce_sum(bb, sh),
              r3 = block_reduce_sum(ag, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x * 4 + 0] = r0;
    part[blockIdx.x * 4 + 1] = r1;
    part[blockIdx.x * 4 + 2] = r2;
    part[blockIdx.x * 4 + 3] = r3;
  }
}
__global__ void grad_diag_finish(const float* __restrict__ part, int nblk, float* __restrict__ out4) {
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int i = 0; i < nblk; ++i) s += part[i * 4 + threadIdx.x];
    out4[threadIdx.x] = s;
  }
}

static int check_segs(const uml_segment* segs, int32_t nseg) {
  UML_REQUIRE(segs && nseg >= 1 && nseg <= UML_MAX_SEGMENTS, "need 1..%d segments", UML_MAX_SEGMENTS);
  for (int i = 0; i < nseg; ++i) {
    UML_REQUIRE(segs[i].n >= 0 && (segs[i].n == 0 || (segs[i].rows && segs[i].labels)), "segment %d: null rows/labels", i);
  }
  return 0;
}

static RowSrc seg_src(const uml_segment* segs, int32_t nseg) {
  RowSrc r;
  for (int i = 0; i < 2; ++i) {
    const uml_segment& s = segs[i < nseg ? i : 0];
    r.base[i] = static_cast<const float*>(s.rows);
    r.idx[i] = s.idx;
    r.ld[i] = s.ld;
  }
  r.n0 = nseg > 1 ? segs[0].n : INT64_MAX;
  return r;
}

}  // namespace uml

static long long g_fused_steps = 0;  // fused-step launches of this process (launch accounting of the callers)

extern "C" {

int uml_head_fwd_ce_f32(const uml_segment* segs, int32_t nseg, int32_t dim, const float* W, int32_t n_classes,
                        float* G, int64_t ldg, float* row_loss, int32_t* row_correct, float* row_dscale,
                        uml_seg_stats* stats, int64_t g_capacity_rows, void* stream) {
  using namespace uml;
  if (check_segs(segs, nseg)) return 1;
  UML_REQUIRE(W && G && row_loss && row_correct && row_dscale && stats && dim > 0 && n_classes > 0 && ldg >= n_classes,
              "head_fwd_ce_f32: bad arguments");
  cudaStream_t st = as_stream(stream);
  const int64_t n0 = segs[0].n, n1 = nseg > 1 ? segs[1].n : 0, total = n0 + n1;
  if (total > 0) {
    DeviceUpdate none;
    memset(&none, 0, sizeof(none));
    // small batches: split the contraction over up to 8 planes of G (when the caller's G has room for them) so that
    // the logit GEMM fills the machine; the softmax kernel adds the planes back in a fixed order
    const int64_t ctas = ((total + 31) / 32) * ((n_classes + 31) / 32);
    int splits = 1;
    if (ctas < 2 * sm_count() && g_capacity_rows >= 2 * total) {
      const int64_t want = (2 * sm_count() + ctas - 1) / ctas, room = g_capacity_rows / total, deep = (dim + 63) / 64;
      splits = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, room), std::min<int64_t>(deep, 8))));
    }
    const int64_t plane = total * ldg;
    if (launch_sgemm<true, true>(seg_src(segs, nseg), dense_src(W, dim), G, ldg, total, n_classes, dim, 1.f, nullptr,
                                 none, st, splits, plane))
      return 1;
    SegInfo si;
    si.n0 = n0;
    si.n1 = n1;
    for (int i = 0; i < 2; ++i) {
      const uml_segment& s = segs[i < nseg ? i : 0];
      si.idx[i] = s.label_idx ? s.label_idx : s.idx;
      si.labels[i] = s.labels;
      si.scale_dev[i] = s.scale_dev;
      si.scale[i] = s.scale;
      si.weight[i] = s.loss_weight;
    }
    softmax_ce_grad_kernel<<<static_cast<unsigned>(total), 256, 0, st>>>(G, ldg, n_classes, si, row_loss, row_correct,
                                                                        row_dscale, splits, plane);
    UML_CUDA(cudaGetLastError());
  }
  seg_stats_kernel<<<nseg, 1024, 0, st>>>(row_loss, row_correct, row_dscale, n0, n1, stats);
  UML_CUDA(cudaGetLastError());
  return 0;
}

int uml_head_step_fused_f32(const uml_segment* segs, int32_t nseg, int32_t dim, float* W, int32_t n_classes, float* G,
                            int64_t ldg, float* row_loss, int32_t* row_correct, float* row_dscale, uml_seg_stats* stats,
                            const uml_update* upd, int32_t* launched, void* stream) {
  using namespace uml;
  UML_REQUIRE(launched != nullptr, "head_step_fused_f32: null launched");
  *launched = 0;
  if (check_segs(segs, nseg)) return 1;
  UML_REQUIRE(W && G && row_loss && row_correct && row_dscale && stats && dim > 0 && n_classes > 0 && ldg >= n_classes,
              "head_step_fused_f32: bad arguments");
  static const bool enabled = [] {
    const char* e = getenv("UML_FUSED_STEP");
    return !(e != nullptr && e[0] == '0');
  }();
  if (!enabled || !upd || upd->kind < 1 || upd->kind > 3 || !upd->m || (upd->kind != 3 && !upd->v)) return 0;
  const int64_t n0 = segs[0].n, n1 = nseg > 1 ? segs[1].n : 0, total = n0 + n1;
  if (total <= 0 || total > kFusedMaxRows || dim % 4 != 0) return 0;
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!aligned(W) || !aligned(upd->m) || (upd->kind != 3 && !aligned(upd->v))) return 0;
  for (int i = 0; i < nseg; ++i)
    if (segs[i].n > 0 && (!aligned(segs[i].rows) || segs[i].ld % 4 != 0)) return 0;
  const int grid = sm_count();
  const int64_t nc_max = (n_classes + grid - 1) / grid + 1;
  const size_t smem = static_cast<size_t>(total) * (dim + 16) * 4 + static_cast<size_t>(nc_max) * dim * 4 +
                      static_cast<size_t>(total) * (nc_max + 3) * 4;
  if (smem > 224 * 1024) return 0;
  static const cudaError_t attr =
      cudaFuncSetAttribute(head_step_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  UML_CUDA(attr);
  FusedStepArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.X = seg_src(segs, nseg);
  fa.seg.n0 = n0;
  fa.seg.n1 = n1;
  for (int i = 0; i < 2; ++i) {
    const uml_segment& sg = segs[i < nseg ? i : 0];
    fa.seg.idx[i] = sg.label_idx ? sg.label_idx : sg.idx;
    fa.seg.labels[i] = sg.labels;
    fa.seg.scale_dev[i] = sg.scale_dev;
    fa.seg.scale[i] = sg.scale;
    fa.seg.weight[i] = sg.loss_weight;
  }
  fa.W = W;
  fa.upd = make_update(upd);
  fa.G = G;
  fa.ldg = ldg;
  fa.C = n_classes;
  fa.D = dim;
  fa.R = static_cast<int>(total);
  fa.nseg = nseg;
  fa.row_loss = row_loss;
  fa.row_correct = row_correct;
  fa.row_dscale = row_dscale;
  fa.stats = stats;
  void* kargs[] = {&fa};
  UML_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(head_step_fused_kernel), dim3(grid), dim3(256), kargs, smem,
                                       as_stream(stream)));
  *launched = 1;
  ++g_fused_steps;
  return 0;
}

int uml_head_step_fused_count(void) { return static_cast<int>(g_fused_steps & 0x7fffffff); }

int uml_head_bwd_dw_f32(const uml_segment* segs, int32_t nseg, int32_t dim, const float* G, int64_t ldg,
                        int32_t n_classes, float* W, float* dW, const uml_update* upd, void* stream) {
  using namespace uml;
  if (check_segs(segs, nseg)) return 1;
  UML_REQUIRE(G && dim > 0 && n_classes > 0, "head_bwd_dw_f32: bad arguments");
  const bool fused = upd && upd->kind != 0;
  UML_REQUIRE(fused || dW, "head_bwd_dw_f32: need dW when no update is fused");
  UML_REQUIRE(!fused || (W && upd->m && (upd->kind == 3 || upd->v)), "head_bwd_dw_f32: fused update needs W, m, v");
  const int64_t total = segs[0].n + (nseg > 1 ? segs[1].n : 0);
  // dW[c,d] = sum_r G[r,c] X[r,d]   (A = G rows over k, M-contiguous; B = X rows over k, gathered)
  return launch_sgemm<false, false>(dense_src(G, ldg), seg_src(segs, nseg), dW, dim, n_classes, dim, total, 1.f, W,
                                    make_update(upd), as_stream(stream));
}

int uml_gemm_nt_f32(const float* A, int64_t lda, const int64_t* a_row_idx, const float* B, int64_t ldb, float* C,
                    int64_t ldc, int64_t m, int64_t n, int64_t k, float alpha, void* stream) {
  using namespace uml;
  UML_REQUIRE(A && B && C, "gemm_nt: null pointer");
  DeviceUpdate none;
  memset(&none, 0, sizeof(none));
  return launch_sgemm<true, true>(dense_src(A, lda, a_row_idx), dense_src(B, ldb), C, ldc, m, n, k, alpha, nullptr, none,
                                  as_stream(stream));
}

int uml_gemm_nn_f32(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t m,
                    int64_t n, int64_t k, float alpha, void* stream) {
  using namespace uml;
  UML_REQUIRE(A && B && C, "gemm_nn: null pointer");
  DeviceUpdate none;
  memset(&none, 0, sizeof(none));
  return launch_sgemm<true, false>(dense_src(A, lda), dense_src(B, ldb), C, ldc, m, n, k, alpha, nullptr, none,
                                   as_stream(stream));
}

int uml_gemm_tn_f32(const float* A, int64_t lda, const float* B, int64_t ldb, const int64_t* b_row_idx, float* C,
                    int64_t ldc, int64_t m, int64_t n, int64_t k, float alpha, float* P, const uml_update* upd,
                    void* stream) {
  using namespace uml;
  const bool fused = upd && upd->kind != 0;
  UML_REQUIRE(A && B && (C || fused), "gemm_tn: null pointer");
  UML_REQUIRE(!fused || (P && upd->m && (upd->kind == 3 || upd->v)), "gemm_tn: fused update needs P, m, v");
  return launch_sgemm<false, false>(dense_src(A, lda), dense_src(B, ldb, b_row_idx), C, ldc, m, n, k, alpha, P,
                                    make_update(upd), as_stream(stream));
}

int uml_eval_f32(const float* feats, int64_t ld, const int64_t* labels, int64_t n_rows, int32_t dim, const float* W,
                 int32_t n_classes, float scale, float* row_loss, int32_t* row_pred, void* stream) {
  using namespace uml;
  UML_REQUIRE(feats && labels && W && row_loss && row_pred && dim > 0 && n_classes > 0 && n_rows >= 0,
              "eval_f32: bad arguments");
  if (n_rows == 0) return 0;
  EvalHeads h;
  memset(&h, 0, sizeof(h));
  h.n = 1;
  h.scale[0] = scale;
  eval_kernel<<<static_cast<unsigned>((n_rows + 31) / 32), 256, 0, as_stream(stream)>>>(
      feats, ld, labels, n_rows, dim, W, 0, n_classes, h, row_loss, row_pred);
  UML_CUDA(cudaGetLastError());
  return 0;
}

int uml_eval_group_f32(const float* feats, int64_t ld, const int64_t* labels, int64_t n_rows, int32_t dim, const float* W,
                       int64_t w_stride, const int32_t* head_ids, const float* scales, int32_t n_heads, int32_t n_classes,
                       float* row_loss, int32_t* row_pred, void* stream) {
  using namespace uml;
  UML_REQUIRE(feats && labels && W && head_ids && scales && row_loss && row_pred && dim > 0 && n_classes > 0 && n_rows >= 0 &&
                  n_heads >= 1 && n_heads <= UML_SWEEP_MAX_HEADS && w_stride >= static_cast<int64_t>(dim) * n_classes,
              "eval_group_f32: bad arguments");
  if (n_rows == 0) return 0;
  EvalHeads h;
  memset(&h, 0, sizeof(h));
  h.n = n_heads;
  for (int i = 0; i < n_heads; ++i) {
    UML_REQUIRE(head_ids[i] >= 0, "eval_group_f32: negative head id");
    h.id[i] = head_ids[i];
    h.scale[i] = scales[i];
  }
  eval_kernel<<<dim3(static_cast<unsigned>((n_rows + 31) / 32), static_cast<unsigned>(n_heads)), 256, 0, as_stream(stream)>>>(
      feats, ld, labels, n_rows, dim, W, w_stride, n_classes, h, row_loss, row_pred);
  UML_CUDA(cudaGetLastError());
  return 0;
}

int uml_eval_reduce_group(const float* row_loss, const int32_t* row_pred, const int64_t* labels, int64_t n_rows,
                          int64_t batch_size, int32_t n_heads, float* out_loss, int32_t* out_correct, void* stream) {
  using namespace uml;
  UML_REQUIRE(row_loss && row_pred && out_loss && out_correct && n_rows > 0 && batch_size > 0 && n_heads >= 1,
              "eval_reduce_group: bad arguments");
  eval_reduce_kernel<<<static_cast<unsigned>(n_heads), 1024, 0, as_stream(stream)>>>(row_loss, row_pred, labels, n_rows, batch_size,
                                                                                  out_loss, out_correct);
  UML_CUDA(cudaGetLastError());
  return 0;
}

int uml_eval_reduce(const float* row_loss, const int32_t* row_pred, const int64_t* labels, int64_t n_rows,
                    int64_t batch_size, float* out_loss, int32_t* out_correct, void* stream) {
  using namespace uml;
  UML_REQUIRE(row_loss && row_pred && out_loss && out_correct && n_rows > 0 && batch_size > 0,
              "eval_reduce: bad arguments");
  eval_reduce_kernel<<<1, 1024, 0, as_stream(stream)>>>(row_loss, row_pred, labels, n_rows, batch_size, out_loss,
                                                       out_correct);
  UML_CUDA(cudaGetLastError());
  return 0;
}

int uml_reduce_seg_stats(const float* row_loss, const int32_t* row_correct, const float* row_dscale,
                         const int64_t* seg_rows, int32_t nseg, uml_seg_stats* stats, void* stream) {
  using namespace uml;
  UML_REQUIRE(row_loss && row_correct && seg_rows && stats && nseg >= 1 && nseg <= UML_MAX_SEGMENTS,
              "reduce_seg_stats: bad arguments");
  seg_stats_kernel<<<nseg, 1024, 0, as_stream(stream)>>>(row_loss, row_correct, row_dscale, seg_rows[0],
                                                        nseg > 1 ? seg_rows[1] : 0, stats);
  UML_CUDA(cudaGetLastError());
  return 0;
}

int uml_grad_diag(const float* a, const float* b, int64_t n, float* workspace, float* out4, void* stream) {
  using namespace uml;
  UML_REQUIRE(a && b && workspace && out4 && n > 0, "grad_diag: bad arguments");
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, UML_DIAG_BLOCKS));
  grad_diag_partial<<<blocks, 256, 0, as_stream(stream)>>>(a, b, n, workspace);
  UML_CUDA(cudaGetLastError());
  grad_diag_finish<<<1, 32, 0, as_stream(stream)>>>(workspace, blocks, out4);
  UML_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
