// a-14 - the linear analogue of the UML step: Gaussian_experiment's SharedAutoencoder
// (reference Gaussian_experiment/model.py:5-49, loop body main.py:47-59).
//
//   recon_m = out_head_m( dec( enc( in_head_m(v_m) ) ) ),   enc/dec = Linear-ReLU-Linear SHARED by both
//   modalities, per-modality in/out heads (all with bias); loss = alpha_x MSE(x) + alpha_y MSE(y) ("xy") or
//   MSE(x) alone ("x": the y heads get no gradient and Adam never touches them); Adam(lr), torch defaults.
//
// The model has ~29 k parameters and the reference batch is 512 rows per modality: the step is pure latency.
// Two launches per step, nothing else:
//   gauss_fwd_bwd_kernel  grid (tiles, 2 modalities): a CTA stages its branch's weights in shared memory
//                         (rows padded to an odd stride, conflict-free in both GEMM directions), gathers its
//                         16 rows by index (UnpairedDataset's `idx % len`), runs the six layers forward, the MSE,
//                         the six layers backward, and writes its weight-gradient contribution to its own slot
//                         of a partial buffer (no atomics: the reduction order is fixed);
//   gauss_update_kernel   sums the slots in a fixed order (shared layers: both modalities), applies Adam, and
//                         reduces the per-tile losses into the step's {loss_x, loss_y} record.
#include <cmath>

#include "common.cuh"

namespace uml {

constexpr int kGaussRows = 16;      // batch rows per CTA
constexpr int kGaussThreads = 256;

struct GaussLayout {
  int obs, com, lat;
  // offsets (floats) into the flat parameter buffer, reference construction order, weight then bias
  int in_w[2], in_b[2], e0_w, e0_b, e2_w, e2_b, d0_w, d0_b, d2_w, d2_b, out_w[2], out_b[2], total;
};

static GaussLayout make_layout(int obs, int com, int lat) {
  GaussLayout L;
  L.obs = obs; L.com = com; L.lat = lat;
  int o = 0;
  auto take = [&](int n) { int r = o; o += n; return r; };
  for (int m = 0; m < 2; ++m) { L.in_w[m] = take(com * obs); L.in_b[m] = take(com); }
  L.e0_w = take(lat * com); L.e0_b = take(lat);
  L.e2_w = take(lat * lat); L.e2_b = take(lat);
  L.d0_w = take(lat * lat); L.d0_b = take(lat);
  L.d2_w = take(com * lat); L.d2_b = take(com);
  for (int m = 0; m < 2; ++m) { L.out_w[m] = take(obs * com); L.out_b[m] = take(obs); }
  L.total = o;
  return L;
}

// ---- shared-memory helpers (all called by the whole CTA) -----------------------------------------------
// W [out, in] global row-major -> shared with leading dimension in + 1
__device__ void stage_weight(const float* __restrict__ g, float* s, int out, int in) {
  for (int i = threadIdx.x; i < out * in; i += blockDim.x) s[(i / in) * (in + 1) + (i % in)] = g[i];
}
// y[r][j] = b[j] + sum_k x[r][k] W[j][k]
__device__ void dense_fwd(const float* x, int ldx, const float* W, const float* b, int in, int out, float* y, int ldy, int R) {
  for (int o = threadIdx.x; o < R * out; o += blockDim.x) {
    const int r = o / out, j = o - r * out;
    const float* w = W + j * (in + 1);
    const float* xr = x + r * ldx;
    float a0 = b[j], a1 = 0.f;
    int k = 0;
    for (; k + 1 < in; k += 2) {
      a0 = fmaf(xr[k], w[k], a0);
      a1 = fmaf(xr[k + 1], w[k + 1], a1);
    }
    if (k < in) a0 = fmaf(xr[k], w[k], a0);
    y[r * ldy + j] = a0 + a1;
  }
}
// dx[r][k] = sum_j dy[r][j] W[j][k]   (optionally masked by pre[r][k] > 0: the ReLU in front of this layer's input)
__device__ void dense_bwd_data(const float* dy, int ldy, const float* W, int in, int out, float* dx, int ldx, int R,
                               const float* pre, int ldp) {
  for (int o = threadIdx.x; o < R * in; o += blockDim.x) {
    const int r = o / in, k = o - r * in;
    const float* d = dy + r * ldy;
    float a = 0.f;
    for (int j = 0; j < out; ++j) a = fmaf(d[j], W[j * (in + 1) + k], a);
    if (pre && !(pre[r * ldp + k] > 0.f)) a = 0.f;
    dx[r * ldx + k] = a;
  }
}
// gW[j][k] = sum_r dy[r][j] x[r][k],  gb[j] = sum_r dy[r][j]   -> this CTA's slot of the partial buffer
__device__ void dense_bwd_weight(const float* dy, int ldy, const float* x, int ldx, int in, int out, int R,
                                 float* __restrict__ gW, float* __restrict__ gb) {
  for (int o = threadIdx.x; o < out * in; o += blockDim.x) {
    const int j = o / in, k = o - j * in;
    float a = 0.f;
    for (int r = 0; r < R; ++r) a = fmaf(dy[r * ldy + j], x[r * ldx + k], a);
    gW[o] = a;
  }
  for (int j = threadIdx.x; j < out; j += blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < R; ++r) a += dy[r * ldy + j];
    gb[j] = a;
  }
}
__device__ void relu_copy(const float* h, float* r, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) r[i] = fmaxf(h[i], 0.f);
}

struct GaussData {
  const float* data[2];   // [n_m, obs] row-major
  int64_t n[2];
  const int64_t* idx;     // [B] sampler indices (nullptr: dense rows 0..B-1, evaluation)
  int64_t B;
  float dscale[2];        // alpha_m * 2 / (B * obs); 0 => no gradient for that modality (forward + loss only)
};

__global__ void __launch_bounds__(kGaussThreads)
    gauss_fwd_bwd_kernel(const float* __restrict__ P, GaussLayout L, GaussData D, float* __restrict__ partial,
                         float* __restrict__ loss_part) {
  extern __shared__ float sm[];
  const int m = blockIdx.y, tile = blockIdx.x, n_tiles = gridDim.x;
  const int obs = L.obs, com = L.com, lat = L.lat;
  constexpr int R = kGaussRows;
  // ---- carve shared memory
  float* p = sm;
  auto carve = [&](int n) { float* r = p; p += n; return r; };
  float* Win = carve(com * (obs + 1));  float* bin = carve(com);
  float* We0 = carve(lat * (com + 1));  float* be0 = carve(lat);
  float* We2 = carve(lat * (lat + 1));  float* be2 = carve(lat);
  float* Wd0 = carve(lat * (lat + 1));  float* bd0 = carve(lat);
  float* Wd2 = carve(com * (lat + 1));  float* bd2 = carve(com);
  float* Wout = carve(obs * (com + 1)); float* bout = carve(obs);
  float* v = carve(R * obs);    // input rows
  float* a0 = carve(R * com);   // in_head output
  float* h1 = carve(R * lat);   float* r1 = carve(R * lat);
  float* z = carve(R * lat);    // latent
  float* h2 = carve(R * lat);   float* r2 = carve(R * lat);
  float* a3 = carve(R * com);   // decoder output
  float* dr = carve(R * obs);   // recon, then d recon
  float* dA = carve(R * com);   // gradient scratch (wide)
  float* dB = carve(R * lat);   // gradient scratch (narrow)
  float* dC = carve(R * lat);
  __shared__ float red[kGaussThreads / 32];

  stage_weight(P + L.in_w[m], Win, com, obs);
  stage_weight(P + L.e0_w, We0, lat, com);
  stage_weight(P + L.e2_w, We2, lat, lat);
  stage_weight(P + L.d0_w, Wd0, lat, lat);
  stage_weight(P + L.d2_w, Wd2, com, lat);
  stage_weight(P + L.out_w[m], Wout, obs, com);
  for (int i = threadIdx.x; i < com; i += blockDim.x) { bin[i] = P[L.in_b[m] + i]; bd2[i] = P[L.d2_b + i]; }
  for (int i = threadIdx.x; i < lat; i += blockDim.x) { be0[i] = P[L.e0_b + i]; be2[i] = P[L.e2_b + i]; bd0[i] = P[L.d0_b + i]; }
  for (int i = threadIdx.x; i < obs; i += blockDim.x) bout[i] = P[L.out_b[m] + i];
  const int64_t row0 = static_cast<int64_t>(tile) * R;
  for (int i = threadIdx.x; i < R * obs; i += blockDim.x) {
    const int r = i / obs, c = i - r * obs;
    float x = 0.f;
    if (row0 + r < D.B) {
      const int64_t src = (D.idx ? D.idx[row0 + r] : row0 + r) % D.n[m];  // UnpairedDataset.__getitem__: idx % len
      x = D.data[m][src * obs + c];
    }
    v[i] = x;
  }
  __syncthreads();

  // ---- forward
  dense_fwd(v, obs, Win, bin, obs, com, a0, com, R);      __syncthreads();
  dense_fwd(a0, com, We0, be0, com, lat, h1, lat, R);     __syncthreads();
  relu_copy(h1, r1, R * lat);                             __syncthreads();
  dense_fwd(r1, lat, We2, be2, lat, lat, z, lat, R);      __syncthreads();
  dense_fwd(z, lat, Wd0, bd0, lat, lat, h2, lat, R);      __syncthreads();
  relu_copy(h2, r2, R * lat);                             __syncthreads();
  dense_fwd(r2, lat, Wd2, bd2, lat, com, a3, com, R);     __syncthreads();
  dense_fwd(a3, com, Wout, bout, com, obs, dr, obs, R);   __syncthreads();

  // ---- loss (sum of squares of this tile) and d recon
  float ss = 0.f;
  const float ds = D.dscale[m];
  for (int i = threadIdx.x; i < R * obs; i += blockDim.x) {
    const bool valid = row0 + i / obs < D.B;
    const float diff = valid ? dr[i] - v[i] : 0.f;
    ss = fmaf(diff, diff, ss);
    dr[i] = diff * ds;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kGaussThreads / 32; ++w) t += red[w];
    loss_part[m * n_tiles + tile] = t;
  }
  if (ds == 0.f) return;  // forward-only modality ("x" mode's y branch, evaluation)

  // ---- backward: every layer writes its gradient slot, then propagates
  float* g = partial + (static_cast<int64_t>(m) * n_tiles + tile) * L.total;
  dense_bwd_weight(dr, obs, a3, com, com, obs, R, g + L.out_w[m], g + L.out_b[m]);
  dense_bwd_data(dr, obs, Wout, com, obs, dA, com, R, nullptr, 0);            __syncthreads();   // d a3
  dense_bwd_weight(dA, com, r2, lat, lat, com, R, g + L.d2_w, g + L.d2_b);
  dense_bwd_data(dA, com, Wd2, lat, com, dB, lat, R, h2, lat);                __syncthreads();   // d h2
  dense_bwd_weight(dB, lat, z, lat, lat, lat, R, g + L.d0_w, g + L.d0_b);
  dense_bwd_data(dB, lat, Wd0, lat, lat, dC, lat, R, nullptr, 0);             __syncthreads();   // d latent
  dense_bwd_weight(dC, lat, r1, lat, lat, lat, R, g + L.e2_w, g + L.e2_b);
  dense_bwd_data(dC, lat, We2, lat, lat, dB, lat, R, h1, lat);                __syncthreads();   // d h1
  dense_bwd_weight(dB, lat, a0, com, com, lat, R, g + L.e0_w, g + L.e0_b);
  dense_bwd_data(dB, lat, We0, com, lat, dA, com, R, nullptr, 0);             __syncthreads();   // d a0
  dense_bwd_weight(dA, com, v, obs, obs, com, R, g + L.in_w[m], g + L.in_b[m]);
}

struct GaussAdam {
  float beta1, beta2, one_m_b1, one_m_b2, eps, step_size, bc2_sqrt_inv;
};

__global__ void __launch_bounds__(256)
    gauss_update_kernel(float* __restrict__ P, float* __restrict__ M, float* __restrict__ V, GaussLayout L,
                        const float* __restrict__ partial, int n_tiles, int grad_x, int grad_y, GaussAdam A,
                        const float* __restrict__ loss_part, float inv_count, float* __restrict__ loss_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x < 2 && loss_out) {  // {loss_x, loss_y} of this step, fixed summation order
    float t = 0.f;
    for (int k = 0; k < n_tiles; ++k) t += loss_part[threadIdx.x * n_tiles + k];
    loss_out[threadIdx.x] = t * inv_count;
  }
  if (i >= L.total || !P) return;
  // which modalities contribute to parameter i: heads belong to one, the shared encoder/decoder to both
  bool use[2] = {grad_x != 0, grad_y != 0};
  if (i < L.in_w[1]) use[1] = false;                                   // in_head_x
  else if (i < L.e0_w) use[0] = false;                                 // in_head_y
  else if (i >= L.out_w[0] && i < L.out_w[1]) use[1] = false;          // out_head_x
  else if (i >= L.out_w[1]) use[0] = false;                            // out_head_y
  if (!use[0] && !use[1]) return;                                      // no gradient: Adam leaves it untouched
  float g = 0.f;
  for (int m = 0; m < 2; ++m)
    if (use[m])
      for (int k = 0; k < n_tiles; ++k) g += partial[(static_cast<int64_t>(m) * n_tiles + k) * L.total + i];
  float mm = M[i], vv = V[i];
  mm = mm + (g - mm) * A.one_m_b1;
  vv = vv * A.beta2 + A.one_m_b2 * g * g;
  M[i] = mm;
  V[i] = vv;
  P[i] = P[i] - A.step_size * (mm / (sqrtf(vv) * A.bc2_sqrt_inv + A.eps));
}

static size_t gauss_smem_bytes(const GaussLayout& L) {
  const int obs = L.obs, com = L.com, lat = L.lat, R = kGaussRows;
  size_t f = static_cast<size_t>(com) * (obs + 1) + com + static_cast<size_t>(lat) * (com + 1) + lat +
             2 * (static_cast<size_t>(lat) * (lat + 1) + lat) + static_cast<size_t>(com) * (lat + 1) + com +
             static_cast<size_t>(obs) * (com + 1) + obs;
  f += static_cast<size_t>(R) * (2 * obs + 3 * com + 7 * lat);
  return f * sizeof(float);
}

static int gauss_launch_fwd_bwd(const float* P, const GaussLayout& L, const GaussData& D, float* partial, float* loss_part,
                                cudaStream_t st) {
  const size_t smem = gauss_smem_bytes(L);
  UML_REQUIRE(smem <= 220 * 1024, "gauss: model too wide for the shared-memory resident kernel (%zu B)", smem);
  static size_t attr = 0;
  if (smem > attr) {
    UML_CUDA(cudaFuncSetAttribute(gauss_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  const int tiles = static_cast<int>((D.B + kGaussRows - 1) / kGaussRows);
  gauss_fwd_bwd_kernel<<<dim3(tiles, 2), kGaussThreads, smem, st>>>(P, L, D, partial, loss_part);
  UML_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace uml

extern "C" {

int uml_gauss_param_count(int32_t dim_obs, int32_t dim_common, int32_t dim_latent) {
  return uml::make_layout(dim_obs, dim_common, dim_latent).total;
}

int uml_gauss_workspace_floats(int32_t dim_obs, int32_t dim_common, int32_t dim_latent, int64_t batch) {
  const int64_t tiles = (batch + uml::kGaussRows - 1) / uml::kGaussRows;
  return static_cast<int>(2 * tiles * uml::make_layout(dim_obs, dim_common, dim_latent).total + 2 * tiles);
}

int uml_gauss_step(float* params, float* adam_m, float* adam_v, int32_t dim_obs, int32_t dim_common, int32_t dim_latent,
                   const float* data_x, int64_t n_x, const float* data_y, int64_t n_y, const int64_t* idx, int64_t batch,
                   int32_t mode_xy, float alpha_x, float alpha_y, double lr, double beta1, double beta2, double eps,
                   int64_t step, float* workspace, float* loss_out, void* stream) {
  using namespace uml;
  UML_REQUIRE(params && adam_m && adam_v && data_x && data_y && idx && workspace && batch > 0 && n_x > 0 && n_y > 0 && step >= 1,
              "gauss_step: bad arguments");
  const GaussLayout L = make_layout(dim_obs, dim_common, dim_latent);
  const int tiles = static_cast<int>((batch + kGaussRows - 1) / kGaussRows);
  float* partial = workspace;
  float* loss_part = workspace + static_cast<int64_t>(2) * tiles * L.total;
  GaussData D;
  D.data[0] = data_x; D.data[1] = data_y;
  D.n[0] = n_x; D.n[1] = n_y;
  D.idx = idx;
  D.B = batch;
  const double cnt = static_cast<double>(batch) * dim_obs;
  // "x" mode: loss = loss_x (no alpha), the y branch only reports its loss (main.py:52-54)
  D.dscale[0] = static_cast<float>((mode_xy ? alpha_x : 1.0) * 2.0 / cnt);
  D.dscale[1] = mode_xy ? static_cast<float>(alpha_y * 2.0 / cnt) : 0.f;
  int rc = gauss_launch_fwd_bwd(params, L, D, partial, loss_part, as_stream(stream));
  if (rc) return rc;
  GaussAdam A;
  const double t = static_cast<double>(step);
  A.beta1 = static_cast<float>(beta1);
  A.beta2 = static_cast<float>(beta2);
  A.one_m_b1 = static_cast<float>(1.0 - beta1);
  A.one_m_b2 = static_cast<float>(1.0 - beta2);
  A.eps = static_cast<float>(eps);
  A.step_size = static_cast<float>(lr / (1.0 - std::pow(beta1, t)));
  A.bc2_sqrt_inv = static_cast<float>(1.0 / std::sqrt(1.0 - std::pow(beta2, t)));
  gauss_update_kernel<<<(L.total + 255) / 256, 256, 0, as_stream(stream)>>>(
      params, adam_m, adam_v, L, partial, tiles, D.dscale[0] != 0.f, D.dscale[1] != 0.f, A, loss_part,
      static_cast<float>(1.0 / cnt), loss_out);
  UML_CUDA(cudaGetLastError());
  return 0;
}

// forward + MSE of both modalities over dense rows (validation, main.py:68-72): loss_out = {loss_x, loss_y}
int uml_gauss_eval(const float* params, int32_t dim_obs, int32_t dim_common, int32_t dim_latent, const float* data_x,
                   const float* data_y, int64_t n_rows, float* workspace, float* loss_out, void* stream) {
  using namespace uml;
  UML_REQUIRE(params && data_x && data_y && workspace && loss_out && n_rows > 0, "gauss_eval: bad arguments");
  const GaussLayout L = make_layout(dim_obs, dim_common, dim_latent);
  const int tiles = static_cast<int>((n_rows + kGaussRows - 1) / kGaussRows);
  GaussData D;
  D.data[0] = data_x; D.data[1] = data_y;
  D.n[0] = D.n[1] = n_rows;
  D.idx = nullptr;
  D.B = n_rows;
  D.dscale[0] = D.dscale[1] = 0.f;
  float* loss_part = workspace;  // 2 * tiles floats
  int rc = gauss_launch_fwd_bwd(params, L, D, nullptr, loss_part, as_stream(stream));
  if (rc) return rc;
  GaussAdam A;
  memset(&A, 0, sizeof(A));
  gauss_update_kernel<<<1, 256, 0, as_stream(stream)>>>(nullptr, nullptr, nullptr, L, nullptr, tiles, 0, 0, A, loss_part,
                                                         static_cast<float>(1.0 / (static_cast<double>(n_rows) * dim_obs)),
                                                         loss_out);
  UML_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
