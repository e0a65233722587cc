// Adam / AdamW arithmetic shared by the optimizer kernels (optim.cu) and the fused data-parallel update (dp.cu).
#pragma once
#include <cmath>

#include "common.cuh"

namespace uml {

struct AdamArgs {
  float decay, beta1, beta2, one_m_b1, one_m_b2, eps, step_size, bc2_sqrt_inv, wd;
  int decoupled;
};

inline AdamArgs make_adam(double lr, double b1, double b2, double eps, double wd, int64_t step, int decoupled) {
  AdamArgs a;
  const double t = static_cast<double>(step);
  a.decay = static_cast<float>(1.0 - lr * wd);
  a.beta1 = static_cast<float>(b1);
  a.beta2 = static_cast<float>(b2);
  a.one_m_b1 = static_cast<float>(1.0 - b1);
  a.one_m_b2 = static_cast<float>(1.0 - b2);
  a.eps = static_cast<float>(eps);
  a.step_size = static_cast<float>(lr / (1.0 - std::pow(b1, t)));
  a.bc2_sqrt_inv = static_cast<float>(1.0 / std::sqrt(1.0 - std::pow(b2, t)));
  a.wd = static_cast<float>(wd);
  a.decoupled = decoupled;
  return a;
}

__device__ __forceinline__ float adam_one(const AdamArgs& a, float w, float g, float& m, float& v) {
  if (a.decoupled) w *= a.decay;
  else if (a.wd != 0.f) g = fmaf(a.wd, w, g);
  m = m + (g - m) * a.one_m_b1;
  v = v * a.beta2 + a.one_m_b2 * g * g;
  const float denom = sqrtf(v) * a.bc2_sqrt_inv + a.eps;
  return w - a.step_size * (m / denom);
}

__device__ __forceinline__ uint2 pack_bf16x4(float4 x) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&lo);
  o.y = *reinterpret_cast<uint32_t*>(&hi);
  return o;
}

}  // namespace uml
