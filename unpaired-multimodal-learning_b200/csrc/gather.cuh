// Row-copy building blocks shared by gather.cu (stand-alone gather kernels) and tc_fwd.cu (the fix-up launch can
// carry the NEXT step's gather as extra CTAs).
#pragma once
#include "common.cuh"

namespace uml {

struct CopySeg {
  const unsigned char* bank;   // row-major, row_bytes per row
  const int64_t* idx;          // gather indices (required)
  const int64_t* labels;       // bank labels (int64) or nullptr
  int64_t n;
};

// what the fix-up launch needs to gather the next step's operand (both runs) as a side job
struct GatherJob {
  CopySeg s0, s1;
  int vec_per_row;             // 16-byte vectors per row
  uint4* out;
  int64_t out_pitch_vec;
  int32_t* out_labels;
  int blocks;                  // CTAs of the launch that work on this job (0 = none)
};

// Fix-up -> dW hand-over at split granularity: the fix-up launch counts its finished 8-row CTAs per K split of the dW
// GEMM (split boundaries are multiples of 64 rows), the dW kernel - running concurrently on another stream - lets
// each split start as soon as ITS rows of G are final instead of waiting for the whole fix-up pass.
struct FixupSignal {
  unsigned* done;     // [n_splits] counters, zeroed before the forward kernel
  int n_splits;
  int64_t num_kb;     // ceil(rows / 64)
};

#ifdef __CUDACC__
// rows warp_id, warp_id + n_warps, ... of the concatenated runs: one warp per row, 16-byte vectors, four in flight
__device__ __forceinline__ void gather_rows_by_warp(const CopySeg& s0, const CopySeg& s1, int vec_per_row, uint4* __restrict__ out,
                                                    int64_t out_pitch_vec, int32_t* __restrict__ out_labels, int64_t warp_id,
                                                    int64_t n_warps) {
  const int64_t n = s0.n + s1.n;
  const int lane = threadIdx.x & 31;
  for (int64_t r = warp_id; r < n; r += n_warps) {
    const bool second = r >= s0.n;
    const CopySeg& sg = second ? s1 : s0;
    const int64_t src = __ldg(sg.idx + (second ? r - s0.n : r));
    const uint4* row = reinterpret_cast<const uint4*>(sg.bank) + src * vec_per_row;
    uint4* dst = out + r * out_pitch_vec;
    if (lane == 0 && out_labels) out_labels[r] = static_cast<int32_t>(__ldg(sg.labels + src));
    for (int v0 = 0; v0 < vec_per_row; v0 += 128) {
      uint4 x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int v = v0 + j * 32 + lane;
        if (v < vec_per_row) x[j] = __ldcs(row + v);  // streamed: a bank row is read once per epoch
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int v = v0 + j * 32 + lane;
        if (v < vec_per_row) __stcs(dst + v, x[j]);  // read again a step or two later, after 0.5 GB of other traffic: not worth L2
      }
    }
  }
}
#endif

}  // namespace uml
