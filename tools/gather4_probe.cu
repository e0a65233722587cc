// Probe of cp.async.bulk.tensor.2d.tile::gather4 on sm_100a (no public docs in this image):
//  (1) does a tensor map with box {64, 1} + SWIZZLE_128B work, and does a gather4 of rows r..r+3 written at
//      tile_base + r*128 produce the same shared-memory image as the regular {64, 128} box load?
//  (2) throughput of a persistent kernel that gathers random rows (128-row x 64-col bf16 tiles, 12 k-blocks).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/gather4_probe tools/gather4_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

// ---- (1) layout equality ------------------------------------------------------------------------------
__global__ void layout_kernel(const __grid_constant__ CUtensorMap box_map, const __grid_constant__ CUtensorMap g4_map,
                              const int* __restrict__ rows, int col0, uint4* out_box, uint4* out_g4) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* a = smem;            // 16 KB: regular box load of rows rows[0]..rows[0]+127 (rows are contiguous in this test)
  unsigned char* b = smem + 16384;    // 16 KB: 32 gather4 loads
  __shared__ __align__(8) uint64_t bar[2];
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect(&bar[0], 16384);
    tma_load_2d(a, &box_map, &bar[0], col0, rows[0]);
    mbar_expect(&bar[1], 16384);
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int l = threadIdx.x;
    tma_gather4(b + l * 512, &g4_map, &bar[1], col0, rows[4 * l], rows[4 * l + 1], rows[4 * l + 2], rows[4 * l + 3]);
  }
  mbar_wait(&bar[0], 0);
  mbar_wait(&bar[1], 0);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    out_box[i] = reinterpret_cast<uint4*>(a)[i];
    out_g4[i] = reinterpret_cast<uint4*>(b)[i];
  }
}

// ---- (2) throughput -----------------------------------------------------------------------------------
constexpr int kStages = 8;
template <bool kGather>
__global__ void __launch_bounds__(64) stream_kernel(const __grid_constant__ CUtensorMap box_map, const __grid_constant__ CUtensorMap g4_map,
                                                    const int* __restrict__ rows, int n_tiles, int num_kb, int passes,
                                                    unsigned long long* sink) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int* r = rows + tile * 128 + lane * 4;
      const int r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3];
      for (int p = 0; p < passes; ++p)
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          if (lane == 0) mbar_expect(&full[s], 16384);
          __syncwarp();
          if (kGather) tma_gather4(smem + s * 16384 + lane * 512, &g4_map, &full[s], kb * 64, r0, r1, r2, r3);
          else if (lane == 0) tma_load_2d(smem + s * 16384, &box_map, &full[s], kb * 64, tile * 128);
        }
    }
  } else {
    uint32_t it = 0;
    unsigned long long acc = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int p = 0; p < passes; ++p)
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(&full[s], ph);
          acc += reinterpret_cast<const unsigned long long*>(smem + s * 16384)[lane * 17];
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
    if (acc == 0x1234567ull) sink[0] = acc;
  }
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int64_t N = 1281167, D = 768;
  const int n_rows = 37888, n_tiles = n_rows / 128;
  __nv_bfloat16* bank;
  CK(cudaMalloc(&bank, N * D * 2));
  std::vector<uint16_t> h(1 << 20);
  for (size_t i = 0; i < h.size(); ++i) h[i] = static_cast<uint16_t>(i * 2654435761u >> 16);
  for (int64_t off = 0; off < N * D; off += h.size())
    CK(cudaMemcpy(reinterpret_cast<uint16_t*>(bank) + off, h.data(), std::min<int64_t>(h.size(), N * D - off) * 2, cudaMemcpyHostToDevice));
  // make rows distinguishable: first element of each row = row id low bits
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  encode_fn enc = reinterpret_cast<encode_fn>(fnp);
  CUtensorMap box_map, g4_map;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(N)}, strides[1] = {static_cast<cuuint64_t>(D * 2)};
  cuuint32_t estr[2] = {1, 1};
  cuuint32_t box_a[2] = {64, 128}, box_g[2] = {64, 1};
  CUresult r1 = enc(&box_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, bank, dims, strides, box_a, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&g4_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, bank, dims, strides, box_g, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box {64,128}: %d   encode gather4 box {64,1}: %d\n", (int)r1, (int)r2);
  if (r1 || r2) return 1;

  // (1) layout: contiguous rows 1000..1127 via both paths
  std::vector<int> rows_h(n_rows);
  for (int i = 0; i < 128; ++i) rows_h[i] = 1000 + i;
  int* rows_d;
  CK(cudaMalloc(&rows_d, n_rows * 4));
  CK(cudaMemcpy(rows_d, rows_h.data(), 128 * 4, cudaMemcpyHostToDevice));
  uint4 *ob, *og;
  CK(cudaMalloc(&ob, 16384));
  CK(cudaMalloc(&og, 16384));
  CK(cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024));
  layout_kernel<<<1, 128, 34 * 1024>>>(box_map, g4_map, rows_d, 128, ob, og);
  CK(cudaDeviceSynchronize());
  std::vector<uint8_t> hb(16384), hg(16384);
  CK(cudaMemcpy(hb.data(), ob, 16384, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hg.data(), og, 16384, cudaMemcpyDeviceToHost));
  int diff = 0;
  for (int i = 0; i < 16384; ++i) diff += hb[i] != hg[i];
  printf("layout: %d differing bytes between the box load and 32 gather4 loads (0 = identical swizzled image)\n", diff);

  // (2) throughput with random rows
  uint32_t st = 12345;
  for (int i = 0; i < n_rows; ++i) { st = st * 1664525u + 1013904223u; rows_h[i] = static_cast<int>((static_cast<uint64_t>(st) * N) >> 32); }
  CK(cudaMemcpy(rows_d, rows_h.data(), n_rows * 4, cudaMemcpyHostToDevice));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  const int smem = kStages * 16384 + 1024;
  CK(cudaFuncSetAttribute(stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  char* flush;
  CK(cudaMalloc(&flush, 256 << 20));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int passes = 1; passes <= 4; passes += 3) {
    for (int mode = 0; mode < 2; ++mode) {
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        CK(cudaMemsetAsync(flush, rep, 256 << 20));
        cudaEventRecord(e0);
        if (mode) stream_kernel<true><<<148, 64, smem>>>(box_map, g4_map, rows_d, n_tiles, 12, passes, sink);
        else stream_kernel<false><<<148, 64, smem>>>(box_map, g4_map, rows_d, n_tiles, 12, passes, sink);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double bytes = static_cast<double>(n_rows) * D * 2 * passes;
      printf("%s passes=%d: %.1f us, %.0f GB/s into shared memory (unique bytes %.1f MB)\n", mode ? "gather4 (random rows)" : "box load (contiguous)",
             passes, best * 1e3, bytes / best / 1e6, n_rows * D * 2 / 1e6);
    }
  }
  return 0;
}
