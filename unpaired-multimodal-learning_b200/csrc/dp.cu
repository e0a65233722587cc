// Data-parallel plumbing: one NCCL communicator per process (one process per GPU), used by the step
// launcher to sum the head gradient over the ranks without returning to Python.  libnccl is resolved at
// run time with dlopen (the same libnccl.so.2 torch.distributed already loaded), so the library has no
// link-time dependency on it and still loads on a box without NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>

#include "common.cuh"
#include "optim.cuh"

namespace {

struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int world = 1;
};
Nccl g;

int load_nccl() {
  if (g.handle) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g.handle) break;
  }
  UML_REQUIRE(g.handle, "dp: cannot dlopen libnccl.so.2 (%s)", dlerror());
  g.GetUniqueId = reinterpret_cast<decltype(g.GetUniqueId)>(dlsym(g.handle, "ncclGetUniqueId"));
  g.CommInitRank = reinterpret_cast<decltype(g.CommInitRank)>(dlsym(g.handle, "ncclCommInitRank"));
  g.AllReduce = reinterpret_cast<decltype(g.AllReduce)>(dlsym(g.handle, "ncclAllReduce"));
  g.CommDestroy = reinterpret_cast<decltype(g.CommDestroy)>(dlsym(g.handle, "ncclCommDestroy"));
  g.GetErrorString = reinterpret_cast<decltype(g.GetErrorString)>(dlsym(g.handle, "ncclGetErrorString"));
  UML_REQUIRE(g.GetUniqueId && g.CommInitRank && g.AllReduce && g.CommDestroy && g.GetErrorString,
              "dp: libnccl lacks a required symbol");
  return 0;
}

#define UML_NCCL(expr)                                                                  \
  do {                                                                                  \
    ncclResult_t _r = (expr);                                                           \
    if (_r != ncclSuccess) UML_FAIL("%s -> %s", #expr, g.GetErrorString(_r));           \
  } while (0)

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// Two-shot all-reduce over NVLink peer memory (one node, NVSwitch): every rank owns one cudaMalloc'd exchange
// block, opened by all peers through CUDA IPC.
//   [ xbuf: max_floats | rbuf: max_floats | flags: 2 x kMaxRanks u32 | counter ]
// One kernel per rank and call:  signal "my xbuf is complete" to every peer -> wait for all peers -> reduce MY
// slice over all ranks' xbuf in rank order (peer loads) and store it into EVERY rank's rbuf (peer stores) ->
// last block signals "my slice is delivered" -> wait for all peers' deliveries.  Each element is summed by
// exactly one rank in a fixed order, so the result is deterministic and bit-identical on all ranks.  NCCL needed
// 38-50 us for this 3 MB message on 8 B200s (85 us inside the step); the message is latency bound, and two
// flag round trips plus one peer read and one peer write of 1/N of the data is all it takes.
namespace {

constexpr int kMaxRanks = 16;
constexpr int kArBlocks = 148, kArThreads = 512;

struct P2P {
  bool ready = false;
  int rank = 0, world = 1;
  int64_t max_floats = 0;
  unsigned char* local = nullptr;         // this rank's exchange block
  unsigned char* peer[kMaxRanks] = {};    // everyone's block (peer[rank] == local)
  uint32_t epoch = 0;
  cudaIpcMemHandle_t handle;
};
P2P p2p;

struct ArPtrs {
  const float* x[kMaxRanks];
  float* r[kMaxRanks];
  uint32_t* flags[kMaxRanks];  // [2][kMaxRanks] per rank
};

// Returns false when a peer did not answer within `timeout` clock cycles (a peer died, or its host stalled for that
// long: UML_DP_TIMEOUT_S, default 20 s) - the caller then leaves the weights untouched and the host raises.
__device__ __forceinline__ bool wait_flags(const volatile uint32_t* f, int world, uint32_t epoch, int* failed, long long timeout) {
  // one warp polls: lane p watches rank p's flag
  const int lane = threadIdx.x & 31;
  const long long t0 = clock64();
  bool ok = false;
  while (!ok) {
    const uint32_t v = lane < world ? f[lane] : epoch;
    ok = __all_sync(0xffffffffu, static_cast<int32_t>(v - epoch) >= 0);
    if (!ok && (clock64() - t0 > timeout || *reinterpret_cast<volatile int*>(failed) != 0)) {
      *failed = 1;
      break;
    }
  }
  __threadfence_system();
  return ok;
}

// Arguments of the optional stages fused around the exchange.
struct FusedUpdate {
  const float* partials;   // phase A: xbuf = sum over n_parts split-K partials (nullptr: xbuf was written by an earlier kernel)
  int n_parts;
  int64_t stride;
  float* w;                // phase D: Adam(W) on every element from rbuf (nullptr: exchange only)
  float* m;
  float* v;
  __nv_bfloat16* shadow;
  uml::AdamArgs adam;
};

// counters[0]: blocks past phase A, counters[1]: blocks past phase B (each reset by its last block)
__global__ void __launch_bounds__(kArThreads)
    p2p_allreduce_kernel(ArPtrs P, int rank, int world, int64_t n4, uint32_t epoch, unsigned int* counters, int* failed,
                         long long timeout, FusedUpdate F) {
  using uml::adam_one;
  uint32_t* my_flags = P.flags[rank];
  __shared__ bool last;
  __shared__ bool alive;
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // ---- phase A: this rank's local gradient sum (fixed split order) into its exchange buffer
  if (F.partials) {
    float4* x = const_cast<float4*>(reinterpret_cast<const float4*>(P.x[rank]));
    for (int64_t i = tid; i < n4; i += nth) {
      float4 g = reinterpret_cast<const float4*>(F.partials)[i];
      for (int sp = 1; sp < F.n_parts; ++sp) {
        const float4 q = reinterpret_cast<const float4*>(F.partials + sp * F.stride)[i];
        g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
      }
      x[i] = g;
    }
    // device-scope fence per block; the block that signals the peers issues the system-scope fence (fences are
    // cumulative: what it observed through the counter is ordered before its flag stores)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&counters[0], 1u) == gridDim.x - 1;
    __syncthreads();
  } else {
    if (threadIdx.x == 0) last = blockIdx.x == 0;  // xbuf complete by stream order: one block signals right away
    __syncthreads();
  }
  // ---- "my xbuf is complete" -> every rank; then wait until everyone's is
  if (last) {
    if (threadIdx.x == 0 && F.partials) counters[0] = 0;
    if (threadIdx.x < world) {
      __threadfence_system();
      reinterpret_cast<volatile uint32_t*>(P.flags[threadIdx.x])[rank] = epoch;
    }
  }
  if (threadIdx.x < 32) {
    const bool ok = wait_flags(reinterpret_cast<const volatile uint32_t*>(my_flags), world, epoch, failed, timeout);
    if (threadIdx.x == 0) alive = ok;
  }
  __syncthreads();
  if (!alive) return;  // a peer is gone: no reduction from stale buffers, no update - weights, moments and shadow stay as they were
  // ---- phase B: reduce my slice in rank order (peer loads), deliver it to every rank's rbuf (peer stores)
  const int64_t per = (n4 + world - 1) / world, lo = per * rank, hi = lo + per < n4 ? lo + per : n4;
  for (int64_t i = lo + tid; i < hi; i += nth) {
    float4 v[kMaxRanks];
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p)
      if (p < world) v[p] = __ldcv(reinterpret_cast<const float4*>(P.x[p]) + i);  // all peer loads in flight at once
    float4 acc = v[0];
#pragma unroll
    for (int p = 1; p < kMaxRanks; ++p)
      if (p < world) { acc.x += v[p].x; acc.y += v[p].y; acc.z += v[p].z; acc.w += v[p].w; }
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p)
      if (p < world) reinterpret_cast<float4*>(P.r[p])[i] = acc;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&counters[1], 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) counters[1] = 0;
    if (threadIdx.x < world) {
      __threadfence_system();
      reinterpret_cast<volatile uint32_t*>(P.flags[threadIdx.x])[kMaxRanks + rank] = epoch;
    }
  }
  // ---- every rank's slice has landed in my rbuf
  if (threadIdx.x < 32) {
    const bool ok = wait_flags(reinterpret_cast<const volatile uint32_t*>(my_flags + kMaxRanks), world, epoch, failed, timeout);
    if (threadIdx.x == 0) alive = ok;
  }
  __syncthreads();
  if (!alive) return;  // (every block of this rank sees the same failure: the flag is sticky and polled by all of them)
  // ---- phase D: the optimizer update, identical on every rank (replicated weights)
  if (F.w) {
    const float4* r = reinterpret_cast<const float4*>(P.r[rank]);
    for (int64_t i = tid; i < n4; i += nth) {
      const float4 g = __ldcv(r + i);
      float4 w = reinterpret_cast<float4*>(F.w)[i];
      float4 mm = reinterpret_cast<float4*>(F.m)[i], vv = reinterpret_cast<float4*>(F.v)[i];
      w.x = adam_one(F.adam, w.x, g.x, mm.x, vv.x);
      w.y = adam_one(F.adam, w.y, g.y, mm.y, vv.y);
      w.z = adam_one(F.adam, w.z, g.z, mm.z, vv.z);
      w.w = adam_one(F.adam, w.w, g.w, mm.w, vv.w);
      reinterpret_cast<float4*>(F.w)[i] = w;
      reinterpret_cast<float4*>(F.m)[i] = mm;
      reinterpret_cast<float4*>(F.v)[i] = vv;
      if (F.shadow) reinterpret_cast<uint2*>(F.shadow)[i] = uml::pack_bf16x4(w);
    }
  }
}

inline int64_t flags_offset(int64_t max_floats) { return 2 * max_floats * static_cast<int64_t>(sizeof(float)); }

}  // namespace

// Where a rank leaves its local gradient sum (xbuf) so that uml_dp_allreduce_p2p needs no staging copy, and where
// the reduced result appears (rbuf).  NULL when the peer-memory path is not set up or the message is too long.
// (Library-internal: used by the step launcher.)
float* uml_dp_p2p_input(int64_t n) {
  return (p2p.ready && n % 4 == 0 && n <= p2p.max_floats) ? reinterpret_cast<float*>(p2p.local) : nullptr;
}
float* uml_dp_p2p_output(int64_t n) {
  return (p2p.ready && n % 4 == 0 && n <= p2p.max_floats) ? reinterpret_cast<float*>(p2p.local) + p2p.max_floats : nullptr;
}

extern "C" {

// Allocates this rank's exchange block for messages of up to max_floats and writes its 64-byte IPC handle.
int uml_dp_p2p_alloc(int64_t max_floats, void* handle_out_64_bytes) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are expected to be 64 bytes");
  UML_REQUIRE(max_floats > 0 && handle_out_64_bytes, "dp_p2p_alloc: bad arguments");
  max_floats = (max_floats + 3) / 4 * 4;
  // Re-sizing frees an IPC-exported block: every importer must have closed its mapping first (CUDA leaves freeing an
  // exported allocation that a peer still has open undefined).  The caller runs uml_dp_p2p_close_peers on every rank,
  // then a barrier, then this.
  UML_REQUIRE(!p2p.ready, "dp_p2p_alloc: peers are still mapped - call uml_dp_p2p_close_peers on every rank and synchronise the ranks first");
  if (p2p.local) {
    cudaFree(p2p.local);
    p2p = P2P();
  }
  const size_t bytes = static_cast<size_t>(flags_offset(max_floats)) + 4096;
  UML_CUDA(cudaMalloc(&p2p.local, bytes));
  UML_CUDA(cudaMemset(p2p.local, 0, bytes));
  UML_CUDA(cudaDeviceSynchronize());
  UML_CUDA(cudaIpcGetMemHandle(&p2p.handle, p2p.local));
  p2p.max_floats = max_floats;
  memcpy(handle_out_64_bytes, &p2p.handle, 64);
  return 0;
}

// Unmaps every peer's block (first half of a re-size; a no-op when nothing is mapped).  Synchronises the device.
int uml_dp_p2p_close_peers(void) {
  if (!p2p.ready) return 0;
  UML_CUDA(cudaDeviceSynchronize());
  for (int p = 0; p < p2p.world; ++p)
    if (p != p2p.rank && p2p.peer[p]) {
      cudaIpcCloseMemHandle(p2p.peer[p]);
      p2p.peer[p] = nullptr;
    }
  p2p.ready = false;
  return 0;
}

// handles: world x 64 bytes in rank order (all-gathered by the caller).  Opens every peer's block.
int uml_dp_p2p_open(const void* handles, int32_t rank, int32_t world) {
  UML_REQUIRE(handles && p2p.local && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "dp_p2p_open: bad arguments");
  for (int p = 0; p < world; ++p) {
    if (p == rank) {
      p2p.peer[p] = p2p.local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const unsigned char*>(handles) + 64 * p, 64);
    void* ptr = nullptr;
    UML_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p2p.peer[p] = static_cast<unsigned char*>(ptr);
  }
  p2p.rank = rank;
  p2p.world = world;
  p2p.epoch = 0;
  p2p.ready = true;
  return 0;
}


static int p2p_launch(int64_t n, const FusedUpdate& F, void* stream) {
  UML_REQUIRE(p2p.ready && n > 0 && n % 4 == 0 && n <= p2p.max_floats, "dp_allreduce_p2p: not initialised or bad size");
  ArPtrs P;
  memset(&P, 0, sizeof(P));
  for (int p = 0; p < p2p.world; ++p) {
    P.x[p] = reinterpret_cast<const float*>(p2p.peer[p]);
    P.r[p] = reinterpret_cast<float*>(p2p.peer[p]) + p2p.max_floats;
    P.flags[p] = reinterpret_cast<uint32_t*>(p2p.peer[p] + flags_offset(p2p.max_floats));
  }
  uint32_t* my = reinterpret_cast<uint32_t*>(p2p.local + flags_offset(p2p.max_floats));
  unsigned int* counters = my + 2 * kMaxRanks;
  int* failed = reinterpret_cast<int*>(my + 2 * kMaxRanks + 2);
  ++p2p.epoch;
  static long long timeout = 0;
  if (timeout == 0) {
    const char* e = getenv("UML_DP_TIMEOUT_S");
    const double sec = e ? atof(e) : 20.0;
    timeout = static_cast<long long>((sec > 0.01 ? sec : 0.01) * 2.0e9);  // clock64 ticks at ~2 GHz
  }
  p2p_allreduce_kernel<<<kArBlocks, kArThreads, 0, uml::as_stream(stream)>>>(P, p2p.rank, p2p.world, n / 4, p2p.epoch, counters,
                                                                              failed, timeout, F);
  UML_CUDA(cudaGetLastError());
  return 0;
}

// Sum of every rank's input half [0..n) -> every rank's output half [0..n).  n % 4 == 0.
int uml_dp_allreduce_p2p(int64_t n, void* stream) {
  FusedUpdate F;
  memset(&F, 0, sizeof(F));
  return p2p_launch(n, F, stream);
}

// The data-parallel tail of a step in ONE kernel: local split-K sum of the dW partials -> exchange over NVLink
// peer memory -> Adam/AdamW on every rank (+ bf16 weight shadow).  partials == NULL: the local sum already is in the
// exchange buffer (fp32 path).
int uml_dp_fused_adam_update(const float* partials, int32_t n_splits, int64_t split_stride, int64_t n, float* p, float* m,
                             float* v, double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                             int32_t decoupled, uint16_t* p_bf16, void* stream) {
  UML_REQUIRE(p && m && v && step >= 1 && (!partials || (n_splits >= 1 && split_stride >= n && split_stride % 4 == 0)),
              "dp_fused_adam_update: bad arguments");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  UML_REQUIRE(al16(p) && al16(m) && al16(v) && (!partials || al16(partials)) && (!p_bf16 || (reinterpret_cast<uintptr_t>(p_bf16) & 7u) == 0),
              "dp_fused_adam_update: buffers must be 16-byte aligned");
  FusedUpdate F;
  memset(&F, 0, sizeof(F));
  F.partials = partials;
  F.n_parts = n_splits;
  F.stride = split_stride;
  F.w = p; F.m = m; F.v = v;
  F.shadow = reinterpret_cast<__nv_bfloat16*>(p_bf16);
  F.adam = uml::make_adam(lr, beta1, beta2, eps, weight_decay, step, decoupled);
  return p2p_launch(n, F, stream);
}

// 1 if a peer stopped answering during an earlier uml_dp_allreduce_p2p (synchronises the device)
int uml_dp_p2p_failed(void) {
  if (!p2p.ready) return 0;
  int f = 0;
  const uint32_t* my = reinterpret_cast<const uint32_t*>(p2p.local + flags_offset(p2p.max_floats));
  if (cudaMemcpy(&f, my + 2 * kMaxRanks + 2, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
  return f != 0;
}

int uml_dp_unique_id(void* out_128_bytes) {
  static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is expected to be 128 bytes");
  UML_REQUIRE(out_128_bytes, "dp_unique_id: null buffer");
  if (load_nccl()) return 1;
  ncclUniqueId id;
  UML_NCCL(g.GetUniqueId(&id));
  memcpy(out_128_bytes, &id, sizeof(id));
  return 0;
}

int uml_dp_init(const void* id_128_bytes, int32_t rank, int32_t world) {
  UML_REQUIRE(id_128_bytes && world >= 1 && rank >= 0 && rank < world, "dp_init: bad arguments");
  if (load_nccl()) return 1;
  if (g.comm) {
    g.CommDestroy(g.comm);
    g.comm = nullptr;
  }
  ncclUniqueId id;
  memcpy(&id, id_128_bytes, sizeof(id));
  UML_NCCL(g.CommInitRank(&g.comm, world, id, rank));
  g.world = world;
  return 0;
}

int uml_dp_allreduce_f32(float* buf, int64_t n, void* stream) {
  UML_REQUIRE(g.comm, "dp_allreduce: uml_dp_init has not been called");
  UML_REQUIRE(buf && n >= 0, "dp_allreduce: bad arguments");
  if (n == 0 || g.world == 1) return 0;
  UML_NCCL(g.AllReduce(buf, buf, static_cast<size_t>(n), ncclFloat, ncclSum, g.comm, uml::as_stream(stream)));
  return 0;
}

int uml_dp_shutdown(void) {
  if (g.comm) {
    g.CommDestroy(g.comm);
    g.comm = nullptr;
  }
  g.world = 1;
  return 0;
}

}  // extern "C"
