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
//   [ xbuf: max_floats | out: max_floats | recv: flagged slots | rll: flagged slots | failed ]
// One kernel per rank and call.  No grid-wide step, no fence and no separate flag inside it: block b of every rank
// owns chunk b of the message from the first load to the optimizer update, and the data carries its own validity -
// it travels in 16-byte SLOTS of three floats and the call's epoch, each written by one 16-byte vector store of one
// thread (the memory system delivers such a store whole - the LL protocol of NCCL, at 16 instead of 8 bytes: 75 %
// payload), so a reader that finds the epoch in a slot has the slot's data.  A thread moves UNITS of three float4
// (twelve floats = four slots).
//   push   : block b sums its chunk of the local gradient and stores sub-slice j of it, as slots, straight into rank
//            j's recv area (peer stores);
//   reduce : block b polls the world copies of ITS sub-slice (local loads) unit by unit, adds them in rank order and
//            stores the sum, as slots again, into EVERY rank's rll area (peer stores);
//   update : block b polls the units of chunk b in its rll area and applies Adam as they arrive.
// Two one-way NVLink hops per chunk.  (Round 1's kernel: grid-wide "last block" hand-offs, system fences, separate
// flags and pulled loads - two flag round trips plus a load round trip, +31 us over the single-GPU update at any rank
// count; a push version with fences and per-block flags still needed 20 us for the exchange alone, as NCCL does; 128-
// byte lines with one flag per line - LL128 - delivered torn lines here: parity failed.)
// Each element is summed by exactly one rank in a fixed order, so the result is deterministic and bit-identical on
// all ranks.
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
  const float* x;              // this rank's local gradient sum (when no split-K partials are given)
  float* out;                  // this rank's plain result (exchange-only calls)
  uint4* recv[kMaxRanks];      // per rank: [source rank][block][unit][4] slots
  uint4* rll[kMaxRanks];       // per rank: [owner rank][block][unit][4] slots
};

// Arguments of the optional stages fused around the exchange.
struct FusedUpdate {
  const float* partials;   // push: chunk = sum over n_parts split-K partials (nullptr: the local sum is in xbuf)
  int n_parts;
  int64_t stride;
  float* w;                // update: Adam(W) on every element (nullptr: exchange only, plain result in `out`)
  float* m;
  float* v;
  __nv_bfloat16* shadow;
  uml::AdamArgs adam;
};

struct Unit {
  float4 a, b, c;
};

__device__ __forceinline__ void st_slot(uint4* p, float x, float y, float z, uint32_t e) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(x)), "r"(__float_as_uint(y)),
               "r"(__float_as_uint(z)), "r"(e)
               : "memory");
}
__device__ __forceinline__ uint4 ld_slot(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// The four slots of unit kk of a region sit 32 slots apart, at (kk & ~31) * 4 + part * 32 + (kk & 31): the 32 units of
// a warp then fill 512 contiguous bytes with each store instruction (full lines on NVLink instead of 16-byte packets).
__device__ __forceinline__ int64_t unit_slot(int64_t kk) { return (kk & ~int64_t(31)) * 4 + (kk & 31); }
__device__ __forceinline__ void st_unit(uint4* p, const Unit& u, uint32_t e) {
  st_slot(p, u.a.x, u.a.y, u.a.z, e);
  st_slot(p + 32, u.a.w, u.b.x, u.b.y, e);
  st_slot(p + 64, u.b.z, u.b.w, u.c.x, e);
  st_slot(p + 96, u.c.y, u.c.z, u.c.w, e);
}
// Polls the four slots of a unit until all carry `epoch`.  false: the peer did not deliver within `timeout` clock
// cycles (it died, or its host stalled that long: UML_DP_TIMEOUT_S, default 20 s) - the kernel then leaves the
// weights untouched and the host raises.
__device__ __forceinline__ bool ld_unit(const uint4* p, uint32_t epoch, Unit& u, int* failed, long long timeout) {
  unsigned polls = 0;
  long long t0 = 0;
  for (;;) {
    const uint4 s0 = ld_slot(p), s1 = ld_slot(p + 32), s2 = ld_slot(p + 64), s3 = ld_slot(p + 96);
    if (s0.w == epoch && s1.w == epoch && s2.w == epoch && s3.w == epoch) {
      u.a = make_float4(__uint_as_float(s0.x), __uint_as_float(s0.y), __uint_as_float(s0.z), __uint_as_float(s1.x));
      u.b = make_float4(__uint_as_float(s1.y), __uint_as_float(s1.z), __uint_as_float(s2.x), __uint_as_float(s2.y));
      u.c = make_float4(__uint_as_float(s2.z), __uint_as_float(s3.x), __uint_as_float(s3.y), __uint_as_float(s3.z));
      return true;
    }
    if ((++polls & 255u) == 0) {
      if (t0 == 0) t0 = clock64();
      if (clock64() - t0 > timeout || *reinterpret_cast<volatile int*>(failed) != 0) {
        *failed = 1;
        return false;
      }
    }
  }
}
__device__ __forceinline__ void add4(float4& a, const float4& b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}

__global__ void __launch_bounds__(kArThreads)
    p2p_allreduce_kernel(ArPtrs P, int rank, int world, int64_t n4, int64_t sub, uint32_t epoch, int* failed, long long timeout,
                         FusedUpdate F) {
  using uml::adam_one;
  const int b = blockIdx.x;
  const int64_t lo = sub * world * b;                // first unit of chunk b; unit u covers float4 [3u, 3u + 3)
  const int64_t sub32 = (sub + 31) & ~int64_t(31);   // a sub-slice's region holds whole warps of units
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  // ---- push: local gradient sum of chunk b (fixed split order), sub-slice j -> rank j, as flagged slots
  {
    const float4* src = reinterpret_cast<const float4*>(F.partials ? F.partials : P.x);
    for (int64_t k = threadIdx.x; k < sub32 * world; k += blockDim.x) {
      const int j = static_cast<int>(k / sub32);
      const int64_t kk = k - j * sub32, i0 = (lo + j * sub + kk) * 3;
      if (kk >= sub || i0 >= n4) continue;
      Unit u;
      float4* part[3] = {&u.a, &u.b, &u.c};
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        float4 g = zero;
        if (i0 + e < n4) {
          g = F.partials ? src[i0 + e] : __ldcv(src + i0 + e);
          for (int sp = 1; F.partials && sp < F.n_parts; ++sp) add4(g, reinterpret_cast<const float4*>(F.partials + sp * F.stride)[i0 + e]);
        }
        *part[e] = g;
      }
      st_unit(P.recv[j] + (static_cast<int64_t>(rank) * kArBlocks + b) * sub32 * 4 + unit_slot(kk), u, epoch);
    }
  }
  // ---- reduce: my sub-slice of chunk b, summed in rank order from the copies the ranks pushed here -> everyone's rll
  for (int64_t kk = threadIdx.x; kk < sub; kk += blockDim.x) {
    if ((lo + sub * rank + kk) * 3 >= n4) break;
    Unit acc, v;
    for (int p = 0; p < world; ++p) {
      if (!ld_unit(P.recv[rank] + (static_cast<int64_t>(p) * kArBlocks + b) * sub32 * 4 + unit_slot(kk), epoch, v, failed, timeout)) return;
      if (p == 0) {
        acc = v;
      } else {
        add4(acc.a, v.a); add4(acc.b, v.b); add4(acc.c, v.c);
      }
    }
    for (int q = 0; q < world; ++q) st_unit(P.rll[q] + (static_cast<int64_t>(rank) * kArBlocks + b) * sub32 * 4 + unit_slot(kk), acc, epoch);
  }
  // ---- update: chunk b of the reduced gradient arrives unit by unit from its owners; the optimizer step is identical
  //      on every rank (replicated weights)
  for (int64_t k = threadIdx.x; k < sub32 * world; k += blockDim.x) {
    const int q = static_cast<int>(k / sub32);
    const int64_t kk = k - q * sub32, i0 = (lo + q * sub + kk) * 3;
    if (kk >= sub || i0 >= n4) continue;
    Unit u;
    if (!ld_unit(P.rll[rank] + (static_cast<int64_t>(q) * kArBlocks + b) * sub32 * 4 + unit_slot(kk), epoch, u, failed, timeout)) return;
    const float4 gs[3] = {u.a, u.b, u.c};
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      const int64_t i = i0 + e;
      if (i >= n4) break;
      const float4 g = gs[e];
      if (F.w) {
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
      } else {
        reinterpret_cast<float4*>(P.out)[i] = g;
      }
    }
  }
}

// layout of an exchange block (bytes): xbuf, out, recv slots, rll slots, failed word
inline int64_t ll_slots(int64_t max_floats) { return (max_floats / 3 + 4) + 4 * 33 * static_cast<int64_t>(kArBlocks) * kMaxRanks + 64; }
inline int64_t recv_offset(int64_t max_floats) { return 2 * max_floats * static_cast<int64_t>(sizeof(float)); }
inline int64_t rll_offset(int64_t max_floats) { return recv_offset(max_floats) + ll_slots(max_floats) * 16; }
constexpr int64_t kFlagWords = 0;  // (the slots carry their own flags) then: the sticky "failed" word
inline int64_t flags_offset(int64_t max_floats) { return rll_offset(max_floats) + ll_slots(max_floats) * 16; }

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
  const size_t bytes = static_cast<size_t>(flags_offset(max_floats)) + kFlagWords * sizeof(uint32_t) + 4096;
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
  P.x = reinterpret_cast<const float*>(p2p.local);
  P.out = reinterpret_cast<float*>(p2p.local) + p2p.max_floats;
  for (int p = 0; p < p2p.world; ++p) {
    P.recv[p] = reinterpret_cast<uint4*>(p2p.peer[p] + recv_offset(p2p.max_floats));
    P.rll[p] = reinterpret_cast<uint4*>(p2p.peer[p] + rll_offset(p2p.max_floats));
  }
  uint32_t* my = reinterpret_cast<uint32_t*>(p2p.local + flags_offset(p2p.max_floats));
  int* failed = reinterpret_cast<int*>(my + kFlagWords);
  ++p2p.epoch;
  static long long timeout = 0;
  if (timeout == 0) {
    const char* e = getenv("UML_DP_TIMEOUT_S");
    const double sec = e ? atof(e) : 20.0;
    timeout = static_cast<long long>((sec > 0.01 ? sec : 0.01) * 2.0e9);  // clock64 ticks at ~2 GHz
  }
  // block b owns units [b * sub * world, (b + 1) * sub * world) of the message (a unit = three float4), rank j reduces
  // its j-th sub-slice
  const int64_t n4 = n / 4, units = (n4 + 2) / 3;
  const int64_t sub = (units + static_cast<int64_t>(kArBlocks) * p2p.world - 1) / (static_cast<int64_t>(kArBlocks) * p2p.world);
  p2p_allreduce_kernel<<<kArBlocks, kArThreads, 0, uml::as_stream(stream)>>>(P, p2p.rank, p2p.world, n4, sub, p2p.epoch, failed,
                                                                              timeout, F);
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
  if (cudaMemcpy(&f, my + kFlagWords, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
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
