// Data-parallel plumbing: one NCCL communicator per process (one process per GPU), used by the step
// launcher to sum the head gradient over the ranks without returning to Python.  libnccl is resolved at
// run time with dlopen (the same libnccl.so.2 torch.distributed already loaded), so the library has no
// link-time dependency on it and still loads on a box without NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

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

extern "C" {

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
