// comm.cu — the path's only collective behind the C ABI: one NCCL all-reduce (sum) of the int64
// (I, U, T) count buffers over the ranks that shard the clips.  The reference never reduces counts (it averages per-rank
// mIoU scalars, base/foundation.py:166-168); summing the integer counts reproduces the single-process result exactly.
//
// NCCL is not linked: a host process that does distributed work has it loaded already (PyTorch ships libnccl.so.2), so
// the five entry points are resolved with dlopen/dlsym at the first call and a process without NCCL gets a clean error
// instead of a load-time failure of libfuvs.so.  One communicator per process (one process per GPU).
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "fuvs_common.cuh"

namespace fuvs {
namespace {

struct NcclUniqueId { char internal[128]; };           // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
using NcclComm = void*;
constexpr int kNcclInt64 = 4;                           // ncclInt64
constexpr int kNcclSum = 0;                             // ncclSum

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi g_api;
NcclComm g_comm = nullptr;
int g_world = 0;
std::mutex g_mu;

bool load_api() {
  if (g_api.ok) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);         // the copy the host process already uses
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  g_api.GetUniqueId = reinterpret_cast<decltype(g_api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  g_api.CommInitRank = reinterpret_cast<decltype(g_api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  g_api.AllReduce = reinterpret_cast<decltype(g_api.AllReduce)>(dlsym(h, "ncclAllReduce"));
  g_api.CommDestroy = reinterpret_cast<decltype(g_api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  g_api.GetErrorString = reinterpret_cast<decltype(g_api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  g_api.ok = g_api.GetUniqueId && g_api.CommInitRank && g_api.AllReduce && g_api.CommDestroy && g_api.GetErrorString;
  return g_api.ok;
}

int nccl_error(const char* what, int rc) {
  return set_error(FUVS_ECUDA, "%s: NCCL error %d (%s)", what, rc, g_api.GetErrorString ? g_api.GetErrorString(rc) : "?");
}

}  // namespace
}  // namespace fuvs

extern "C" int fuvs_comm_unique_id(void* id128) {
  using namespace fuvs;
  std::lock_guard<std::mutex> lock(g_mu);
  if (!id128) return set_error(FUVS_EINVAL, "comm_unique_id: NULL buffer (128 bytes)");
  if (!load_api()) return set_error(FUVS_ENODEV, "comm: libnccl.so.2 is not available in this process");
  NcclUniqueId id;
  if (int rc = g_api.GetUniqueId(&id)) return nccl_error("comm_unique_id", rc);
  std::memcpy(id128, &id, sizeof(id));
  return FUVS_OK;
}

extern "C" int fuvs_comm_init(const void* id128, int rank, int world_size) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  std::lock_guard<std::mutex> lock(g_mu);
  if (!id128 || world_size < 1 || rank < 0 || rank >= world_size)
    return set_error(FUVS_EINVAL, "comm_init: bad arguments rank=%d world_size=%d", rank, world_size);
  if (g_comm) return set_error(FUVS_EINVAL, "comm_init: a communicator exists already (fuvs_comm_destroy first)");
  if (!load_api()) return set_error(FUVS_ENODEV, "comm: libnccl.so.2 is not available in this process");
  NcclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  if (int rc = g_api.CommInitRank(&g_comm, world_size, id, rank)) {
    g_comm = nullptr;
    return nccl_error("comm_init", rc);
  }
  g_world = world_size;
  return FUVS_OK;
}

extern "C" int fuvs_comm_world_size(void) { return fuvs::g_comm ? fuvs::g_world : 0; }

extern "C" int fuvs_allreduce_counts(long long* counts, long long n, fuvs_stream_t stream) {
  using namespace fuvs;
  if (!counts || n < 0) return set_error(FUVS_EINVAL, "allreduce_counts: bad arguments n=%lld", n);
  if (!g_comm) return set_error(FUVS_EINVAL, "allreduce_counts: no communicator (fuvs_comm_init)");
  if (n == 0) return FUVS_OK;
  if (int rc = g_api.AllReduce(counts, counts, static_cast<size_t>(n), kNcclInt64, kNcclSum, g_comm,
                               static_cast<cudaStream_t>(stream)))
    return nccl_error("allreduce_counts", rc);
  count_launch();
  return FUVS_OK;
}

extern "C" int fuvs_comm_destroy(void) {
  using namespace fuvs;
  std::lock_guard<std::mutex> lock(g_mu);
  if (!g_comm) return FUVS_OK;
  const int rc = g_api.CommDestroy(g_comm);
  g_comm = nullptr;
  g_world = 0;
  return rc ? nccl_error("comm_destroy", rc) : FUVS_OK;
}
