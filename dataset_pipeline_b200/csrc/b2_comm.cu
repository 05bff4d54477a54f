// Data-parallel exchange owned by the library: one NCCL communicator per process (one process per GPU), used for the single
// collective of the ICP path — a sum-allreduce of the packed normal equations per pass, over NVLink / NVSwitch.
// NCCL is bound at run time with dlopen("libnccl.so.2"): inside a PyTorch process this resolves to the NCCL torch already
// loaded (no second copy), in a C++ host to the system library. The few declarations needed are restated here so the build has
// no NCCL header dependency.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "b2_common.cuh"

namespace b2 {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0, kNcclSum = 0, kNcclUint8 = 1, kNcclInt32 = 2, kNcclFloat32 = 7, kNcclFloat64 = 8 };

struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return;
    api.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(lib, "ncclCommInitRank");
    api.CommDestroy = (int (*)(ncclComm_t))dlsym(lib, "ncclCommDestroy");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllReduce");
    api.Broadcast = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclBroadcast");
    api.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce;
  });
  return api;
}

}  // namespace b2

using namespace b2;

struct b2_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

extern "C" {

int b2_comm_unique_id(unsigned char id[128]) {
  if (!id) return set_error(B2_ERR_ARG, "null");
  if (!nccl().ok) return set_error(B2_ERR_COMM, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "symbols missing");
  ncclUniqueId u;
  const int rc = nccl().GetUniqueId(&u);
  if (rc != kNcclSuccess) return set_error(B2_ERR_COMM, "ncclGetUniqueId failed: %s", nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
  std::memcpy(id, u.internal, 128);
  return B2_OK;
}

int b2_comm_create(int rank, int world_size, const unsigned char id[128], int device, b2_comm** out) {
  if (!out || !id || world_size < 1 || rank < 0 || rank >= world_size) return set_error(B2_ERR_ARG, "bad argument");
  *out = nullptr;
  if (!nccl().ok) return set_error(B2_ERR_COMM, "libnccl.so.2 could not be loaded");
  int dev = 0, sms = 0;
  B2_TRY(select_device(device, &dev, &sms));
  b2_comm* c = new b2_comm();
  c->rank = rank; c->world = world_size; c->device = dev;
  ncclUniqueId u; std::memcpy(u.internal, id, 128);
  const int rc = nccl().CommInitRank(&c->comm, world_size, u, rank);
  if (rc != kNcclSuccess) { delete c; return set_error(B2_ERR_COMM, "ncclCommInitRank failed: %s", nccl().GetErrorString ? nccl().GetErrorString(rc) : "?"); }
  *out = c;
  return B2_OK;
}

int b2_comm_destroy(b2_comm* c) {
  if (!c) return B2_OK;
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
  return B2_OK;
}

// In-place sum-allreduce of `count` doubles on `stream` (used by b2_icp_run; exported so a host can exercise the path).
int b2_comm_allreduce_f64(b2_comm* c, double* buf_dev, size_t count, void* stream) {
  if (!c || !buf_dev) return set_error(B2_ERR_ARG, "null");
  const int rc = nccl().AllReduce(buf_dev, buf_dev, count, kNcclFloat64, kNcclSum, c->comm, (cudaStream_t)stream);
  if (rc != kNcclSuccess) return set_error(B2_ERR_COMM, "ncclAllReduce failed: %s", nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
  return B2_OK;
}

int b2_comm_allreduce(b2_comm* c, void* buf_dev, size_t count, int dtype, void* stream) {
  if (!c || !buf_dev) return set_error(B2_ERR_ARG, "null");
  const int t = dtype == B2_F64 ? kNcclFloat64 : dtype == B2_F32 ? kNcclFloat32 : dtype == B2_I32 ? kNcclInt32 : -1;
  if (t < 0) return set_error(B2_ERR_ARG, "bad dtype");
  if (count == 0) return B2_OK;
  const int rc = nccl().AllReduce(buf_dev, buf_dev, count, t, kNcclSum, c->comm, (cudaStream_t)stream);
  if (rc != kNcclSuccess) return set_error(B2_ERR_COMM, "ncclAllReduce failed: %s", nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
  return B2_OK;
}

// In-place broadcast of `bytes` bytes from rank `root` on `stream` (the sharded upload of b2_icp_add_cloud: the owner's copy of a scan
// reaches the other GPUs over NVLink instead of every process pushing every scan through PCIe).
int b2_comm_broadcast(b2_comm* c, void* buf_dev, size_t bytes, int root, void* stream) {
  if (!c || (!buf_dev && bytes)) return set_error(B2_ERR_ARG, "null");
  if (root < 0 || root >= c->world) return set_error(B2_ERR_ARG, "bad root %d", root);
  if (!nccl().Broadcast) return set_error(B2_ERR_COMM, "ncclBroadcast not found in libnccl");
  if (bytes == 0) return B2_OK;
  const int rc = nccl().Broadcast(buf_dev, buf_dev, bytes, kNcclUint8, root, c->comm, (cudaStream_t)stream);
  if (rc != kNcclSuccess) return set_error(B2_ERR_COMM, "ncclBroadcast failed: %s", nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
  return B2_OK;
}

int b2_comm_info(b2_comm* c, int* rank, int* world_size) {
  if (!c) return set_error(B2_ERR_ARG, "null");
  if (rank) *rank = c->rank;
  if (world_size) *world_size = c->world;
  return B2_OK;
}

}  // extern "C"
