// comm.cpp -- NCCL plumbing for the row-partitioned multi-GPU path.
// One process per GPU; rank 0 creates the ncclUniqueId, the launcher (bench.py over
// torch.distributed, or any other out-of-band channel) shares its 128 bytes, every rank passes
// them to eigkl_create.
#include "internal.h"
#ifdef EIGKL_WITH_NCCL
#include <nccl.h>
#include <dlfcn.h>
#endif

namespace eigkl {

#ifdef EIGKL_WITH_NCCL
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  bool ok = false;
};
NcclApi &nccl();
#define EIGKL_NCCL(call)                                                                            \
  do {                                                                                              \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != ncclSuccess)                                                                         \
      throw Error(EIGKL_E_NCCL, std::string("NCCL error in ") + __FILE__ + ":" + std::to_string(__LINE__) + ": " + nccl().GetErrorString(r__)); \
  } while (0)
#endif

#ifdef EIGKL_WITH_NCCL
// NCCL is bound at run time (dlopen by SONAME) instead of at link time: inside a Python process
// that also imports torch there must be exactly one libnccl.so.2, torch's bundled one; a link-time
// dependency on the system copy would be loaded first and break torch's own import.
NcclApi &nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (lib) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(lib, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(lib, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(lib, "ncclAllReduce");
      api.AllGather = (decltype(api.AllGather))dlsym(lib, "ncclAllGather");
      api.Broadcast = (decltype(api.Broadcast))dlsym(lib, "ncclBroadcast");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(lib, "ncclGetErrorString");
      api.GetVersion = (decltype(api.GetVersion))dlsym(lib, "ncclGetVersion");
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather &&
               api.Broadcast && api.GetErrorString;
    }
  }
  if (!api.ok) throw Error(EIGKL_E_NCCL, "libnccl.so.2 could not be loaded");
  return api;
}

void comm_unique_id(void *id128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  EIGKL_NCCL(nccl().GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
}
void comm_init(eigkl_handle *h) {
  EIGKL_REQUIRE(h->opts.nccl_unique_id != nullptr, EIGKL_E_ARG, "nranks > 1 needs eigkl_opts.nccl_unique_id");
  ncclUniqueId id;
  memcpy(&id, h->opts.nccl_unique_id, sizeof(id));
  ncclComm_t comm;
  EIGKL_NCCL(nccl().CommInitRank(&comm, h->opts.nranks, id, h->opts.rank));
  h->nccl_comm = comm;
}
void comm_destroy(eigkl_handle *h) {
  if (h->nccl_comm) nccl().CommDestroy((ncclComm_t)h->nccl_comm);
  h->nccl_comm = nullptr;
}
void comm_allreduce_sum_f64(eigkl_handle *h, double *buf, size_t count) {
  EIGKL_NCCL(nccl().AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
}
void comm_allreduce_max_u64(eigkl_handle *h, unsigned long long *buf, size_t count) {
  EIGKL_NCCL(nccl().AllReduce(buf, buf, count, ncclUint64, ncclMax, (ncclComm_t)h->nccl_comm, h->stream));
}
void comm_allgather_f64(eigkl_handle *h, const double *send, double *recv, size_t count_per_rank) {
  EIGKL_NCCL(nccl().AllGather(send, recv, count_per_rank, ncclDouble, (ncclComm_t)h->nccl_comm, h->stream));
}
void comm_broadcast_bytes(eigkl_handle *h, void *buf, size_t bytes, int root) {
  EIGKL_NCCL(nccl().Broadcast(buf, buf, bytes, ncclChar, root, (ncclComm_t)h->nccl_comm, h->stream));
}
void comm_allgather_bytes(eigkl_handle *h, const void *send, void *recv, size_t bytes_per_rank) {
  EIGKL_NCCL(nccl().AllGather(send, recv, bytes_per_rank, ncclChar, (ncclComm_t)h->nccl_comm, h->stream));
}
void comm_allreduce_min_i32(eigkl_handle *h, int32_t *buf, size_t count) {
  EIGKL_NCCL(nccl().AllReduce(buf, buf, count, ncclInt32, ncclMin, (ncclComm_t)h->nccl_comm, h->stream));
}
#else
void comm_unique_id(void *) { throw Error(EIGKL_E_NCCL, "libeigkl was built without NCCL"); }
void comm_init(eigkl_handle *) { throw Error(EIGKL_E_NCCL, "libeigkl was built without NCCL"); }
void comm_destroy(eigkl_handle *) {}
void comm_allreduce_sum_f64(eigkl_handle *, double *, size_t) { throw Error(EIGKL_E_NCCL, "libeigkl was built without NCCL"); }
void comm_allreduce_max_u64(eigkl_handle *, unsigned long long *, size_t) { throw Error(EIGKL_E_NCCL, "libeigkl was built without NCCL"); }
void comm_allgather_f64(eigkl_handle *, const double *, double *, size_t) { throw Error(EIGKL_E_NCCL, "libeigkl was built without NCCL"); }
void comm_broadcast_bytes(eigkl_handle *, void *, size_t, int) { throw Error(EIGKL_E_NCCL, "libeigkl was built without NCCL"); }
void comm_allgather_bytes(eigkl_handle *, const void *, void *, size_t) { throw Error(EIGKL_E_NCCL, "libeigkl was built without NCCL"); }
void comm_allreduce_min_i32(eigkl_handle *, int32_t *, size_t) { throw Error(EIGKL_E_NCCL, "libeigkl was built without NCCL"); }
#endif

}  // namespace eigkl
