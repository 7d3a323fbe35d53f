// comm.cpp -- NCCL plumbing for the row-partitioned multi-GPU path.
// One process per GPU; rank 0 creates the ncclUniqueId, the launcher (bench.py over
// torch.distributed, or any other out-of-band channel) shares its 128 bytes, every rank passes
// them to eigkl_create.
#include "internal.h"
#ifdef EIGKL_WITH_NCCL
#include <nccl.h>
#endif

namespace eigkl {

#ifdef EIGKL_WITH_NCCL
#define EIGKL_NCCL(call)                                                                            \
  do {                                                                                              \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != ncclSuccess)                                                                         \
      throw Error(EIGKL_E_NCCL, std::string("NCCL error in ") + __FILE__ + ":" + std::to_string(__LINE__) + ": " + ncclGetErrorString(r__)); \
  } while (0)

void comm_unique_id(void *id128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  EIGKL_NCCL(ncclGetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
}
void comm_init(eigkl_handle *h) {
  EIGKL_REQUIRE(h->opts.nccl_unique_id != nullptr, EIGKL_E_ARG, "nranks > 1 needs eigkl_opts.nccl_unique_id");
  ncclUniqueId id;
  memcpy(&id, h->opts.nccl_unique_id, sizeof(id));
  ncclComm_t comm;
  EIGKL_NCCL(ncclCommInitRank(&comm, h->opts.nranks, id, h->opts.rank));
  h->nccl_comm = comm;
}
void comm_destroy(eigkl_handle *h) {
  if (h->nccl_comm) ncclCommDestroy((ncclComm_t)h->nccl_comm);
  h->nccl_comm = nullptr;
}
#else
void comm_unique_id(void *) { throw Error(EIGKL_E_NCCL, "libeigkl was built without NCCL"); }
void comm_init(eigkl_handle *) { throw Error(EIGKL_E_NCCL, "libeigkl was built without NCCL"); }
void comm_destroy(eigkl_handle *) {}
#endif

}  // namespace eigkl
