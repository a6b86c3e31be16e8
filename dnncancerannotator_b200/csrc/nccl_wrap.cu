// Thin NCCL wrappers of the C ABI (`dnnca_nccl_*`): the exchange step of the path -- MirroredStrategy's
// SUM all-reduce of the parameter gradients (annotator/engine.py:260-263) -- for hosts that bind libdnnca.so without
// torch.distributed.  The Python host of this repo keeps torch.distributed for the plumbing (same NCCL underneath);
// tests/test_dp_nccl.py drives these entry points with two ranks and checks them against torch's collective.
//
// libnccl is NOT a link-time dependency: it is opened on first use (libnccl.so.2; a process that already loaded NCCL --
// e.g. through torch -- gets that same copy by SONAME), so single-GPU users never need it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "common.cuh"

namespace dnnca {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!api.handle) {
    set_error("dnnca_nccl: cannot open libnccl.so.2 (%s)", dlerror());
    return nullptr;
  }
#define DNNCA_NCCL_SYM(field, name)                                          \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name)); \
  if (!api.field) {                                                          \
    set_error("dnnca_nccl: symbol %s missing in libnccl", name);             \
    return nullptr;                                                          \
  }
  DNNCA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  DNNCA_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  DNNCA_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  DNNCA_NCCL_SYM(AllReduce, "ncclAllReduce")
  DNNCA_NCCL_SYM(Broadcast, "ncclBroadcast")
  DNNCA_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef DNNCA_NCCL_SYM
  api.ok = true;
  return &api;
}

static int nccl_fail(NcclApi* api, ncclResult_t r, const char* what) {
  set_error("%s: NCCL error %d (%s)", what, (int)r, api->GetErrorString(r));
  return DNNCA_ERR_NCCL;
}

}  // namespace dnnca

using namespace dnnca;

static_assert(sizeof(ncclUniqueId) == DNNCA_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");

extern "C" int dnnca_nccl_unique_id(unsigned char* id) {
  DNNCA_CHECK_ARG(id, "nccl_unique_id: bad arguments");
  NcclApi* api = nccl_api();
  if (!api) return DNNCA_ERR_NCCL;
  ncclUniqueId u;
  ncclResult_t r = api->GetUniqueId(&u);
  if (r != ncclSuccess) return nccl_fail(api, r, "nccl_unique_id");
  memcpy(id, &u, sizeof(u));
  return DNNCA_OK;
}

extern "C" int dnnca_nccl_comm_init_rank(void** comm, int nranks, const unsigned char* id, int rank) {
  DNNCA_CHECK_ARG(comm && id && nranks >= 1 && rank >= 0 && rank < nranks, "nccl_comm_init_rank: bad arguments");
  NcclApi* api = nccl_api();
  if (!api) return DNNCA_ERR_NCCL;
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t c = nullptr;
  ncclResult_t r = api->CommInitRank(&c, nranks, u, rank);          // binds to the calling thread's current device
  if (r != ncclSuccess) return nccl_fail(api, r, "nccl_comm_init_rank");
  *comm = c;
  return DNNCA_OK;
}

extern "C" int dnnca_nccl_comm_destroy(void* comm) {
  if (!comm) return DNNCA_OK;
  NcclApi* api = nccl_api();
  if (!api) return DNNCA_ERR_NCCL;
  ncclResult_t r = api->CommDestroy(reinterpret_cast<ncclComm_t>(comm));
  return r == ncclSuccess ? DNNCA_OK : nccl_fail(api, r, "nccl_comm_destroy");
}

extern "C" int dnnca_nccl_allreduce_bucket(void* comm, void* stream, void* buf, int64_t count, int dtype) {
  DNNCA_CHECK_ARG(comm && buf && count > 0 && (dtype == DNNCA_F32 || dtype == DNNCA_BF16), "nccl_allreduce_bucket: bad arguments");
  NcclApi* api = nccl_api();
  if (!api) return DNNCA_ERR_NCCL;
  ncclResult_t r = api->AllReduce(buf, buf, (size_t)count, dtype == DNNCA_F32 ? ncclFloat32 : ncclBfloat16, ncclSum,
                                  reinterpret_cast<ncclComm_t>(comm), (cudaStream_t)stream);
  return r == ncclSuccess ? DNNCA_OK : nccl_fail(api, r, "nccl_allreduce_bucket");
}

extern "C" int dnnca_nccl_broadcast(void* comm, void* stream, void* buf, int64_t bytes, int root) {
  DNNCA_CHECK_ARG(comm && buf && bytes > 0 && root >= 0, "nccl_broadcast: bad arguments");
  NcclApi* api = nccl_api();
  if (!api) return DNNCA_ERR_NCCL;
  ncclResult_t r = api->Broadcast(buf, buf, (size_t)bytes, ncclUint8, root, reinterpret_cast<ncclComm_t>(comm), (cudaStream_t)stream);
  return r == ncclSuccess ? DNNCA_OK : nccl_fail(api, r, "nccl_broadcast");
}
