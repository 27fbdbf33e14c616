// crb_comm.cu — the multi-GPU side of libcrb: an NCCL communicator owned by the context.
//
// the batch shards over GPUs by contiguous agent ranges and the only traffic that ever
// crosses NVLink is (1) one all-gather of CRB_STATS_LEN doubles per rank per call (crb_gather_stats) and,
// for the particle filter, (2) one all-reduce of the partial weight sums / moments (src/particle_filter.cpp:104
// `pw / pw.sum()` across shards, crb_pf_estimate with a communicator).  Both are latency-bound (64-208
// bytes), so NCCL's own kernels are used as they are: there is no compute to fuse them with.
//
// A C++ host drives the multi-GPU path with nothing but this ABI: one crb_ctx per device (one thread or one
// process each), crb_comm_get_unique_id on rank 0, the 128 bytes handed to every rank by whatever channel
// the host has (a shared variable between threads, MPI, a socket, torch.distributed's store), then
// crb_comm_init_rank everywhere.  tests/cpp/multi_gpu_demo.cpp does exactly that with threads.
//
// NCCL is opened at run time (dlopen), so libcrb.so has no link-time dependency on it and loads on boxes
// without NCCL; the first crb_comm_* call reports CRB_ERR_UNSUPPORTED if no library can be found.  Search
// order: $CRB_NCCL_LIB, an already loaded / default-path "libnccl.so.2" (inside a PyTorch process that is
// torch's bundled copy, so both sides speak the same NCCL), "libnccl.so".
#include <dlfcn.h>

#include "crb_common.cuh"

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { CRB_NCCL_SUCCESS = 0 };
enum { CRB_NCCL_FLOAT64 = 8 };  // ncclDouble
enum { CRB_NCCL_SUM = 0 };      // ncclSum

struct NcclApi {
  void* handle;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*GetVersion)(int*);
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static int state = 0;  // 0 untried, 1 ok, -1 unavailable
  if (state == 0) {
    const char* names[3] = {getenv("CRB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (int k = 0; k < 3 && !h; ++k)
      if (names[k] && names[k][0]) h = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.handle = h;
      api.GetUniqueId = (ncclResult_t(*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (ncclResult_t(*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (ncclResult_t(*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
      api.AllGather = (ncclResult_t(*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
      api.AllReduce = (ncclResult_t(*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
      api.GetErrorString = (const char* (*)(ncclResult_t))dlsym(h, "ncclGetErrorString");
      api.GetVersion = (ncclResult_t(*)(int*))dlsym(h, "ncclGetVersion");
      state = (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce) ? 1 : -1;
    } else {
      state = -1;
    }
  }
  return state == 1 ? &api : nullptr;
}

#define CRB_NCCL(api, call)                                                                       \
  do {                                                                                            \
    ncclResult_t r_ = (call);                                                                     \
    if (r_ != CRB_NCCL_SUCCESS) {                                                                 \
      crb_set_error("%s:%d %s -> NCCL error %d (%s)", __FILE__, __LINE__, #call, (int)r_,         \
                    (api)->GetErrorString ? (api)->GetErrorString(r_) : "?");                     \
      return CRB_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

static NcclApi* need_nccl() {
  NcclApi* a = nccl_api();
  if (!a) crb_set_error("NCCL library not found (tried $CRB_NCCL_LIB, libnccl.so.2, libnccl.so)");
  return a;
}

extern "C" {

int crb_comm_nccl_version(void) {
  NcclApi* a = nccl_api();
  int v = 0;
  if (!a || !a->GetVersion || a->GetVersion(&v) != CRB_NCCL_SUCCESS) return 0;
  return v;
}

int crb_comm_get_unique_id(void* id_out) {
  CRB_REQUIRE(id_out != nullptr, "id_out is NULL");
  NcclApi* a = need_nccl();
  if (!a) return CRB_ERR_UNSUPPORTED;
  ncclUniqueId id;
  CRB_NCCL(a, a->GetUniqueId(&id));
  memcpy(id_out, &id, CRB_COMM_ID_BYTES);
  return CRB_OK;
}

int crb_comm_init_rank(crb_ctx* ctx, int world, int rank, const void* id) {
  CRB_REQUIRE(ctx != nullptr && id != nullptr, "ctx / id is NULL");
  CRB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank out of range");
  CRB_REQUIRE(ctx->comm == nullptr, "this context already has a communicator");
  NcclApi* a = need_nccl();
  if (!a) return CRB_ERR_UNSUPPORTED;
  CRB_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  memcpy(&uid, id, CRB_COMM_ID_BYTES);
  ncclComm_t comm = nullptr;
  CRB_NCCL(a, a->CommInitRank(&comm, world, uid, rank));
  ctx->comm = comm;
  ctx->comm_world = world;
  ctx->comm_rank = rank;
  return CRB_OK;
}

int crb_comm_destroy(crb_ctx* ctx) {
  CRB_REQUIRE(ctx != nullptr, "ctx is NULL");
  if (!ctx->comm) return CRB_OK;
  NcclApi* a = need_nccl();
  if (!a) return CRB_ERR_UNSUPPORTED;
  cudaStreamSynchronize(ctx->stream);
  CRB_NCCL(a, a->CommDestroy((ncclComm_t)ctx->comm));
  ctx->comm = nullptr;
  ctx->comm_world = 1;
  ctx->comm_rank = 0;
  return CRB_OK;
}

int crb_comm_world(crb_ctx* ctx) { return ctx && ctx->comm ? ctx->comm_world : 1; }
int crb_comm_rank(crb_ctx* ctx) { return ctx && ctx->comm ? ctx->comm_rank : 0; }

int crb_gather_stats(crb_ctx* ctx, const double* stats_dev, double* all_dev) {
  CRB_REQUIRE(ctx != nullptr && stats_dev != nullptr && all_dev != nullptr, "NULL argument");
  if (!ctx->comm) {  // single GPU: the "gather" of one rank
    if (all_dev != stats_dev)
      CRB_CUDA(cudaMemcpyAsync(all_dev, stats_dev, CRB_STATS_LEN * sizeof(double), cudaMemcpyDeviceToDevice,
                               ctx->stream));
    return CRB_OK;
  }
  NcclApi* a = need_nccl();
  if (!a) return CRB_ERR_UNSUPPORTED;
  CRB_NCCL(a, a->AllGather(stats_dev, all_dev, CRB_STATS_LEN, CRB_NCCL_FLOAT64, (ncclComm_t)ctx->comm, ctx->stream));
  return CRB_OK;
}

int crb_comm_allreduce_sum_f64(crb_ctx* ctx, double* buf_dev, int64_t count) {
  CRB_REQUIRE(ctx != nullptr && buf_dev != nullptr && count >= 0, "NULL argument or count < 0");
  if (!ctx->comm || count == 0) return CRB_OK;
  NcclApi* a = need_nccl();
  if (!a) return CRB_ERR_UNSUPPORTED;
  CRB_NCCL(a, a->AllReduce(buf_dev, buf_dev, (size_t)count, CRB_NCCL_FLOAT64, CRB_NCCL_SUM, (ncclComm_t)ctx->comm,
                           ctx->stream));
  return CRB_OK;
}

}  // extern "C"
