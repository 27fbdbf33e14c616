// crb_common.cuh — context object, error plumbing and launch helpers shared by the libcrb TUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/crb.h"

#define CRB_N_PIPE 3  // host-pointer entry points pipeline chunks over this many streams

struct crb_ctx {
  int device;
  int sm_count;
  cudaStream_t own_stream;
  cudaStream_t stream;  // the stream kernels are enqueued on (own_stream or caller's)
  int64_t launches;
  cudaEvent_t ev_start, ev_stop;
  // staging for the *_host entry points: one arena per pipeline slot, grown on demand
  cudaStream_t pipe_stream[CRB_N_PIPE];
  void* pipe_buf[CRB_N_PIPE];
  size_t pipe_cap[CRB_N_PIPE];
  // scratch for reductions / solver workspaces, grown on demand
  void* scratch;
  size_t scratch_cap;
  // MPC solver workspace (two trajectory buffers + gains per problem) for the device-pointer entry
  void* mpc_ws;
  size_t mpc_ws_cap;
  void* host_scratch;  // small pinned buffer for reduction results
};

void crb_set_error(const char* fmt, ...);

#define CRB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      crb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
      return e_ == cudaErrorMemoryAllocation ? CRB_ERR_ALLOC : CRB_ERR_CUDA;                \
    }                                                                                       \
  } while (0)

#define CRB_REQUIRE(cond, msg)                                  \
  do {                                                          \
    if (!(cond)) {                                              \
      crb_set_error("%s: invalid argument: %s", __func__, msg); \
      return CRB_ERR_INVALID_ARG;                               \
    }                                                           \
  } while (0)

// Grow-only arenas.
int crb_ctx_pipe_reserve(crb_ctx* ctx, int slot, size_t bytes);
int crb_ctx_scratch_reserve(crb_ctx* ctx, size_t bytes);
int crb_ctx_mpc_ws_reserve(crb_ctx* ctx, size_t bytes);

// Strided host<->device block copy for the *_host pipelines.  Measured on this platform
// (scripts/pcie_probe.py + bench e2e): plain 1-D copies reach 48 (H2D) / 57 (D2H) GB/s, one
// cudaMemcpy2DAsync per array ~31 GB/s per direction in duplex, and splitting it into per-row 1-D copies is
// SLOWER (enqueue-bound: 2.0-2.8e8 vs 3.45e8 EKF updates/s end to end), so the 2-D copy stays.
static inline cudaError_t crb_copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch,
                                        size_t width, size_t rows, cudaMemcpyKind kind, cudaStream_t st) {
  return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st);
}

static inline int crb_grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// Streaming (evict-first) global accesses for data that is touched exactly once per launch.
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
