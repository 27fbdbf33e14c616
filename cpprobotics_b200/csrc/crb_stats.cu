// crb_stats.cu — per-GPU summary statistics: the only payload that ever crosses NVLink.
//
// No reference counterpart (the reference is single-robot).  Each rank reduces its shard's per-agent
// results to CRB_STATS_LEN doubles on the device; the caller all-gathers 64 bytes per rank (NCCL).
// Deterministic: fixed block count, fixed slice per block, partials combined in index order.
#include <math.h>

#include "crb_common.cuh"

#define ST_BLOCKS 512
#define ST_THREADS 256

__global__ void __launch_bounds__(ST_THREADS)
crb_stats_partial_kernel(int64_t n, int64_t i0, const float* __restrict__ values,
                         const int32_t* __restrict__ status, const int32_t* __restrict__ iters,
                         double* __restrict__ partial /*[blocks][8]*/) {
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t b0 = (int64_t)blockIdx.x * per;
  const int64_t b1 = b0 + per < n ? b0 + per : n;
  double sum = 0.0, mn = INFINITY, mx = -INFINITY, nonfinite = 0.0, conv = 0.0, its = 0.0,
         chk = 0.0;
  for (int64_t i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
    const float v = values[i];
    if (isfinite(v)) {
      sum += (double)v;
      mn = fmin(mn, (double)v);
      mx = fmax(mx, (double)v);
      chk += (double)v * (double)((i0 + i) % 251 + 1);
    } else {
      nonfinite += 1.0;
    }
    if (status) conv += status[i] == CRB_MPC_CONVERGED ? 1.0 : 0.0;
    if (iters) its += (double)iters[i];
  }
  __shared__ double sm[7][ST_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double v7[7] = {sum, mn, mx, nonfinite, conv, its, chk};
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    double t = v7[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_down_sync(0xffffffffu, t, o);
      t = k == 1 ? fmin(t, other) : (k == 2 ? fmax(t, other) : t + other);
    }
    if (lane == 0) sm[k][wid] = t;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    const int k = threadIdx.x;
    double t = sm[k][0];
    for (int w = 1; w < ST_THREADS / 32; ++w)
      t = k == 1 ? fmin(t, sm[k][w]) : (k == 2 ? fmax(t, sm[k][w]) : t + sm[k][w]);
    partial[(size_t)blockIdx.x * 8 + k] = t;
  }
}

// One CTA of ST_BLOCKS threads: thread b takes block b's partial, then a fixed shuffle/shared-memory
// tree (same shape every run -> bit-reproducible).
__global__ void __launch_bounds__(ST_BLOCKS)
crb_stats_combine_kernel(int64_t n, const double* __restrict__ partial, double* __restrict__ out) {
  __shared__ double sm[7][ST_BLOCKS / 32];
  const int b = threadIdx.x, lane = b & 31, wid = b >> 5;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    double t = partial[(size_t)b * 8 + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_down_sync(0xffffffffu, t, o);
      t = k == 1 ? fmin(t, other) : (k == 2 ? fmax(t, other) : t + other);
    }
    if (lane == 0) sm[k][wid] = t;
  }
  __syncthreads();
  if (b < 7) {
    double t = sm[b][0];
    for (int w = 1; w < ST_BLOCKS / 32; ++w)
      t = b == 1 ? fmin(t, sm[b][w]) : (b == 2 ? fmax(t, sm[b][w]) : t + sm[b][w]);
    out[b] = t;
  }
  if (b == 7) out[7] = (double)n;
}

extern "C" int crb_stats_reduce(crb_ctx* ctx, int64_t n, int64_t i0, const float* values,
                                const int32_t* status, const int32_t* iters, double* stats_dev) {
  CRB_REQUIRE(ctx != nullptr, "ctx is NULL");
  CRB_DEVICE_GUARD(ctx);   // ctx->device is current for this call, the caller's device is restored after it
  CRB_REQUIRE(n > 0 && values && stats_dev, "n <= 0 or NULL array");
  int rc = crb_ctx_scratch_reserve(ctx, (size_t)ST_BLOCKS * 8 * sizeof(double));
  if (rc) return rc;
  double* partial = (double*)ctx->scratch;
  crb_stats_partial_kernel<<<ST_BLOCKS, ST_THREADS, 0, ctx->stream>>>(n, i0, values, status, iters,
                                                                      partial);
  crb_stats_combine_kernel<<<1, ST_BLOCKS, 0, ctx->stream>>>(n, partial, stats_dev);
  CRB_CUDA(cudaGetLastError());
  ctx->launches += 2;
  return CRB_OK;
}
