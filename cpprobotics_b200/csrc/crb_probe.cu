// crb_probe.cu — measurement aids: the roofline denominators that MEASURED_PEAKS.json does not carry.
//
// crb_probe_fp32_peak: the non-tensor binary32 FMA rate of THIS device at ITS current clocks, measured by a
// kernel that does nothing else (8 independent FFMA chains per thread, 32 resident warps per SM, operands in
// registers).  It is the denominator of the MPC / LQR roofline fractions in bench.py ("fp32 peak, measured in
// the same run"), instead of the nominal SMs x 128 lanes x 2 x clock.
#include "crb_common.cuh"

__global__ void __launch_bounds__(256, 4) crb_probe_ffma_kernel(float* out, int iters, float a, float b) {
  float v0 = threadIdx.x * 1.0e-3f, v1 = v0 + 1.0f, v2 = v0 + 2.0f, v3 = v0 + 3.0f;
  float v4 = v0 + 4.0f, v5 = v0 + 5.0f, v6 = v0 + 6.0f, v7 = v0 + 7.0f;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      v0 = fmaf(v0, a, b); v1 = fmaf(v1, a, b); v2 = fmaf(v2, a, b); v3 = fmaf(v3, a, b);
      v4 = fmaf(v4, a, b); v5 = fmaf(v5, a, b); v6 = fmaf(v6, a, b); v7 = fmaf(v7, a, b);
    }
  }
  const float s = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the chains alive
}

extern "C" int crb_probe_fp32_peak(crb_ctx* ctx, double* tflops_out) {
  CRB_REQUIRE(ctx != nullptr && tflops_out != nullptr, "NULL argument");
  CRB_DEVICE_GUARD(ctx);
  int rc = crb_ctx_scratch_reserve(ctx, (size_t)ctx->sm_count * 8 * 256 * sizeof(float));
  if (rc) return rc;
  const int grid = ctx->sm_count * 8, iters = 4096;   // ~0.3 ms per launch
  float* out = (float*)ctx->scratch;
  cudaStream_t st = ctx->stream;
  for (int w = 0; w < 2; ++w) crb_probe_ffma_kernel<<<grid, 256, 0, st>>>(out, iters, 0.999f, 1.0e-3f);
  double best = 0.0;
  for (int r = 0; r < 5; ++r) {
    CRB_CUDA(cudaEventRecord(ctx->ev_start, st));
    crb_probe_ffma_kernel<<<grid, 256, 0, st>>>(out, iters, 0.999f, 1.0e-3f);
    CRB_CUDA(cudaEventRecord(ctx->ev_stop, st));
    CRB_CUDA(cudaEventSynchronize(ctx->ev_stop));
    float ms = 0.0f;
    CRB_CUDA(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)grid * 256.0;
    const double tf = flops / ((double)ms * 1.0e-3) / 1.0e12;
    if (tf > best) best = tf;
  }
  CRB_CUDA(cudaGetLastError());
  ctx->launches += 7;
  *tflops_out = best;
  return CRB_OK;
}
