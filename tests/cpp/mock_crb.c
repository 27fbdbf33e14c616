/* mock_crb.c - TEST-ONLY stand-in for libcrb.so: the handful of C-ABI entry points the C++ shims
 * resampling<NP>(), solve_DARE() and dlqr() use, implemented on the CPU with the oracle (liboracle.so).
 * It lets tests/test_reference_api.py execute those shims on a machine without a GPU and check their
 * marshalling (AoS <-> SoA, column-major matrices, by-value RNG) against direct oracle calls.
 * Never shipped, never linked into the product; "device" memory is malloc. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/crb.h"
#include "../../oracle/crb_oracle.h"

struct crb_ctx { int unused; };
static struct crb_ctx g_ctx;

const char* crb_last_error_string(void) { return "mock_crb: no error text"; }
int crb_init(crb_ctx** out, int device) { (void)device; *out = &g_ctx; return CRB_OK; }
int crb_destroy(crb_ctx* ctx) { (void)ctx; return CRB_OK; }
int crb_device_alloc(crb_ctx* ctx, void** out, size_t bytes) { (void)ctx; *out = malloc(bytes ? bytes : 1); return *out ? CRB_OK : -1; }
int crb_device_free(crb_ctx* ctx, void* p) { (void)ctx; free(p); return CRB_OK; }
int crb_memcpy_h2d(crb_ctx* ctx, void* d, const void* s, size_t n) { (void)ctx; memcpy(d, s, n); return CRB_OK; }
int crb_memcpy_d2h(crb_ctx* ctx, void* d, const void* s, size_t n) { (void)ctx; memcpy(d, s, n); return CRB_OK; }

int crb_pf_resample(crb_ctx* ctx, int64_t n, float* px, float* pw, float* px_tmp, const float* uniforms,
                    uint64_t seed, float nth, int* did_resample_host, double* neff_host) {
  (void)ctx; (void)px_tmp; (void)seed;
  double* u = (double*)malloc((size_t)n * sizeof(double));
  for (int64_t i = 0; i < n; ++i) u[i] = (double)uniforms[i];
  float neff = 0.0f;
  const int did = crb_oracle_pf_resample(n, px, pw, u, nth, /*reference_mode*/ 0, &neff);
  free(u);
  if (did_resample_host) *did_resample_host = did;
  if (neff_host) *neff_host = (double)neff;
  return CRB_OK;
}

int crb_lqr_dlqr_batched(crb_ctx* ctx, int64_t n, int nx, int nu, const float* A, const float* B,
                         const float* Q, const float* R, int maxiter, float eps, float* K, float* X,
                         int32_t* iters) {
  (void)ctx;
  float xtmp[25];
  int32_t it = 0;
  if (n != 1) return -1;
  crb_oracle_dlqr_batched(1, nx, nu, A, B, Q, R, maxiter, eps, K, X ? X : xtmp, iters ? iters : &it, 1);
  return CRB_OK;
}

/* the other entry points the header's shims reference: present so that the link succeeds, never called here */
int crb_ekf_step_batched_host(crb_ctx* c, int64_t n, float* x, float* P, const float* z, const float* u,
                              const crb_ekf_params* p, int s) { (void)c; (void)n; (void)x; (void)P; (void)z; (void)u; (void)p; (void)s; return -1; }
void crb_ekf_default_params(crb_ekf_params* p) { memset(p, 0, sizeof(*p)); }
void crb_pf_default_params(crb_pf_params* p) { memset(p, 0, sizeof(*p)); }
void crb_mpc_default_params(crb_mpc_params* p) { memset(p, 0, sizeof(*p)); }
int crb_pf_predict_weight_batched(crb_ctx* c, int64_t n, float* px, float* pw, const float* noise, uint64_t seed,
                                  const float* lm, int n_lm, const crb_pf_params* p) { (void)c; (void)n; (void)px; (void)pw; (void)noise; (void)seed; (void)lm; (void)n_lm; (void)p; return -1; }
int crb_pf_estimate(crb_ctx* c, int64_t n, const float* px, float* pw, float* xe, float* pe, double* sw) { (void)c; (void)n; (void)px; (void)pw; (void)xe; (void)pe; (void)sw; return -1; }
int crb_mpc_solve_batched_host(crb_ctx* c, int64_t n, int T, const float* x0, const float* xref, const float* ui,
                               const crb_mpc_params* p, float* sol, float* u0, float* cost, int32_t* st,
                               int32_t* it) { (void)c; (void)n; (void)T; (void)x0; (void)xref; (void)ui; (void)p; (void)sol; (void)u0; (void)cost; (void)st; (void)it; return -1; }
int crb_mpc_plant_update_batched(crb_ctx* c, int64_t n, float* s, const float* u0, const crb_mpc_params* p) { (void)c; (void)n; (void)s; (void)u0; (void)p; return -1; }
int crb_mpc_calc_ref_trajectory_batched(crb_ctx* c, int64_t n, int T, const float* s, const float* cx, const float* cy,
                                        const float* cyaw, const float* sp, int32_t nc, float dl, int32_t* ti,
                                        float* xr, const crb_mpc_params* p) { (void)c; (void)n; (void)T; (void)s; (void)cx; (void)cy; (void)cyaw; (void)sp; (void)nc; (void)dl; (void)ti; (void)xr; (void)p; return -1; }
