// crb_mpc_tasks.cuh — launcher of the resident-slot MPC kernel (crb_mpc_tasks.cu), used by crb_mpc.cu.
#pragma once
#include "crb_common.cuh"

struct MpcP;
#define MPC_TASK_MAX_WARPS 6

// bytes of device scratch (header + per-CTA slab of stage records) for `count` problems of horizon T on a
// device with sm_count SMs
size_t crb_mpc_tasks_scratch_bytes(int sm_count, int T, int64_t count);
// inputs x0 / xref / u_init have leading dimension ld (>= count); outputs ld_out.  `scratch` must hold
// crb_mpc_tasks_scratch_bytes() bytes (any alignment); it starts with the launch header (problem counter, error word,
// histogram of the hints, bin cursors).  hint (device, [count], may be NULL): expected work per problem, larger =
// started earlier (counting sort by decreasing hint, crb_mpc_hint_perm_kernel); it changes the order in which
// problems start, nothing else.
int crb_mpc_tasks_launch(crb_ctx* ctx, cudaStream_t st, int64_t count, int64_t ld, int T,
                         const float* x0, const float* xref, const float* u_init, void* scratch,
                         int64_t ld_out, float* sol, float* u0, float* cost, int32_t* status,
                         int32_t* iters, const MpcP& p, const int32_t* hint);
