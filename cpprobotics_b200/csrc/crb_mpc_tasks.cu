// crb_mpc_tasks.cu — resident-slot MPC solver for sm_100a: persistent CTAs, roll-outs in shared memory,
// sweeps regrouped by kind.
//
// Same solver, same bits as crb_mpc_solve_kernel (crb_mpc.cu) and oracle/crb_oracle_mpc.c; replaces
// mpc_solve() + FG_EVAL, src/model_predictive_control.cpp:188-346.  What changes is WHERE a problem lives
// and WHO advances it:
//
//   * One persistent CTA per SM.  A CTA owns S problem SLOTS.  Both roll-out buffers of a slot (X, U: the
//     data every sweep reads and writes with a dependent chain behind it) and its scalar state stay in
//     SHARED MEMORY for the whole solve (976 B per slot at T = 20, ~230 slots per SM).  The stage gains and
//     the translated reference (written once per backward sweep / once per problem, read by the next
//     forward sweeps) go to a per-CTA slab of 80-byte stage records in global memory that the CTA reuses for
//     every problem it ever solves: 148 x S x 1.5 KB ~ 50 MB, resident in the 126 MB L2.  The first version
//     of the solver streamed a 2.3 KB per-problem workspace through HBM ~20 times per solve (3.1 GB of DRAM
//     traffic for 53 MB of input + output); here DRAM sees the inputs and the outputs.
//   * A problem is advanced one SWEEP at a time (refill = retire + load + initial roll-out, backward,
//     forward); between sweeps everything it needs is in its slot, so ANY thread can run its next sweep.
//     Each warp repeatedly takes up to 32 slots that wait for the same kind of sweep and runs that sweep
//     with one problem per lane.  An adaptive solver diverges badly under a fixed problem-to-lane map
//     (4..15 outer iterations, 1..5 line-search passes: 16.5 of 32 lanes active in the first version);
//     regrouped, the lanes of a warp always execute the same sweep kind for the same stage count.
//   * Problems are pulled from a global counter 32 at a time, so the 148 CTAs balance themselves and the
//     inputs of a refill are coalesced 128-byte rows of the SoA arrays.
//
// Scheduling state per CTA (shared memory): phase[S] (what each slot waits for), one lock word.  A warp
// takes the lock (~150 cycles, once per ~10^4-cycle sweep), picks the kind with the most waiting slots,
// marks up to 32 of them BUSY, releases.  Results are published with a block-scope fence before the
// slot's new phase is stored.  All waiting loops are bounded: on overrun the kernel raises the error word
// in the slab header instead of hanging the GPU.
#include "crb_common.cuh"
#include "crb_mpc_core.cuh"
#include "crb_mpc_tasks.cuh"

#define MPC_TASK_MAX_CHUNKS 8  // S <= 256 slots

struct MpcTaskArgs {
  int64_t count, ld_in, ld_out;
  int T, S, slot_words;
  const float* x0;
  const float* xref;
  const float* u_init;
  unsigned long long* header;  // [0] next problem index, [1] error word
  float* slab;                 // [grid][S][T-1][MPC_REC]
  float* sol;
  float* u0;
  float* cost;
  int32_t* status;
  int32_t* iters;
};

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__global__ void __launch_bounds__(MPC_TASK_MAX_WARPS * 32, 1)
crb_mpc_tasks_kernel(const __grid_constant__ MpcTaskArgs A, const __grid_constant__ MpcP p) {
  extern __shared__ __align__(16) float smem[];
  const int S = A.S, T = A.T, N = T - 1, SW = A.slot_words;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* phase = reinterpret_cast<int*>(smem + (size_t)S * SW);  // [MPC_TASK_MAX_CHUNKS * 32]
  int* claim = phase + MPC_TASK_MAX_CHUNKS * 32;                // [MPC_TASK_MAX_WARPS][32]
  int* lock = claim + MPC_TASK_MAX_WARPS * 32;
  for (int s = threadIdx.x; s < MPC_TASK_MAX_CHUNKS * 32; s += blockDim.x)
    phase[s] = s < S ? MPC_PH_REFILL : MPC_PH_DEAD;
  for (int s = threadIdx.x; s < S; s += blockDim.x)
    reinterpret_cast<int*>(smem + (size_t)s * SW + mpc_slot_tr_words(T))[MPC_SW_PROB] = -1;
  if (threadIdx.x == 0) *lock = 0;
  __syncthreads();
  float* const slab = A.slab + (size_t)blockIdx.x * S * N * MPC_REC;
  const int tr_words = mpc_slot_tr_words(T);
  const unsigned FULL = 0xffffffffu;
  int idle_spins = 0;

  for (;;) {
    // ---- claim: up to 32 slots waiting for the same kind of sweep ---------------------------------
    int got = 1;
    if (lane == 0) {
      int tries = 0;
      while (atomicCAS(lock, 0, 1) != 0) {
        __nanosleep(64);
        if (++tries > (1 << 22)) { atomicExch(&A.header[1], 1ull); got = 0; break; }
      }
    }
    got = __shfl_sync(FULL, got, 0);
    if (!got) break;  // never hang the GPU on a scheduling bug: raise the error word and leave
    __threadfence_block();
    int ph[MPC_TASK_MAX_CHUNKS];
    int cnt_refill = 0, cnt_bw = 0, cnt_fw = 0, cnt_busy = 0;
#pragma unroll
    for (int c = 0; c < MPC_TASK_MAX_CHUNKS; ++c) {
      ph[c] = *reinterpret_cast<volatile int*>(phase + c * 32 + lane);
      cnt_refill += __popc(__ballot_sync(FULL, ph[c] == MPC_PH_REFILL));
      cnt_bw += __popc(__ballot_sync(FULL, ph[c] == MPC_PH_BW));
      cnt_fw += __popc(__ballot_sync(FULL, ph[c] == MPC_PH_FW));
      cnt_busy += __popc(__ballot_sync(FULL, ph[c] == MPC_PH_BUSY));
    }
    int kind = MPC_PH_FW, best = cnt_fw;
    if (cnt_bw > best) { kind = MPC_PH_BW; best = cnt_bw; }
    if (cnt_refill > best) { kind = MPC_PH_REFILL; best = cnt_refill; }
    int taken = 0;
    if (best > 0) {
#pragma unroll
      for (int c = 0; c < MPC_TASK_MAX_CHUNKS; ++c) {
        const bool mine = ph[c] == kind;
        const unsigned m = __ballot_sync(FULL, mine);
        const int r = taken + __popc(m & lanemask_lt());
        if (mine && r < 32) {
          phase[c * 32 + lane] = MPC_PH_BUSY;
          claim[warp * 32 + r] = c * 32 + lane;
        }
        taken += __popc(m);
      }
      if (taken > 32) taken = 32;
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicExch(lock, 0);
    if (best == 0) {
      if (cnt_busy == 0) break;  // every slot is dead: this CTA is done
      __nanosleep(256);
      if (++idle_spins > (1 << 22)) { if (lane == 0) atomicExch(&A.header[1], 2ull); break; }
      continue;
    }
    idle_spins = 0;
    const bool active = lane < taken;
    const int slot = active ? claim[warp * 32 + lane] : 0;
    __syncwarp();
    MpcSlot sl;
    sl.tr = smem + (size_t)slot * SW;
    sl.sw = sl.tr + tr_words;
    sl.rec = slab + (size_t)slot * N * MPC_REC;

    int next = MPC_PH_DEAD;
    if (kind == MPC_PH_BW) {
      if (active) next = mpc_task_bw(sl, T, p);
    } else if (kind == MPC_PH_FW) {
      if (active) next = mpc_task_fw(sl, T, p);
    } else {
      // retire what the slot holds, then pull the next problems: consecutive indices for consecutive
      // lanes, so the loads of a refill are coalesced rows of x0 / xref
      if (active && mpc_sw_int(sl, MPC_SW_PROB) >= 0)
        mpc_task_retire(sl, T, p, A.ld_out, A.sol, A.u0, A.cost, A.status, A.iters);
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&A.header[0], (unsigned long long)taken);
      base = __shfl_sync(FULL, base, 0);
      const int64_t i = (int64_t)base + lane;
      if (active) {
        if (i < A.count) {
          next = mpc_task_init(sl, T, p, i, A.ld_in, A.x0, A.xref, A.u_init);
        } else {
          mpc_sw_int(sl, MPC_SW_PROB) = -1;
          next = MPC_PH_DEAD;
        }
      }
    }
    // publish: data first, then the phase
    __threadfence_block();
    if (active) *reinterpret_cast<volatile int*>(phase + slot) = next;
  }
}

// Launch geometry for `count` problems of horizon T: warps per CTA, slots per CTA, CTAs, shared memory.
struct MpcTaskGeom {
  int nwarps, S, grid;
  size_t smem;
};
static bool mpc_tasks_geometry(int sm_count, int T, int64_t count, MpcTaskGeom* g) {
  const size_t smem_cap = 227 * 1024;
  static int env_warps = -1, env_slots = -1;  // A/B knobs (process-wide, read once)
  if (env_warps < 0) {
    const char* e = getenv("CRB_MPC_WARPS");
    env_warps = e ? atoi(e) : 0;
    const char* f = getenv("CRB_MPC_SLOTS");
    env_slots = f ? atoi(f) : 0;
  }
  int nwarps = env_warps > 0 ? env_warps : MPC_TASK_MAX_WARPS;
  if (nwarps > MPC_TASK_MAX_WARPS) nwarps = MPC_TASK_MAX_WARPS;
  const size_t fixed = (size_t)(MPC_TASK_MAX_CHUNKS * 32 + MPC_TASK_MAX_WARPS * 32 + 4) * sizeof(int);
  size_t s = (smem_cap - fixed) / ((size_t)mpc_slot_words(T) * sizeof(float));
  if (s > MPC_TASK_MAX_CHUNKS * 32) s = MPC_TASK_MAX_CHUNKS * 32;
  int S = (int)s;
  if (env_slots > 0 && env_slots < S) S = env_slots;
  if (S < 32) return false;
  // as many warps as leave ~40 slots waiting, so that full warps of one kind can form
  while (nwarps > 1 && nwarps * 32 + 40 > S) --nwarps;
  // one CTA per SM; fewer when the batch has less than one task per warp
  int64_t grid = (count + (int64_t)nwarps * 32 - 1) / ((int64_t)nwarps * 32);
  if (grid > sm_count) grid = sm_count;
  if (grid < 1) grid = 1;
  g->nwarps = nwarps;
  g->S = S;
  g->grid = (int)grid;
  g->smem = (size_t)S * mpc_slot_words(T) * sizeof(float) + fixed;
  return true;
}

// header (256 B) + slab of stage records + 256 B of alignment slack
size_t crb_mpc_tasks_scratch_bytes(int sm_count, int T, int64_t count) {
  MpcTaskGeom g;
  if (!mpc_tasks_geometry(sm_count, T, count, &g)) return 512;
  return 512 + (size_t)g.grid * g.S * (size_t)(T - 1) * MPC_REC * sizeof(float);
}

int crb_mpc_tasks_launch(crb_ctx* ctx, cudaStream_t st, int64_t count, int64_t ld, int T,
                         const float* x0, const float* xref, const float* u_init, void* scratch,
                         int64_t ld_out, float* sol, float* u0, float* cost, int32_t* status,
                         int32_t* iters, const MpcP& p) {
  CRB_REQUIRE(count < ((int64_t)1 << 31), "more than 2^31 problems in one launch");
  MpcTaskGeom g;
  if (!mpc_tasks_geometry(ctx->sm_count, T, count, &g)) {
    crb_set_error("crb_mpc_tasks_launch: T = %d does not fit the resident-slot kernel", T);
    return CRB_ERR_UNSUPPORTED;
  }
  if (!ctx->mpc_tasks_attr_set) {
    CRB_CUDA(cudaFuncSetAttribute(crb_mpc_tasks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
    ctx->mpc_tasks_attr_set = 1;
  }
  char* base = (char*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  MpcTaskArgs a;
  a.count = count; a.ld_in = ld; a.ld_out = ld_out;
  a.T = T; a.S = g.S; a.slot_words = mpc_slot_words(T);
  a.x0 = x0; a.xref = xref; a.u_init = u_init;
  a.header = (unsigned long long*)base;
  a.slab = (float*)(base + 256);
  a.sol = sol; a.u0 = u0; a.cost = cost; a.status = status; a.iters = iters;
  CRB_CUDA(cudaMemsetAsync(base, 0, 256, st));
  crb_mpc_tasks_kernel<<<(unsigned)g.grid, g.nwarps * 32, g.smem, st>>>(a, p);
  CRB_CUDA(cudaGetLastError());
  ctx->launches++;
  return CRB_OK;
}
