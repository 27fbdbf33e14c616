// mpc_tasks_sim.cpp — TEST INFRASTRUCTURE: drives the task state machine of crb_mpc_tasks.cu on the CPU.
//
// Compiles cpprobotics_b200/csrc/crb_mpc_core.cuh (the arithmetic and the per-sweep task functions the
// CUDA kernel inlines) with g++ and emulates ONE CTA of the resident-slot kernel: S slots, a problem
// counter, and a scheduler that repeatedly picks a sweep kind and up to 32 waiting slots of that kind — in a
// pseudo-random order instead of the kernel's "kind with the most waiting slots", so that the test also
// shows the result does not depend on the schedule.  tests/test_mpc_tasks_sim.py compares the outputs bit for
// bit with oracle/crb_oracle_mpc.c.  Nothing in libcrb calls this; the product has no CPU path.
//
// Build: g++ -O2 -ffp-contract=off -fno-strict-aliasing [-mfma] -shared -fPIC -DCRB_HOST_SIM ...
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../cpprobotics_b200/csrc/crb_mpc_core.cuh"

struct sim_params {  // == crb_mpc_params (include/crb.h)
  float dt, wb, max_steer, max_accel, max_speed, min_speed;
  float w_a, w_delta, w_da, w_ddelta, w_x, w_y, w_yaw, w_v;
  int max_iter;
  float du_th;
  int max_ls;
  float j_tol;
};

static uint64_t rng_next(uint64_t* s) {
  *s += 0x9E3779B97F4A7C15ull;
  uint64_t z = *s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

extern "C" int mpc_tasks_sim_solve(int64_t n, int T, const float* x0, const float* xref, const float* u_init,
                                   const sim_params* prm, float* sol, float* u0, float* cost, int32_t* status,
                                   int32_t* iters, int S, uint64_t seed, int64_t* task_counts /*[3] or NULL*/,
                                   const int32_t* hint /*[n] or NULL: the kernel's hinted order*/) {
  MpcP p;
  p.dt = prm->dt;
  p.inv_dt = 1.0f / prm->dt;
  p.inv_wb = 1.0f / prm->wb;
  p.max_steer = prm->max_steer; p.max_accel = prm->max_accel;
  p.max_speed = prm->max_speed; p.min_speed = prm->min_speed;
  p.w_a = prm->w_a; p.w_delta = prm->w_delta; p.w_da = prm->w_da; p.w_ddelta = prm->w_ddelta;
  p.wq[0] = prm->w_x; p.wq[1] = prm->w_y; p.wq[2] = prm->w_yaw; p.wq[3] = prm->w_v;
  p.max_iter = prm->max_iter; p.du_th = prm->du_th; p.max_ls = prm->max_ls; p.j_tol = prm->j_tol;

  const int N = T - 1, SW = mpc_slot_words(T), trw = mpc_slot_tr_words(T);
  float* smem = (float*)aligned_alloc(16, (size_t)S * SW * sizeof(float));
  float* slab = (float*)aligned_alloc(16, (size_t)S * N * MPC_REC * sizeof(float));
  // poison: a sweep that reads something no earlier sweep wrote should not get zeros by luck
  for (size_t k = 0; k < (size_t)S * SW; ++k) smem[k] = 1.0e30f;
  for (size_t k = 0; k < (size_t)S * N * MPC_REC; ++k) slab[k] = -1.0e30f;
  std::vector<int> phase(S, MPC_PH_REFILL);
  auto slot_of = [&](int s) {
    MpcSlot sl;
    sl.tr = smem + (size_t)s * SW;
    sl.sw = sl.tr + trw;
    sl.rec = slab + (size_t)s * N * MPC_REC;
    sl.pol = 0;
    sl.ring = 0;
    sl.ring_bulk = 0;
    sl.mbar = 0;
    sl.rounds[0] = sl.rounds[1] = sl.rounds[2] = 0;
    return sl;
  };
  for (int s = 0; s < S; ++s) mpc_sw_int(slot_of(s), MPC_SW_PROB) = -1;
  int64_t next_problem = 0;   // the kernel's global counter
  int64_t counts[3] = {0, 0, 0};
  // hinted order: the problems with hint >= thr in the order of decreasing (clamped) hint, then all the others;
  // inside a bin the GPU's order depends on the run, here it is by index
  std::vector<int64_t> perm;
  if (hint) {
    unsigned hist[MPC_HINT_BINS] = {0};
    for (int64_t i = 0; i < n; ++i) hist[mpc_hint_clamp(hint[i])] += 1u;
    const int thr = mpc_hint_threshold(hist, n);
    perm.reserve((size_t)n);
    for (int b = MPC_HINT_BINS - 1; b >= thr; --b)
      for (int64_t i = 0; i < n; ++i)
        if (mpc_hint_clamp(hint[i]) == b) perm.push_back(i);
    for (int64_t i = 0; i < n; ++i)
      if (mpc_hint_clamp(hint[i]) < thr) perm.push_back(i);
    if ((int64_t)perm.size() != n) return 1;
  }
  uint64_t rs = seed;
  for (;;) {
    // waiting slots by kind
    std::vector<int> w[4];
    for (int s = 0; s < S; ++s)
      if (phase[s] >= MPC_PH_REFILL && phase[s] <= MPC_PH_FW) w[phase[s]].push_back(s);
    int kinds[3], nk = 0;
    for (int k = MPC_PH_REFILL; k <= MPC_PH_FW; ++k)
      if (!w[k].empty()) kinds[nk++] = k;
    if (nk == 0) break;
    const int kind = kinds[rng_next(&rs) % nk];
    std::vector<int>& q = w[kind];
    // random subset of up to 32 (sometimes fewer) in random order
    for (size_t a = q.size(); a > 1; --a) {
      const size_t b = rng_next(&rs) % a;
      const int tmp = q[a - 1]; q[a - 1] = q[b]; q[b] = tmp;
    }
    size_t take = q.size() < 32 ? q.size() : 32;
    if (take > 1 && (rng_next(&rs) & 3) == 0) take = 1 + rng_next(&rs) % take;
    counts[kind - 1] += 1;
    if (kind == MPC_PH_REFILL) {
      for (size_t k = 0; k < take; ++k) {
        MpcSlot sl = slot_of(q[k]);
        if (mpc_sw_int(sl, MPC_SW_PROB) >= 0) mpc_task_retire(sl, T, p, n, sol, u0, cost, status, iters);
      }
      // positions in the start order of the lanes: consecutive; the problem is the position itself or perm[position]
      std::vector<int64_t> pick(take, n);
      for (size_t k = 0; k < take; ++k) {
        const int64_t pos = next_problem + (int64_t)k;
        if (pos < n) pick[k] = hint ? perm[(size_t)pos] : pos;
      }
      next_problem += (int64_t)take;
      for (size_t k = 0; k < take; ++k) {
        MpcSlot sl = slot_of(q[k]);
        const int64_t i = pick[k];
        if (i < n) {
          phase[q[k]] = mpc_task_init(sl, T, p, i, n, x0, xref, u_init);
        } else {
          mpc_sw_int(sl, MPC_SW_PROB) = -1;
          phase[q[k]] = MPC_PH_DEAD;
        }
      }
    } else {
      for (size_t k = 0; k < take; ++k) {
        MpcSlot sl = slot_of(q[k]);
        phase[q[k]] = kind == MPC_PH_BW ? mpc_task_bw(sl, T, p) : mpc_task_fw<false>(sl, T, p);
      }
    }
  }
  if (task_counts) memcpy(task_counts, counts, sizeof(counts));
  free(smem);
  free(slab);
  return 0;
}
