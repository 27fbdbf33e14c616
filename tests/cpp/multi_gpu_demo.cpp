// multi_gpu_demo.cpp — a C++ host drives the multi-GPU path with nothing but include/crb.h.
//
// BASELINE.json configs[4] in miniature (the caller loop of src/model_predictive_control.cpp:372-378 for a
// sharded batch): one host thread per GPU, one crb_ctx each, a contiguous shard of the agents per GPU, one
// MPC solve per shard, the per-shard cost statistics reduced on the device and ALL-GATHERED through libcrb's
// own NCCL communicator (crb_comm_* / crb_gather_stats).  No torch, no Python, no NCCL header.
//
//   g++ -std=c++11 -O1 -pthread -I include tests/cpp/multi_gpu_demo.cpp -L cpprobotics_b200/lib -lcrb -o demo
//   ./demo <n_gpus> <agents_per_gpu>        exit code 0 and a line "multi_gpu_demo OK ..." on success
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "crb.h"

static const int T = 20;
static std::mutex g_print;

#define CHECK(call)                                                                              \
  do {                                                                                           \
    int rc_ = (call);                                                                            \
    if (rc_ != CRB_OK) {                                                                         \
      std::lock_guard<std::mutex> lk(g_print);                                                   \
      std::fprintf(stderr, "rank %d: %s failed (%d): %s\n", rank, #call, rc_, crb_last_error_string()); \
      *failed = 1;                                                                               \
      return;                                                                                    \
    }                                                                                            \
  } while (0)

// index-addressed inputs: agent g (global index) near a sinusoidal course, like the bench's generator
static void make_inputs(int64_t i0, int64_t n, std::vector<float>& x0, std::vector<float>& xref) {
  x0.assign(4 * n, 0.0f);
  xref.assign((size_t)4 * T * n, 0.0f);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t g = i0 + i;
    const float s0 = (float)(g % 4000) * 0.1f, v = 1.0f + (float)(g % 7) * 0.5f;
    const float yaw = std::atan2(std::cos(s0 / 20.0f), 1.0f);
    x0[0 * n + i] = s0;
    x0[1 * n + i] = 20.0f * std::sin(s0 / 20.0f) + 0.3f * (float)((g % 5) - 2);
    x0[2 * n + i] = yaw + 0.05f * (float)((g % 3) - 1);
    x0[3 * n + i] = v;
    for (int t = 0; t < T; ++t) {
      const float s = s0 + v * 0.2f * (float)t;
      xref[((size_t)4 * t + 0) * n + i] = s;
      xref[((size_t)4 * t + 1) * n + i] = 20.0f * std::sin(s / 20.0f);
      xref[((size_t)4 * t + 2) * n + i] = std::atan2(std::cos(s / 20.0f), 1.0f);
      xref[((size_t)4 * t + 3) * n + i] = 10.0f / 3.6f;
    }
  }
}

static void worker(int rank, int world, int64_t n_per, const char* uid, std::vector<double>* table, int* failed) {
  crb_ctx* ctx = nullptr;
  CHECK(crb_init(&ctx, rank));
  CHECK(crb_comm_init_rank(ctx, world, rank, uid));
  std::vector<float> x0, xref;
  make_inputs((int64_t)rank * n_per, n_per, x0, xref);
  void *dx0, *dxr, *dcost, *dstat, *dit, *du0, *dstats, *dall;
  CHECK(crb_device_alloc(ctx, &dx0, x0.size() * 4));
  CHECK(crb_device_alloc(ctx, &dxr, xref.size() * 4));
  CHECK(crb_device_alloc(ctx, &dcost, n_per * 4));
  CHECK(crb_device_alloc(ctx, &dstat, n_per * 4));
  CHECK(crb_device_alloc(ctx, &dit, n_per * 4));
  CHECK(crb_device_alloc(ctx, &du0, 2 * n_per * 4));
  CHECK(crb_device_alloc(ctx, &dstats, CRB_STATS_LEN * 8));
  CHECK(crb_device_alloc(ctx, &dall, (size_t)world * CRB_STATS_LEN * 8));
  CHECK(crb_memcpy_h2d(ctx, dx0, x0.data(), x0.size() * 4));
  CHECK(crb_memcpy_h2d(ctx, dxr, xref.data(), xref.size() * 4));
  crb_mpc_params prm;
  crb_mpc_default_params(&prm);
  CHECK(crb_mpc_solve_batched(ctx, n_per, T, (const float*)dx0, (const float*)dxr, nullptr, &prm, nullptr,
                              (float*)du0, (float*)dcost, (int32_t*)dstat, (int32_t*)dit));
  CHECK(crb_stats_reduce(ctx, n_per, (int64_t)rank * n_per, (const float*)dcost, (const int32_t*)dstat,
                         (const int32_t*)dit, (double*)dstats));
  CHECK(crb_gather_stats(ctx, (const double*)dstats, (double*)dall));   // the one collective of the data path
  table->assign((size_t)world * CRB_STATS_LEN, 0.0);
  CHECK(crb_memcpy_d2h(ctx, table->data(), dall, table->size() * 8));
  crb_device_free(ctx, dx0); crb_device_free(ctx, dxr); crb_device_free(ctx, dcost); crb_device_free(ctx, dstat);
  crb_device_free(ctx, dit); crb_device_free(ctx, du0); crb_device_free(ctx, dstats); crb_device_free(ctx, dall);
  CHECK(crb_destroy(ctx));
}

int main(int argc, char** argv) {
  const int world = argc > 1 ? std::atoi(argv[1]) : 2;
  const int64_t n_per = argc > 2 ? std::atoll(argv[2]) : 8192;
  char uid[CRB_COMM_ID_BYTES];
  if (crb_comm_get_unique_id(uid) != CRB_OK) {
    std::fprintf(stderr, "crb_comm_get_unique_id: %s\n", crb_last_error_string());
    return 3;
  }
  std::vector<std::vector<double> > tables(world);
  std::vector<int> failed(world, 0);
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r) th.emplace_back(worker, r, world, n_per, uid, &tables[r], &failed[r]);
  for (auto& t : th) t.join();
  for (int r = 0; r < world; ++r)
    if (failed[r]) return 3;
  // every rank holds the same table; row r describes shard r
  double n_sum = 0, conv = 0, cost = 0;
  for (int r = 0; r < world; ++r) {
    if (std::memcmp(tables[r].data(), tables[0].data(), tables[0].size() * 8) != 0) {
      std::fprintf(stderr, "rank %d gathered a different table\n", r);
      return 4;
    }
    n_sum += tables[0][r * CRB_STATS_LEN + 7];
    conv += tables[0][r * CRB_STATS_LEN + 4];
    cost += tables[0][r * CRB_STATS_LEN + 0];
  }
  if (n_sum != (double)world * (double)n_per || conv < 0.99 * n_sum || !(cost > 0.0)) {
    std::fprintf(stderr, "bad statistics: n %.0f converged %.0f cost %.6g\n", n_sum, conv, cost);
    return 5;
  }
  std::printf("multi_gpu_demo OK world %d agents %.0f converged %.0f mean_cost %.6f nccl %d\n", world, n_sum, conv,
              cost / n_sum, crb_comm_nccl_version());
  return 0;
}
