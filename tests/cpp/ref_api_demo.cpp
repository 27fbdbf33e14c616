// Exercises include/crb/reference_api.hpp the way a CppRobotics main() would: the reference's function
// names, argument order and value semantics.  Inputs come from a raw float32 file written by the
// Python test, outputs go to another one; tests/test_reference_api.py compares them with the oracle.
// Exit code 3 + message when no GPU is usable (there is no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "crb/reference_api.hpp"

using namespace cpprobotics;
static std::vector<float> in;
static size_t pos = 0;
static float rd() { return in.at(pos++); }
static std::vector<float> out;

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  {
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    in.resize(n / 4);
    if (std::fread(in.data(), 4, in.size(), f) != in.size()) return 2;
    std::fclose(f);
  }
  try {
    // ---- EKF, exactly like main() of src/extended_kalman_filter.cpp:131-151,183 ----------------
    const int K = (int)rd();
    crb::Vector4f xEst = crb::Vector4f::Zero();
    crb::Matrix4f PEst = crb::Matrix4f::Identity();
    crb::Matrix4f Q = crb::Matrix4f::Identity();
    Q(0, 0) = 0.1 * 0.1;
    Q(1, 1) = 0.1 * 0.1;
    Q(2, 2) = (1.0 / 180 * M_PI) * (1.0 / 180 * M_PI);
    Q(3, 3) = 0.1 * 0.1;
    crb::Matrix2f R = crb::Matrix2f::Identity();
    for (int k = 0; k < K; ++k) {
      crb::Vector2f z, ud;
      z(0) = rd(); z(1) = rd(); ud(0) = rd(); ud(1) = rd();
      ekf_estimation(xEst, PEst, z, ud, Q, R);
    }
    for (int i = 0; i < 4; ++i) out.push_back(xEst(i));
    for (int i = 0; i < 16; ++i) out.push_back(PEst.data()[i]);
    // ---- MPC: calc_ref_trajectory -> mpc_solve -> update, like mpc_simulation :372-376 -----------
    constexpr int T = 6;
    const int nc = (int)rd();
    Vec_f cx(nc), cy(nc), cyaw(nc), ck(nc, 0.0f), sp(nc);
    for (auto& v : cx) v = rd();
    for (auto& v : cy) v = rd();
    for (auto& v : cyaw) v = rd();
    for (auto& v : sp) v = rd();
    const float sx = rd(), sy = rd(), syaw = rd(), sv = rd();  // (argument evaluation order is unspecified)
    State state(sx, sy, syaw, sv);
    int target_ind = (int)rd();
    crb::M_XREF<T> xref;
    calc_ref_trajectory<T>(state, cx, cy, cyaw, ck, sp, 1.0f, target_ind, xref);
    for (int i = 0; i < 4 * T; ++i) out.push_back(xref.data()[i]);
    out.push_back((float)target_ind);
    int32_t status = -1;
    Vec_f output = mpc_solve<T>(state, xref, nullptr, &status);
    for (float v : output) out.push_back(v);
    out.push_back((float)status);
    const int a_start = 4 * T + (T - 1), delta_start = 4 * T;   // :54-60
    update(state, output[a_start], output[delta_start]);
    out.push_back(state.x); out.push_back(state.y); out.push_back(state.yaw); out.push_back(state.v);
    // ---- PF, like main() of src/particle_filter.cpp:270 ------------------------------------------
    constexpr int NP = 100;
    crb::Mat<4, NP> px;
    crb::Mat<NP, 1> pw;
    for (int i = 0; i < 4 * NP; ++i) px.data()[i] = rd();
    for (int i = 0; i < NP; ++i) pw.data()[i] = rd();
    const int nz = (int)rd();
    std::vector<crb::RowVector3f> z(nz);
    for (auto& it : z) { it(0, 0) = rd(); it(0, 1) = rd(); it(0, 2) = rd(); }
    crb::Vector2f u; u(0) = 1.0f; u(1) = 0.1f;
    crb::Matrix2f Rsim = crb::Matrix2f::Identity();
    Rsim(0, 0) = 1.0;
    Rsim(1, 1) = (30.0 / 180 * M_PI) * (30.0 / 180 * M_PI);
    const unsigned seed = (unsigned)rd();
    std::mt19937 gen{seed};
    std::normal_distribution<> gaussian_d{0, 1};
    crb::Vector4f xe; crb::Matrix4f Pe;
    pf_localization<NP>(px, pw, xe, Pe, z, u, Rsim, (float)(0.1 * 0.1), gen, gaussian_d);
    for (int i = 0; i < 4 * NP; ++i) out.push_back(px.data()[i]);
    for (int i = 0; i < NP; ++i) out.push_back(pw.data()[i]);
    for (int i = 0; i < 4; ++i) out.push_back(xe(i));
    for (int i = 0; i < 16; ++i) out.push_back(Pe.data()[i]);
    // the draws pf_localization made from its by-value copy of gen (:78): same sequence here
    std::mt19937 gen2{seed};
    std::normal_distribution<> d2{0, 1};
    for (int i = 0; i < 2 * NP; ++i) out.push_back((float)d2(gen2));
  } catch (const std::exception& e) {
    std::cerr << "reference_api: " << e.what() << std::endl;
    return 3;
  }
  FILE* f = std::fopen(argv[2], "wb");
  std::fwrite(out.data(), 4, out.size(), f);
  std::fclose(f);
  return 0;
}
