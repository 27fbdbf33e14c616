// Compile-only translation unit: every reference-signature shim of include/crb/reference_api.hpp is
// instantiated with the argument types a CppRobotics main() passes (tests/test_reference_api.py compiles it
// with -std=c++11 -Wall -Werror and links it against libcrb.so; it is never run).
#include <crb/reference_api.hpp>

int main() {
  crb::Vector4f xEst = crb::Vector4f::Zero();
  crb::Matrix4f PEst = crb::Matrix4f::Identity(), Q = crb::Matrix4f::Identity();
  crb::Matrix2f R = crb::Matrix2f::Identity();
  crb::Vector2f z = crb::Vector2f::Zero(), u = crb::Vector2f::Zero();
  ekf_estimation(xEst, PEst, z, u, Q, R);                                 // extended_kalman_filter.cpp:64

  std::mt19937 gen(1);
  std::normal_distribution<> gaussian_d(0, 1);
  std::uniform_real_distribution<> uni_d(1.0, 2.0);
  crb::Mat<4, 100> px = crb::Mat<4, 100>::Zero();
  crb::Mat<100, 1> pw = crb::Mat<100, 1>::Zero();
  std::vector<crb::RowVector3f> zs;
  pf_localization<100>(px, pw, xEst, PEst, zs, u, R, 0.01f, gen, gaussian_d);   // particle_filter.cpp:73
  resampling<100>(px, pw, gen, uni_d);                                    // particle_filter.cpp:120

  cpprobotics::State st(0.f, 0.f, 0.f, 0.f);
  crb::M_XREF<6> xref = crb::M_XREF<6>::Zero();
  cpprobotics::Vec_f sol = mpc_solve<6>(st, xref);                        // model_predictive_control.cpp:255
  update(st, sol[0], sol[1]);                                             // :69
  cpprobotics::Vec_f cx(10, 0.f), cy(10, 0.f), cyaw(10, 0.f), ck(10, 0.f), sp(10, 1.f);
  int target_ind = 0;
  calc_ref_trajectory<6>(st, cx, cy, cyaw, ck, sp, 1.0f, target_ind, xref);   // :130

  crb::Vector4f B4 = crb::Vector4f::Zero();
  crb::Matrix4f X4 = solve_DARE(Q, B4, Q, 1.0f);                          // lqr_steer_control.cpp:75
  crb::RowVector4f K4 = dlqr(Q, B4, Q, 1.0f);                             // :92
  crb::Matrix5f A5 = crb::Matrix5f::Identity();
  crb::Matrix52f B5 = crb::Matrix52f::Zero();
  crb::Matrix5f X5 = solve_DARE(A5, B5, A5, R);                           // lqr_speed_steer_control.cpp:85
  crb::Matrix25f K5 = dlqr(A5, B5, A5, R);                                // :101
  return (int)(X4(0, 0) + K4(0, 0) + X5(0, 0) + K5(0, 0));
}
