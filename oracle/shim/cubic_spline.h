// oracle/shim/cubic_spline.h — stand-in for the reference's include/cubic_spline.h (dynamic-size Eigen QR,
// out of the hot path): only what main() of src/model_predictive_control.cpp:467-492 names, so that the
// translation unit compiles.  Never executed.
#ifndef CRB_SHIM_CUBIC_SPLINE_
#define CRB_SHIM_CUBIC_SPLINE_
#include <array>
#include "cpprobotics_types.h"
namespace cpprobotics {
struct Spline2D {
  Vec_f s;
  Spline2D(Vec_f, Vec_f) { s.push_back(0.0f); }
  Poi_f calc_postion(float) { return Poi_f{{0.0f, 0.0f}}; }
  float calc_yaw(float) { return 0.0f; }
  float calc_curvature(float) { return 0.0f; }
};
}
#endif
