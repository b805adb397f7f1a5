// host_capi.cpp — extern "C" shim over hsb200::SlamProcessor so that the C++ host façade can be
// driven from ctypes in the tests (the façade itself is header-only C++).
#include "SlamProcessorB200.hpp"

#include <cstring>

extern "C" {

void* hsbp_create(float res, int sx, int sy, float startx, float starty, int levels, int device) {
  try {
    return new hsb200::SlamProcessor(res, sx, sy, startx, starty, levels, device);
  } catch (const std::exception&) {
    return 0;
  }
}
void hsbp_destroy(void* p) { delete static_cast<hsb200::SlamProcessor*>(p); }

int hsbp_update(void* p, const float* pts, int n, const float* origo, const float hint[3], int map_without_matching,
                float out_pose[3], float out_cov[9]) {
  hsb200::SlamProcessor* s = static_cast<hsb200::SlamProcessor*>(p);
  try {
    s->update(pts, n, origo, hsb200::Pose{hint[0], hint[1], hint[2]}, map_without_matching != 0);
  } catch (const std::exception&) {
    return -1;
  }
  const hsb200::Pose& q = s->getLastScanMatchPose();
  out_pose[0] = q.x;
  out_pose[1] = q.y;
  out_pose[2] = q.psi;
  if (out_cov) std::memcpy(out_cov, s->getLastScanMatchCovariance(), 9 * sizeof(float));
  return 0;
}
int hsbp_reset(void* p) {
  try {
    static_cast<hsb200::SlamProcessor*>(p)->reset();
  } catch (const std::exception&) {
    return -1;
  }
  return 0;
}
void hsbp_set_update_factors(void* p, float ffree, float focc) {
  static_cast<hsb200::SlamProcessor*>(p)->setUpdateFactorFree(ffree);
  static_cast<hsb200::SlamProcessor*>(p)->setUpdateFactorOccupied(focc);
}
void hsbp_set_map_update_thresholds(void* p, float dist, float ang) {
  static_cast<hsb200::SlamProcessor*>(p)->setMapUpdateMinDistDiff(dist);
  static_cast<hsb200::SlamProcessor*>(p)->setMapUpdateMinAngleDiff(ang);
}
int hsbp_get_grid_map(void* p, int level, float* out) {
  try {
    static_cast<hsb200::SlamProcessor*>(p)->getGridMap(level, out);
  } catch (const std::exception&) {
    return -1;
  }
  return 0;
}
int hsbp_pose_difference_larger_than(const float a[3], const float b[3], float dist, float ang) {
  return hsb200::SlamProcessor::poseDifferenceLargerThan(hsb200::Pose{a[0], a[1], a[2]}, hsb200::Pose{b[0], b[1], b[2]},
                                                         dist, ang)
             ? 1
             : 0;
}
void* hsbp_handle(void* p) { return static_cast<hsb200::SlamProcessor*>(p)->handle(); }
}
