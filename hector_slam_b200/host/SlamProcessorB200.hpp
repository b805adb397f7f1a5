// SlamProcessorB200.hpp — standalone C++ host façade over the C-ABI (no Eigen, no ROS).
//
// Mirrors hectorslam::HectorSlamProcessor (hector_mapping/include/hector_slam_lib/slam_main/
// HectorSlamProcessor.h:50-154) method for method — update(), reset(), the getters and setters —
// for callers that do not have the reference's headers around.  update() is one C-ABI call
// (hsb_slam_update): match, gate on pose difference, map write and epoch bump run back to back on
// the GPU; updateUnfused() keeps the reference's host-side control flow over the two heavy calls.  Where the reference's own
// headers are available, use MapRepB200.h instead: it plugs into the unmodified reference façade.
#ifndef HECTOR_SLAM_B200_SLAMPROCESSOR_HPP
#define HECTOR_SLAM_B200_SLAMPROCESSOR_HPP

#include <cfloat>
#include <cmath>
#include <stdexcept>
#include <string>

#include "hector_slam_b200.h"

namespace hsb200 {

struct Pose {
  float x, y, psi;
};

class SlamProcessor {
 public:
  // HectorSlamProcessor.h:54
  SlamProcessor(float mapResolution, int mapSizeX, int mapSizeY, float startX, float startY, int multi_res_size,
                int device = 0)
      : h_(0) {
    hsb_config cfg = hsb_config();
    cfg.map_resolution = mapResolution;
    cfg.map_size_x = mapSizeX;
    cfg.map_size_y = mapSizeY;
    cfg.start_x = startX;
    cfg.start_y = startY;
    cfg.levels = multi_res_size;
    cfg.device = device;
    if (hsb_create(&cfg, &h_) != HSB_OK) throw std::runtime_error(std::string("hsb_create: ") + hsb_last_error(0));
    reset();
    setMapUpdateMinDistDiff(0.4f * 1.0f);    // :62
    setMapUpdateMinAngleDiff(0.13f * 1.0f);  // :63
    for (int i = 0; i < 9; ++i) lastScanMatchCov_[i] = 0.f;
  }
  ~SlamProcessor() { hsb_destroy(h_); }
  SlamProcessor(const SlamProcessor&) = delete;
  SlamProcessor& operator=(const SlamProcessor&) = delete;

  // HectorSlamProcessor.h:71-113.  points: n x 2 floats in level-0 cell units (DataContainer).
  void update(const float* points_xy, int n, const float origo[2], const Pose& poseHintWorld,
              bool map_without_matching = false) {
    // one fused call: match (:78), gate (:89), updateByScan (:91), onMapUpdated (:93) run back to back on the
    // device; the host only mirrors the two poses the getters expose
    const float hint[3] = {poseHintWorld.x, poseHintWorld.y, poseHintWorld.psi};
    float out[3];
    int updated = 0;
    check(hsb_slam_update(h_, hint, points_xy, n, origo, map_without_matching ? 1 : 0, out, lastScanMatchCov_, &updated));
    lastScanMatchPose_ = Pose{out[0], out[1], out[2]};  // :83
    if (updated) lastMapUpdatePose_ = lastScanMatchPose_;  // :94
  }

  // The same step with the control flow on the host (two C-ABI calls and a round trip in between), as the
  // reference's own HectorSlamProcessor drives MapRepB200; kept for comparison and tests.
  void updateUnfused(const float* points_xy, int n, const float origo[2], const Pose& poseHintWorld,
                     bool map_without_matching = false) {
    Pose newPose;
    if (!map_without_matching) {
      const float hint[3] = {poseHintWorld.x, poseHintWorld.y, poseHintWorld.psi};
      float out[3];
      check(hsb_match_data(h_, hint, points_xy, n, origo, out, lastScanMatchCov_));  // :78
      newPose = Pose{out[0], out[1], out[2]};
    } else {
      newPose = poseHintWorld;  // :80
    }
    lastScanMatchPose_ = newPose;  // :83
    if (poseDifferenceLargerThan(newPose, lastMapUpdatePose_, minDist_, minAngle_) || map_without_matching) {  // :89
      const float p[3] = {newPose.x, newPose.y, newPose.psi};
      check(hsb_update_by_scan(h_, points_xy, n, origo, p));  // :91
      check(hsb_on_map_updated(h_));                          // :93
      lastMapUpdatePose_ = newPose;                           // :94
      check(hsb_set_last_map_update_pose(h_, p));             // keep the fused path's gate state in step (mixing allowed)
    }
  }

  // HectorSlamProcessor.h:115-124
  void reset() {
    lastMapUpdatePose_ = Pose{FLT_MAX, FLT_MAX, FLT_MAX};
    lastScanMatchPose_ = Pose{0.f, 0.f, 0.f};
    check(hsb_reset(h_));
  }

  const Pose& getLastScanMatchPose() const { return lastScanMatchPose_; }
  const float* getLastScanMatchCovariance() const { return lastScanMatchCov_; }
  float getScaleToMap() const { return hsb_get_scale_to_map(h_); }
  int getMapLevels() const { return hsb_get_map_levels(h_); }
  void setUpdateFactorFree(float f) { check(hsb_set_update_factor_free(h_, f)); }
  void setUpdateFactorOccupied(float f) { check(hsb_set_update_factor_occupied(h_, f)); }
  void setMapUpdateMinDistDiff(float minDist) {
    minDist_ = minDist;
    check(hsb_set_map_update_min_dist_diff(h_, minDist));
  }
  void setMapUpdateMinAngleDiff(float angleChange) {
    minAngle_ = angleChange;
    check(hsb_set_map_update_min_angle_diff(h_, angleChange));
  }
  // getGridMap(level): the log-odds plane, row-major [sizeY][sizeX]
  void getGridMap(int level, float* logodds_out) const { check(hsb_download_level(h_, level, logodds_out)); }
  hsb_handle* handle() { return h_; }

  // util::poseDifferenceLargerThan, util/UtilFunctions.h:73-92 (float norm, double pi arithmetic)
  static bool poseDifferenceLargerThan(const Pose& p1, const Pose& p2, float distThresh, float angleThresh) {
    const float dx = p1.x - p2.x, dy = p1.y - p2.y;
    if (std::sqrt(dx * dx + dy * dy) > distThresh) return true;
    float angleDiff = p1.psi - p2.psi;
    if (angleDiff > M_PI) {
      angleDiff -= M_PI * 2.0f;
    } else if (angleDiff < -M_PI) {
      angleDiff += M_PI * 2.0f;
    }
    return std::fabs(angleDiff) > angleThresh;
  }

 private:
  void check(int st) const {
    if (st != HSB_OK) throw std::runtime_error(std::string("hector_slam_b200: ") + hsb_last_error(h_));
  }
  hsb_handle* h_;
  Pose lastMapUpdatePose_, lastScanMatchPose_;
  float lastScanMatchCov_[9];
  float minDist_, minAngle_;
};

}  // namespace hsb200
#endif
