// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, not product code.
//
// A thin extern "C" driver over the UNMODIFIED reference headers
// (/root/reference/hector_mapping/include/hector_slam_lib, included where they lie; nothing is
// copied) compiled against oracle/shim (mini-Eigen + tf stub).  Built by oracle/Makefile into
// oracle/_ref/libhsref.so.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
// `--impl reference` legs may load it.
//
// Every entry point forwards to a reference call:
//   hsref_update            -> HectorSlamProcessor::update          slam_main/HectorSlamProcessor.h:71
//   hsref_match             -> MapRepMultiMap::matchData            slam_main/MapRepMultiMap.h:116
//   hsref_update_by_scan    -> MapRepMultiMap::updateByScan         slam_main/MapRepMultiMap.h:134
//   hsref_on_map_updated    -> MapRepMultiMap::onMapUpdated         slam_main/MapRepMultiMap.h:107
//   hsref_hessian_derivs    -> OccGridMapUtil::getCompleteHessianDerivs   map/OccGridMapUtil.h:64
//   hsref_match_level       -> ScanMatcher::matchData               matcher/ScanMatcher.h:54
//   hsref_update_level      -> OccGridMapBase::updateByScan         map/OccGridMapBase.h:121
//   hsref_get_prob          -> OccGridMapBase::getGridProbabilityMap map/OccGridMapBase.h:74
//   hsref_covariance_for_pose -> OccGridMapUtil::getCovarianceForPose / getCovMatrixWorldCoords  map/OccGridMapUtil.h:106-187
#include <iostream>
#include <vector>
#include <climits>
#include <cstring>
#include <cfloat>
#include <chrono>
#include <thread>
#include <atomic>
#include <streambuf>

#include "slam_main/HectorSlamProcessor.h"

namespace {

class NullBuf : public std::streambuf {
 protected:
  int overflow(int c) override { return c; }
};
NullBuf g_nullbuf;
std::streambuf* g_saved_cout = nullptr;

// mapRep is a protected member of the reference façade (HectorSlamProcessor.h:143); a subclass
// may name it.  Nothing of the reference's behaviour is overridden.
class Proc : public hectorslam::HectorSlamProcessor {
 public:
  Proc(float res, int sx, int sy, const Eigen::Vector2f& start, int levels)
      : hectorslam::HectorSlamProcessor(res, sx, sy, start, levels, 0, 0) {}
  hectorslam::MapRepresentationInterface* rep() { return mapRep; }
  hectorslam::GridMap& grid(int level) { return const_cast<hectorslam::GridMap&>(mapRep->getGridMap(level)); }
};

void fill(hectorslam::DataContainer& dc, const float* pts, int n, const float* origo) {
  dc.clear();
  for (int i = 0; i < n; ++i) dc.add(Eigen::Vector2f(pts[2 * i], pts[2 * i + 1]));
  if (origo)
    dc.setOrigo(Eigen::Vector2f(origo[0], origo[1]));
  else
    dc.setOrigo(Eigen::Vector2f(0.0f, 0.0f));
}

void put_pose(const Eigen::Vector3f& p, float* out) {
  out[0] = p[0];
  out[1] = p[1];
  out[2] = p[2];
}
// row-major 3x3 out (H is symmetric, so the order is moot; stated anyway)
void put_mat(const Eigen::Matrix3f& m, float* out) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out[3 * r + c] = m(r, c);
}

struct Handle {
  Proc* proc;
  float res;
  int sx, sy, levels;
  float startx, starty;
  float ffree, focc;
};

}  // namespace

extern "C" {

void hsref_silence(int on) {
  if (on && !g_saved_cout) {
    g_saved_cout = std::cout.rdbuf(&g_nullbuf);
  } else if (!on && g_saved_cout) {
    std::cout.rdbuf(g_saved_cout);
    g_saved_cout = nullptr;
  }
}

void* hsref_create(float res, int sx, int sy, float startx, float starty, int levels) {
  Handle* h = new Handle;
  h->proc = new Proc(res, sx, sy, Eigen::Vector2f(startx, starty), levels);
  h->res = res;
  h->sx = sx;
  h->sy = sy;
  h->levels = levels;
  h->startx = startx;
  h->starty = starty;
  h->ffree = 0.4f;   // GridMapLogOdds.h:117
  h->focc = 0.6f;    // GridMapLogOdds.h:118
  return h;
}

void hsref_destroy(void* hv) {
  Handle* h = static_cast<Handle*>(hv);
  delete h->proc;
  delete h;
}

void hsref_reset(void* hv) { static_cast<Handle*>(hv)->proc->reset(); }

void hsref_set_update_factors(void* hv, float ffree, float focc) {
  Handle* h = static_cast<Handle*>(hv);
  h->proc->setUpdateFactorFree(ffree);
  h->proc->setUpdateFactorOccupied(focc);
  h->ffree = ffree;
  h->focc = focc;
}

void hsref_set_map_update_thresholds(void* hv, float dist, float ang) {
  Handle* h = static_cast<Handle*>(hv);
  h->proc->setMapUpdateMinDistDiff(dist);
  h->proc->setMapUpdateMinAngleDiff(ang);
}

int hsref_levels(void* hv) { return static_cast<Handle*>(hv)->proc->getMapLevels(); }
int hsref_size_x(void* hv, int level) { return static_cast<Handle*>(hv)->proc->getGridMap(level).getSizeX(); }
int hsref_size_y(void* hv, int level) { return static_cast<Handle*>(hv)->proc->getGridMap(level).getSizeY(); }
float hsref_cell_length(void* hv, int level) { return static_cast<Handle*>(hv)->proc->getGridMap(level).getCellLength(); }
float hsref_scale_to_map(void* hv) { return static_cast<Handle*>(hv)->proc->getScaleToMap(); }

// world <-> level-map pose conversion exactly as the matcher does it (GridMapBase.h:226-239)
void hsref_map_coords_pose(void* hv, int level, const float world[3], float out[3]) {
  Handle* h = static_cast<Handle*>(hv);
  put_pose(h->proc->getGridMap(level).getMapCoordsPose(Eigen::Vector3f(world[0], world[1], world[2])), out);
}
void hsref_world_coords_pose(void* hv, int level, const float map[3], float out[3]) {
  Handle* h = static_cast<Handle*>(hv);
  put_pose(h->proc->getGridMap(level).getWorldCoordsPose(Eigen::Vector3f(map[0], map[1], map[2])), out);
}

void hsref_update(void* hv, const float* pts, int n, const float* origo, const float hint[3],
                  int map_without_matching, float out_pose[3], float out_cov[9]) {
  Handle* h = static_cast<Handle*>(hv);
  hectorslam::DataContainer dc;
  fill(dc, pts, n, origo);
  h->proc->update(dc, Eigen::Vector3f(hint[0], hint[1], hint[2]), map_without_matching != 0);
  put_pose(h->proc->getLastScanMatchPose(), out_pose);
  if (out_cov) put_mat(h->proc->getLastScanMatchCovariance(), out_cov);
}

void hsref_match(void* hv, const float hint[3], const float* pts, int n, const float* origo,
                 float out_pose[3], float cov_inout[9]) {
  Handle* h = static_cast<Handle*>(hv);
  hectorslam::DataContainer dc;
  fill(dc, pts, n, origo);
  Eigen::Matrix3f cov;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) cov(r, c) = cov_inout ? cov_inout[3 * r + c] : 0.0f;
  Eigen::Vector3f p = h->proc->rep()->matchData(Eigen::Vector3f(hint[0], hint[1], hint[2]), dc, cov);
  put_pose(p, out_pose);
  if (cov_inout) put_mat(cov, cov_inout);
}

void hsref_update_by_scan(void* hv, const float* pts, int n, const float* origo, const float pose[3]) {
  Handle* h = static_cast<Handle*>(hv);
  hectorslam::DataContainer dc;
  fill(dc, pts, n, origo);
  h->proc->rep()->updateByScan(dc, Eigen::Vector3f(pose[0], pose[1], pose[2]));
}

void hsref_on_map_updated(void* hv) { static_cast<Handle*>(hv)->proc->rep()->onMapUpdated(); }

void hsref_get_logodds(void* hv, int level, float* out) {
  Handle* h = static_cast<Handle*>(hv);
  hectorslam::GridMap& g = h->proc->grid(level);
  int n = g.getSizeX() * g.getSizeY();
  for (int i = 0; i < n; ++i) out[i] = g.getCell(i).logOddsVal;
}

void hsref_set_logodds(void* hv, int level, const float* in) {
  Handle* h = static_cast<Handle*>(hv);
  hectorslam::GridMap& g = h->proc->grid(level);
  int n = g.getSizeX() * g.getSizeY();
  for (int i = 0; i < n; ++i) g.getCell(i).logOddsVal = in[i];
  h->proc->rep()->onMapUpdated();  // invalidate the probability cache (MapRepMultiMap.h:107-114)
}

void hsref_get_prob(void* hv, int level, float* out) {
  Handle* h = static_cast<Handle*>(hv);
  const hectorslam::GridMap& g = h->proc->getGridMap(level);
  int n = g.getSizeX() * g.getSizeY();
  for (int i = 0; i < n; ++i) out[i] = g.getGridProbabilityMap(i);
}

// log-odds increments the map applies (GridMapLogOdds.h:187-203), recovered by applying them to
// a scratch cell — the members are protected.
void hsref_get_logodds_increments(void* hv, float out[2]) {
  Handle* h = static_cast<Handle*>(hv);
  GridMapLogOddsFunctions fn;
  fn.setUpdateFreeFactor(h->ffree);
  fn.setUpdateOccupiedFactor(h->focc);
  LogOddsCell c;
  c.resetGridCell();
  fn.updateSetFree(c);
  out[0] = c.logOddsVal;
  c.resetGridCell();
  fn.updateSetOccupied(c);
  out[1] = c.logOddsVal;
}

void hsref_hessian_derivs(void* hv, int level, const float pose_map[3], const float* pts_level, int n,
                          float H_out[9], float dTr_out[3]) {
  Handle* h = static_cast<Handle*>(hv);
  hectorslam::GridMap& g = h->proc->grid(level);
  hectorslam::OccGridMapUtilConfig<hectorslam::GridMap> util(&g);
  hectorslam::DataContainer dc;
  fill(dc, pts_level, n, 0);
  Eigen::Matrix3f H;
  Eigen::Vector3f dTr;
  util.getCompleteHessianDerivs(Eigen::Vector3f(pose_map[0], pose_map[1], pose_map[2]), dc, H, dTr);
  put_mat(H, H_out);
  put_pose(dTr, dTr_out);
}

void hsref_match_level(void* hv, int level, const float hint_world[3], const float* pts_level, int n,
                       int max_iterations, float out_pose[3], float out_cov[9]) {
  Handle* h = static_cast<Handle*>(hv);
  hectorslam::GridMap& g = h->proc->grid(level);
  hectorslam::OccGridMapUtilConfig<hectorslam::GridMap> util(&g);
  hectorslam::ScanMatcher<hectorslam::OccGridMapUtilConfig<hectorslam::GridMap> > matcher(0, 0);
  hectorslam::DataContainer dc;
  fill(dc, pts_level, n, 0);
  Eigen::Matrix3f cov = Eigen::Matrix3f::Zero();
  Eigen::Vector3f p =
      matcher.matchData(Eigen::Vector3f(hint_world[0], hint_world[1], hint_world[2]), util, dc, cov, max_iterations);
  put_pose(p, out_pose);
  if (out_cov) put_mat(cov, out_cov);
}

void hsref_update_level(void* hv, int level, const float* pts_level, int n, const float* origo_level,
                        const float pose_world[3]) {
  Handle* h = static_cast<Handle*>(hv);
  hectorslam::DataContainer dc;
  fill(dc, pts_level, n, origo_level);
  h->proc->grid(level).updateByScan(dc, Eigen::Vector3f(pose_world[0], pose_world[1], pose_world[2]));
  h->proc->rep()->onMapUpdated();
}

// OccGridMapUtil::getLikelihoodForState — map/OccGridMapUtil.h:189-221 (state in the level's cells)
float hsref_likelihood(void* hv, int level, const float pose_map[3], const float* pts_level, int n) {
  Handle* h = static_cast<Handle*>(hv);
  hectorslam::GridMap& g = h->proc->grid(level);
  hectorslam::OccGridMapUtilConfig<hectorslam::GridMap> util(&g);
  hectorslam::DataContainer dc;
  fill(dc, pts_level, n, 0);
  return util.getLikelihoodForState(Eigen::Vector3f(pose_map[0], pose_map[1], pose_map[2]), dc);
}

// OccGridMapUtil::getCovarianceForPose — map/OccGridMapUtil.h:106-160 (sigma points in the level's cells) and
// getCovMatrixWorldCoords — :162-187.  The function prints its likelihoods to std::cout (:139): silenced for the call.
void hsref_covariance_for_pose(void* hv, int level, const float pose_map[3], const float* pts_level, int n, float out_map[9],
                               float out_world[9]) {
  Handle* h = static_cast<Handle*>(hv);
  hectorslam::GridMap& g = h->proc->grid(level);
  hectorslam::OccGridMapUtilConfig<hectorslam::GridMap> util(&g);
  hectorslam::DataContainer dc;
  fill(dc, pts_level, n, 0);
  std::streambuf* saved = std::cout.rdbuf(&g_nullbuf);
  Eigen::Matrix3f cm = util.getCovarianceForPose(Eigen::Vector3f(pose_map[0], pose_map[1], pose_map[2]), dc);
  std::cout.rdbuf(saved);
  Eigen::Matrix3f cw = util.getCovMatrixWorldCoords(cm);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      out_map[3 * r + c] = cm(r, c);
      out_world[3 * r + c] = cw(r, c);
    }
}

// Batch of independent matches against the handle's current (frozen) map.
// nthreads <= 1: the handle's own processor, calling thread.
// nthreads  > 1: OccGridMapUtil is not re-entrant (mutable members, OccGridMapUtil.h:376-378), so
//                every worker builds a private HectorSlamProcessor, clones the log-odds planes into
//                it (untimed) and runs MapRepMultiMap::matchData on its contiguous share.
// Returns the wall-clock seconds spent inside the matching phase (max over workers).
double hsref_match_batch(void* hv, int B, const float* hints, const float* pts, const int* offsets,
                         float* out_poses, float* out_cov, int nthreads) {
  Handle* h = static_cast<Handle*>(hv);
  typedef std::chrono::steady_clock clk;
  if (nthreads <= 1) {
    hectorslam::DataContainer dc;
    clk::time_point t0 = clk::now();
    for (int b = 0; b < B; ++b) {
      int n = offsets[b + 1] - offsets[b];
      fill(dc, pts + 2 * (size_t)offsets[b], n, 0);
      Eigen::Matrix3f cov = Eigen::Matrix3f::Zero();
      Eigen::Vector3f p =
          h->proc->rep()->matchData(Eigen::Vector3f(hints[3 * b], hints[3 * b + 1], hints[3 * b + 2]), dc, cov);
      put_pose(p, out_poses + 3 * b);
      if (out_cov) put_mat(cov, out_cov + 9 * b);
    }
    return std::chrono::duration<double>(clk::now() - t0).count();
  }
  std::vector<std::vector<float> > planes(h->levels);
  for (int l = 0; l < h->levels; ++l) {
    planes[l].resize((size_t)hsref_size_x(hv, l) * hsref_size_y(hv, l));
    hsref_get_logodds(hv, l, planes[l].data());
  }
  std::vector<double> secs(nthreads, 0.0);
  std::atomic<int> ready(0);
  std::atomic<int> go(0);
  std::vector<std::thread> workers;
  for (int t = 0; t < nthreads; ++t) {
    workers.emplace_back([&, t]() {
      Proc* p = new Proc(h->res, h->sx, h->sy, Eigen::Vector2f(h->startx, h->starty), h->levels);
      p->setUpdateFactorFree(h->ffree);
      p->setUpdateFactorOccupied(h->focc);
      for (int l = 0; l < h->levels; ++l) {
        hectorslam::GridMap& g = p->grid(l);
        const std::vector<float>& src = planes[l];
        for (size_t i = 0; i < src.size(); ++i) g.getCell((int)i).logOddsVal = src[i];
      }
      p->rep()->onMapUpdated();
      int b0 = (int)((long long)B * t / nthreads), b1 = (int)((long long)B * (t + 1) / nthreads);
      hectorslam::DataContainer dc;
      // untimed warm pass over this worker's share: touches the private planes and fills the
      // probability cache, so the timed pass below is the reference's steady state (its best case)
      for (int b = b0; b < b1; ++b) {
        int n = offsets[b + 1] - offsets[b];
        fill(dc, pts + 2 * (size_t)offsets[b], n, 0);
        Eigen::Matrix3f cov = Eigen::Matrix3f::Zero();
        p->rep()->matchData(Eigen::Vector3f(hints[3 * b], hints[3 * b + 1], hints[3 * b + 2]), dc, cov);
      }
      ready.fetch_add(1);
      while (go.load() == 0) std::this_thread::yield();
      clk::time_point t0 = clk::now();
      for (int b = b0; b < b1; ++b) {
        int n = offsets[b + 1] - offsets[b];
        fill(dc, pts + 2 * (size_t)offsets[b], n, 0);
        Eigen::Matrix3f cov = Eigen::Matrix3f::Zero();
        Eigen::Vector3f r =
            p->rep()->matchData(Eigen::Vector3f(hints[3 * b], hints[3 * b + 1], hints[3 * b + 2]), dc, cov);
        put_pose(r, out_poses + 3 * b);
        if (out_cov) put_mat(cov, out_cov + 9 * b);
      }
      secs[t] = std::chrono::duration<double>(clk::now() - t0).count();
      delete p;
    });
  }
  while (ready.load() < nthreads) std::this_thread::yield();
  go.store(1);
  for (size_t t = 0; t < workers.size(); ++t) workers[t].join();
  double mx = 0.0;
  for (int t = 0; t < nthreads; ++t) mx = secs[t] > mx ? secs[t] : mx;
  return mx;
}

}  // extern "C"
