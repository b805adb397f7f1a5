// oracle/dropin_check.cpp — TEST INFRASTRUCTURE.
//
// Runs the UNMODIFIED reference façade (HectorSlamProcessor::update, slam_main/HectorSlamProcessor.h:71)
// twice over the same synthetic scan sequence: once with the reference's own CPU map representation
// (MapRepMultiMap) and once with hector_slam_b200/host/MapRepB200.h plugged into the same façade,
// then compares every pose, the returned covariance and the final maps (read back through
// getGridMap(), i.e. through the lazily synchronised host mirror).
//
// Built by oracle/Makefile into oracle/_ref/dropin_check (needs /root/reference for the headers);
// the binary travels to the GPU box and is run by tests/test_gpu_dropin.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>
#include <streambuf>
#include <vector>

#include "MapRepB200.h"

namespace {

class NullBuf : public std::streambuf {
 protected:
  int overflow(int c) override { return c; }
};

struct CountingMutex : public MapLockerInterface {
  int locks = 0, unlocks = 0;
  virtual void lockMap() { ++locks; }
  virtual void unlockMap() { ++unlocks; }
};

// 16 m x 10 m room with three pillars, analytic ray casting.
struct World {
  double hx = 8.0, hy = 5.0;
  double pil[3][3] = {{-3.0, 1.5, 0.8}, {2.5, -2.0, 0.6}, {4.5, 2.5, 1.0}};
  double cast(double x, double y, double a) const {
    double dx = std::cos(a), dy = std::sin(a), t = 1e9;
    if (dx > 0) t = std::min(t, (hx - x) / dx);
    if (dx < 0) t = std::min(t, (-hx - x) / dx);
    if (dy > 0) t = std::min(t, (hy - y) / dy);
    if (dy < 0) t = std::min(t, (-hy - y) / dy);
    for (int k = 0; k < 3; ++k) {
      double ox = x - pil[k][0], oy = y - pil[k][1];
      double b = ox * dx + oy * dy, c = ox * ox + oy * oy - pil[k][2] * pil[k][2];
      double disc = b * b - c;
      if (disc >= 0) {
        double t0 = -b - std::sqrt(disc);
        if (t0 > 1e-9) t = std::min(t, t0);
      }
    }
    return t;
  }
};

void make_scan(const World& w, const double pose[3], std::mt19937& rng, float scale, hectorslam::DataContainer& dc) {
  std::normal_distribution<double> noise(0.0, 0.01);
  dc.clear();
  dc.setOrigo(Eigen::Vector2f(0.0f, 0.0f));
  float angle = -135.0f * 3.14159265f / 180.0f;
  const float inc = 0.25f * 3.14159265f / 180.0f;
  for (int i = 0; i < 1081; ++i) {
    float dist = static_cast<float>(w.cast(pose[0], pose[1], pose[2] + angle) + noise(rng));
    if (dist > 0.1f && dist < 29.9f) {  // HectorMappingRos.cpp:499
      dist *= scale;
      dc.add(Eigen::Vector2f(std::cos(angle) * dist, std::sin(angle) * dist));
    }
    angle += inc;
  }
}

double wrap(double a) { return std::fabs(std::remainder(a, 2.0 * 3.14159265358979323846)); }

}  // namespace

int main(int argc, char** argv) {
  const int steps = argc > 1 ? std::atoi(argv[1]) : 40;
  NullBuf nb;
  std::streambuf* saved = std::cout.rdbuf(&nb);  // the reference prints level banners / clamp notes

  const float res = 0.05f;
  const int size = 1024, levels = 3;
  hectorslam::HectorSlamProcessor cpu(res, size, size, Eigen::Vector2f(0.5f, 0.5f), levels);
  hectorslam::HectorSlamProcessorB200 gpu(res, size, size, Eigen::Vector2f(0.5f, 0.5f), levels);
  hectorslam::HectorSlamProcessor* procs[2] = {&cpu, &gpu};
  CountingMutex* mtx[2] = {new CountingMutex, new CountingMutex};
  for (int k = 0; k < 2; ++k) {
    procs[k]->setUpdateFactorFree(0.4f);
    procs[k]->setUpdateFactorOccupied(0.9f);
    procs[k]->setMapUpdateMinDistDiff(0.2f);   // some scans write the map, some do not (gate logic)
    procs[k]->setMapUpdateMinAngleDiff(0.06f);
    procs[k]->addMapMutex(0, mtx[k]);
  }

  World world;
  std::mt19937 rng(7);
  double pose[3] = {-4.0, -2.5, 0.3};
  Eigen::Vector3f hint[2];
  hint[0] = hint[1] = Eigen::Vector3f(static_cast<float>(pose[0]), static_cast<float>(pose[1]), static_cast<float>(pose[2]));
  double max_dp = 0, max_da = 0, max_dcov = 0, max_truth = 0;
  hectorslam::DataContainer dc;
  for (int s = 0; s < steps; ++s) {
    make_scan(world, pose, rng, cpu.getScaleToMap(), dc);
    for (int k = 0; k < 2; ++k) procs[k]->update(dc, hint[k]);
    const Eigen::Vector3f a = cpu.getLastScanMatchPose(), b = gpu.getLastScanMatchPose();
    max_dp = std::max(max_dp, static_cast<double>(std::max(std::fabs(a[0] - b[0]), std::fabs(a[1] - b[1]))));
    max_da = std::max(max_da, wrap(static_cast<double>(a[2]) - b[2]));
    max_truth = std::max(max_truth, std::max(std::fabs(a[0] - pose[0]), std::fabs(a[1] - pose[1])));
    const Eigen::Matrix3f ca = cpu.getLastScanMatchCovariance(), cb = gpu.getLastScanMatchCovariance();
    double scale = 1e-6, diff = 0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        scale = std::max(scale, static_cast<double>(std::fabs(ca(r, c))));
        diff = std::max(diff, static_cast<double>(std::fabs(ca(r, c) - cb(r, c))));
      }
    if (s > 0) max_dcov = std::max(max_dcov, diff / scale);
    hint[0] = a;  // each side feeds its own estimate back, like the node (HectorMappingRos.cpp:314)
    hint[1] = b;
    pose[0] += 0.09;
    pose[1] += 0.035;
    pose[2] += 0.015;
  }
  // final maps through getGridMap() (host mirror on the B200 side)
  long differing = 0, touched = 0;
  double max_cell = 0;
  for (int l = 0; l < levels; ++l) {
    const hectorslam::GridMap& ma = cpu.getGridMap(l);
    const hectorslam::GridMap& mb = gpu.getGridMap(l);
    const int n = ma.getSizeX() * ma.getSizeY();
    for (int i = 0; i < n; ++i) {
      const float va = ma.getCell(i).logOddsVal, vb = mb.getCell(i).logOddsVal;
      if (va != 0.0f) ++touched;
      const double d = std::fabs(static_cast<double>(va) - vb);
      if (d > 1e-5) ++differing;
      if (d > max_cell) max_cell = d;
    }
  }
  const int upd_cpu = cpu.getGridMap(0).getUpdateIndex(), upd_gpu = gpu.getGridMap(0).getUpdateIndex();
  std::cout.rdbuf(saved);
  std::printf("steps=%d max_dpos=%.3e max_dang=%.3e max_dcov_rel=%.3e err_vs_truth=%.3e\n", steps, max_dp, max_da, max_dcov,
              max_truth);
  std::printf("cells touched=%ld differing(>1e-5)=%ld max_cell_diff=%.3e\n", touched, differing, max_cell);
  std::printf("map writes: cpu updateIndex=%d gpu updateIndex=%d; mutex locks cpu=%d/%d gpu=%d/%d\n", upd_cpu, upd_gpu,
              mtx[0]->locks, mtx[0]->unlocks, mtx[1]->locks, mtx[1]->unlocks);
  bool ok = max_dp <= 1e-4 && max_da <= 1e-4 && differing <= 2 && upd_cpu == upd_gpu &&
            mtx[0]->locks == mtx[1]->locks && mtx[1]->locks == mtx[1]->unlocks && max_truth < 0.05;
  std::printf("%s\n", ok ? "DROPIN OK" : "DROPIN MISMATCH");
  return ok ? 0 : 1;
}
