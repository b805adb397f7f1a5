// oracle/ros_conv_driver.cpp — TEST INFRASTRUCTURE, not product code.
//
// The two scan -> DataContainer converters of the reference NODE, compiled from the reference source where it
// lies: oracle/Makefile cuts hector_mapping/src/HectorMappingRos.cpp:483-542 (rosLaserScanToDataContainer and
// rosPointCloudToDataContainer, unmodified text) into the build directory oracle/_ref/ (git-ignored, never
// committed) and this file #includes that extract as member functions of a stand-in `HectorMappingRos` class that
// has only the four parameter members the functions read.  The ROS message types are restated as plain structs with
// the same field names; hectorslam::DataContainer is the reference's own header on the Eigen shim.
//
// Third-party arithmetic not in /root/reference: tf (ROS geometry, Bullet LinearMath; tfScalar = double).
// The node calls `laserTransform * tf::Vector3(x, y, z)`, `getOrigin()`, `.x()/.y()/.z()`.  Restated below from the
// published tf/LinearMath headers (geometry 1.12/1.13, ROS Melodic/Noetic):
//   Transform::operator*(const Vector3& x) = Vector3(m_basis[0].dot(x) + m_origin.x(), [1].., [2]..)   (Transform.h)
//   Vector3::dot(v) = m_floats[0]*v[0] + m_floats[1]*v[1] + m_floats[2]*v[2]                           (Vector3.h)
// all in double, no contraction on a stock x86-64 build.
#include <math.h>
#include <cmath>
#include <cstring>
#include <iostream>
#include <vector>

#include <Eigen/Core>
#include "scan/DataPointContainer.h"

namespace geometry_msgs {
struct Point32 {
  float x, y, z;
};
}  // namespace geometry_msgs
namespace sensor_msgs {
struct LaserScan {
  float angle_min, angle_increment, range_min, range_max;
  std::vector<float> ranges;
};
struct PointCloud {
  std::vector<geometry_msgs::Point32> points;
};
}  // namespace sensor_msgs
namespace tf {
typedef double tfScalar;
class Vector3 {
 public:
  tfScalar m_floats[4];
  Vector3() {}
  Vector3(const tfScalar& x, const tfScalar& y, const tfScalar& z) {
    m_floats[0] = x;
    m_floats[1] = y;
    m_floats[2] = z;
    m_floats[3] = tfScalar(0.);
  }
  tfScalar dot(const Vector3& v) const {
    return m_floats[0] * v.m_floats[0] + m_floats[1] * v.m_floats[1] + m_floats[2] * v.m_floats[2];
  }
  const tfScalar& x() const { return m_floats[0]; }
  const tfScalar& y() const { return m_floats[1]; }
  const tfScalar& z() const { return m_floats[2]; }
};
class Transform {
 public:
  Vector3 m_basis[3];
  Vector3 m_origin;
  Vector3 operator*(const Vector3& x) const {
    return Vector3(m_basis[0].dot(x) + m_origin.x(), m_basis[1].dot(x) + m_origin.y(), m_basis[2].dot(x) + m_origin.z());
  }
  const Vector3& getOrigin() const { return m_origin; }
};
class StampedTransform : public Transform {};
}  // namespace tf

class HectorMappingRos {
 public:
  void rosLaserScanToDataContainer(const sensor_msgs::LaserScan& scan, hectorslam::DataContainer& dataContainer, float scaleToMap);
  void rosPointCloudToDataContainer(const sensor_msgs::PointCloud& pointCloud, const tf::StampedTransform& laserTransform,
                                    hectorslam::DataContainer& dataContainer, float scaleToMap);
  float p_sqr_laser_min_dist_;
  float p_sqr_laser_max_dist_;
  float p_laser_z_min_value_;
  float p_laser_z_max_value_;
};

#include "_ref/ros_conv_extract.inc"  // HectorMappingRos.cpp:483-542, cut at build time (oracle/Makefile)

static int dump(const hectorslam::DataContainer& dc, float* out_xy, float* out_origo) {
  const int n = dc.getSize();
  for (int i = 0; i < n; ++i) {
    out_xy[2 * i] = dc.getVecEntry(i)[0];
    out_xy[2 * i + 1] = dc.getVecEntry(i)[1];
  }
  if (out_origo) {
    const Eigen::Vector2f o = dc.getOrigo();
    out_origo[0] = o[0];
    out_origo[1] = o[1];
  }
  return n;
}

extern "C" {

int hsref_scan_to_points(const float* ranges, int n_beams, float angle_min, float angle_increment, float range_min,
                         float range_max, float scale_to_map, float* out_xy) {
  sensor_msgs::LaserScan scan;
  scan.angle_min = angle_min;
  scan.angle_increment = angle_increment;
  scan.range_min = range_min;
  scan.range_max = range_max;
  scan.ranges.assign(ranges, ranges + n_beams);
  hectorslam::DataContainer dc;
  HectorMappingRos node;
  node.rosLaserScanToDataContainer(scan, dc, scale_to_map);
  return dump(dc, out_xy, nullptr);
}

// transform: 12 doubles, rows of [R | t] (base_frame <- laser frame)
int hsref_cloud_to_points(const float* xyz, int n, const double* transform, float sqr_min_dist, float sqr_max_dist,
                          float z_min, float z_max, float scale_to_map, float* out_xy, float* out_origo) {
  sensor_msgs::PointCloud cloud;
  cloud.points.resize(n);
  if (n > 0) memcpy(cloud.points.data(), xyz, (size_t)n * sizeof(geometry_msgs::Point32));
  tf::StampedTransform T;
  for (int r = 0; r < 3; ++r) T.m_basis[r] = tf::Vector3(transform[4 * r], transform[4 * r + 1], transform[4 * r + 2]);
  T.m_origin = tf::Vector3(transform[3], transform[7], transform[11]);
  hectorslam::DataContainer dc;
  HectorMappingRos node;
  node.p_sqr_laser_min_dist_ = sqr_min_dist;
  node.p_sqr_laser_max_dist_ = sqr_max_dist;
  node.p_laser_z_min_value_ = z_min;
  node.p_laser_z_max_value_ = z_max;
  node.rosPointCloudToDataContainer(cloud, T, dc, scale_to_map);
  return dump(dc, out_xy, out_origo);
}

}  // extern "C"
