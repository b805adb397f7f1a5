// oracle/shim/tf/LinearMath/Transform.h — TEST INFRASTRUCTURE.
// hector_slam_lib/util/UtilFunctions.h:33 includes this ROS header only for getYawFromQuat
// (:94-97), which is not on the scan-matching path.  The stub supplies the three names that
// function mentions.  It also includes <math.h>: UtilFunctions.h:88 calls unqualified
// `abs(float)`, which resolves to the floating overload only when the C header's global
// overloads are visible (as they are through the real tf headers); with <cmath> alone g++ picks
// `int abs(int)` and the angle gate of poseDifferenceLargerThan silently truncates.
#ifndef HS_ORACLE_SHIM_TF_TRANSFORM_H
#define HS_ORACLE_SHIM_TF_TRANSFORM_H
#include <math.h>
#include <stdlib.h>
#include <cmath>
#include <cstdlib>
namespace geometry_msgs {
struct Quaternion {
  double x, y, z, w;
};
}  // namespace geometry_msgs
namespace tf {
struct Quaternion {
  double x_, y_, z_, w_;
  Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
};
inline double getYaw(const Quaternion& q) {
  return atan2(2.0 * (q.w_ * q.z_ + q.x_ * q.y_), 1.0 - 2.0 * (q.y_ * q.y_ + q.z_ * q.z_));
}
}  // namespace tf
#endif
