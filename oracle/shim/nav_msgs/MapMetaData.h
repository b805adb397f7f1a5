// oracle/shim/nav_msgs/MapMetaData.h — TEST INFRASTRUCTURE.
// Plain-struct stand-in for the ROS message nav_msgs/MapMetaData (fields and types as in the .msg definition:
// float32 resolution, uint32 width / height, geometry_msgs/Pose origin with float64 members), enough for
// hector_map_tools/HectorMapTools.h to compile unmodified in oracle/_ref.
#ifndef HS_ORACLE_SHIM_NAV_MSGS_MAPMETADATA_H
#define HS_ORACLE_SHIM_NAV_MSGS_MAPMETADATA_H
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <memory>
#include <vector>
namespace nav_msgs {
struct MapMetaData {
  float resolution;
  uint32_t width, height;
  struct {
    struct { double x, y, z; } position;
    struct { double x, y, z, w; } orientation;
  } origin;
};
}  // namespace nav_msgs
#endif
