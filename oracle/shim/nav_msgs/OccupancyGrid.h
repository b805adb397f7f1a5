// oracle/shim/nav_msgs/OccupancyGrid.h — TEST INFRASTRUCTURE.  Stand-in for nav_msgs/OccupancyGrid
// (MapMetaData info; int8[] data) and its ConstPtr (boost::shared_ptr in ROS1; std::shared_ptr here).
#ifndef HS_ORACLE_SHIM_NAV_MSGS_OCCUPANCYGRID_H
#define HS_ORACLE_SHIM_NAV_MSGS_OCCUPANCYGRID_H
#include "MapMetaData.h"
namespace nav_msgs {
struct OccupancyGrid {
  MapMetaData info;
  std::vector<int8_t> data;
};
typedef std::shared_ptr<const OccupancyGrid> OccupancyGridConstPtr;
}  // namespace nav_msgs
#endif
