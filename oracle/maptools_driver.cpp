// oracle/maptools_driver.cpp — TEST INFRASTRUCTURE, not product code.
//
// extern "C" driver over the UNMODIFIED hector_map_tools/HectorMapTools.h (included where it lies under
// /root/reference/hector_map_tools/include), compiled against oracle/shim (mini-Eigen + a plain-struct
// nav_msgs/OccupancyGrid).  Pins SURVEY.md §8f N4 on the reference itself:
//   hsref_maptools_raycast  -> DistanceMeasurementProvider::checkOccupancyBresenhami   HectorMapTools.h:148-214
//   hsref_maptools_get_dist -> DistanceMeasurementProvider::getDist                    HectorMapTools.h:133-147
#include <iostream>
#include <vector>
#include <climits>
#include <cstring>

#include <nav_msgs/OccupancyGrid.h>
#include "hector_map_tools/HectorMapTools.h"

namespace {
struct MapTools {
  std::shared_ptr<nav_msgs::OccupancyGrid> grid;
  HectorMapTools::DistanceMeasurementProvider dist;
};
}  // namespace

extern "C" {

// data: width*height occupancy values (0 / 100 / -1) as HectorMappingRos::publishMap writes them
void* hsref_maptools_create(int width, int height, float resolution, double origin_x, double origin_y, const int8_t* data) {
  MapTools* m = new MapTools();
  m->grid = std::make_shared<nav_msgs::OccupancyGrid>();
  m->grid->info.resolution = resolution;
  m->grid->info.width = (uint32_t)width;
  m->grid->info.height = (uint32_t)height;
  m->grid->info.origin.position.x = origin_x;
  m->grid->info.origin.position.y = origin_y;
  m->grid->info.origin.position.z = 0.0;
  m->grid->data.assign(data, data + (size_t)width * height);
  m->dist.setMap(m->grid);
  return m;
}

void hsref_maptools_destroy(void* p) { delete static_cast<MapTools*>(p); }

float hsref_maptools_raycast(void* p, int x0, int y0, int x1, int y1, int hit[2]) {
  MapTools* m = static_cast<MapTools*>(p);
  Eigen::Vector2i h(-1, -1);
  float d = m->dist.checkOccupancyBresenhami(Eigen::Vector2i(x0, y0), Eigen::Vector2i(x1, y1), &h);
  hit[0] = h[0];
  hit[1] = h[1];
  return d;
}

// returns getDist's value; *found = 0 when nothing was hit (the reference then leaves hitCoords built from an
// uninitialised vector, HectorMapTools.h:135,141 — not reported here)
float hsref_maptools_get_dist(void* p, const float begin_world[2], const float end_world[2], float hit_world[2], int* found) {
  MapTools* m = static_cast<MapTools*>(p);
  const float raw = m->dist.getDist(Eigen::Vector2f(begin_world[0], begin_world[1]), Eigen::Vector2f(end_world[0], end_world[1]));
  *found = raw >= 0.0f;
  hit_world[0] = hit_world[1] = 0.0f;
  if (*found) {
    Eigen::Vector2f hw;
    m->dist.getDist(Eigen::Vector2f(begin_world[0], begin_world[1]), Eigen::Vector2f(end_world[0], end_world[1]), &hw);
    hit_world[0] = hw[0];
    hit_world[1] = hw[1];
  }
  return raw;
}

}  // extern "C"
