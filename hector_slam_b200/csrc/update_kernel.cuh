// update_kernel.cuh — K2 (Bresenham log-odds map update) and K3 (probability-plane refresh), sm_100a.
//
// What K2 replaces (paths under /root/reference/hector_mapping/include/hector_slam_lib/):
//   OccGridMapBase::updateByScan          map/OccGridMapBase.h:121-168
//   OccGridMapBase::updateLineBresenhami  map/OccGridMapBase.h:170-214
//   OccGridMapBase::bresenham2D           map/OccGridMapBase.h:243-260
//   bresenhamCellFree / bresenhamCellOcc  map/OccGridMapBase.h:216-241
//   GridMapLogOddsFunctions::updateSetOccupied/SetFree/UnsetFree   map/GridMapLogOdds.h:135-156
// and K3:
//   GridMapLogOddsFunctions::getGridProbability   map/GridMapLogOdds.h:163-166
//   (the reference memoises it per cell in GridMapCacheArray, map/GridMapCacheArray.h:80-102,
//    invalidated by onMapUpdated; here the probability plane is simply kept current)
//
// The reference walks beams one after another and uses per-cell `updateIndex` stamps so that a
// cell is touched once per scan, an end point (occupied) overriding any free marking.  That is
// order-free, so the device does it in two grid-wide phases with one warp per beam and lanes
// striding along the line (cell i of a Bresenham line has a closed form, no serial walk):
//   MARK : atomicMax(stamp[cell], base+1) along the line, atomicMax(stamp[end], base+2)
//   APPLY: the first thread to raise a marked cell's stamp to base+3 owns it and does the
//          log-odds update (free: l += lf ; occupied: if (l < 50) l += lo) and rewrites the
//          cell's probability (and the texture twin through a surface store).
// A cell that the reference first frees and then hits ends as ((l + lf) - lf) + lo there and as
// l + lo here — equal to ~1 ulp; tests compare planes with abs tol 1e-5.
#ifndef HSB_UPDATE_KERNEL_CUH
#define HSB_UPDATE_KERNEL_CUH

#include "hsb_internal.h"
#include "sincosf_glibc.h"

namespace hsb {

// P = e^l / (e^l + 1) in fp32 like the reference (expf, then one division).  exp is evaluated in
// double and rounded once, which reproduces a correctly rounded expf (glibc's is, to 0.502 ulp).
__device__ __forceinline__ float prob_from_logodds(float l) {
  const float odds = (float)exp((double)l);
  return __fdiv_rn(odds, __fadd_rn(odds, 1.0f));
}

__global__ void refresh_prob_kernel(const float* __restrict__ logodds, float* __restrict__ prob, cudaSurfaceObject_t surf,
                                    int sx, int sy) {
  const size_t n = (size_t)sx * (size_t)sy;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float p = prob_from_logodds(logodds[i]);
    prob[i] = p;
    if (surf) {
      const int y = (int)(i / (size_t)sx), x = (int)(i - (size_t)y * sx);
      surf2Dwrite(p, surf, x * (int)sizeof(float), y);
    }
  }
}

// LogOddsCell::resetGridCell for a whole level (GridMapBase.h:71-82): l = 0, P(0) = 0.5, stamps 0.
__global__ void clear_level_kernel(float* __restrict__ logodds, float* __restrict__ prob, uint32_t* __restrict__ stamp,
                                   cudaSurfaceObject_t surf, int sx, int sy) {
  const size_t n = (size_t)sx * (size_t)sy;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    logodds[i] = 0.0f;
    prob[i] = 0.5f;
    stamp[i] = 0u;
    if (surf) {
      const int y = (int)(i / (size_t)sx), x = (int)(i - (size_t)y * sx);
      surf2Dwrite(0.5f, surf, x * (int)sizeof(float), y);
    }
  }
}

// N1: nav_msgs/OccupancyGrid values as HectorMappingRos::publishMap derives them
// (hector_mapping/src/HectorMappingRos.cpp:448-468; isFree / isOccupied GridMapLogOdds.h:76-84).
__global__ void occupancy_kernel(const float* __restrict__ logodds, int8_t* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float l = logodds[i];
    out[i] = l < 0.0f ? (int8_t)0 : (l > 0.0f ? (int8_t)100 : (int8_t)-1);
  }
}

// N4: DistanceMeasurementProvider::checkOccupancyBresenhami (HectorMapTools.h:133-216), one warp per ray.
// Same closed-form line as K2; 32 cells are tested per step and the first occupied one wins.
__global__ void __launch_bounds__(256)
    raycast_kernel(const float* __restrict__ logodds, int sx, int sy, int B, const int2* __restrict__ begin,
                   const int2* __restrict__ end, float* __restrict__ out_dist, int2* __restrict__ out_hit) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b < B; b += nwarps) {
    const int2 p0 = begin[b], p1 = end[b];
    float dist = -1.0f;
    int2 hit = make_int2(-1, -1);
    const bool ok = p0.x >= 0 && p0.x < sx && p0.y >= 0 && p0.y < sy && p1.x >= 0 && p1.x < sx && p1.y >= 0 && p1.y < sy;  // :141-154
    if (ok) {
      const int dx = p1.x - p0.x, dy = p1.y - p0.y;
      const unsigned adx = (unsigned)abs(dx), ady = (unsigned)abs(dy);
      const int off_dx = dx > 0 ? 1 : -1, off_dy = (dy > 0 ? 1 : -1) * sx;   // :162-163
      unsigned ada, adb;
      int off_a, off_b;
      if (adx >= ady) { ada = adx; adb = ady; off_a = off_dx; off_b = off_dy; }   // :170-177
      else            { ada = ady; adb = adx; off_a = off_dy; off_b = off_dx; }
      const unsigned err0 = ada / 2u;
      const unsigned start = (unsigned)p0.y * (unsigned)sx + (unsigned)p0.x;
      const unsigned steps = min(5000u, ada);                                    // :203
      for (unsigned base = 0; base < steps; base += 32u) {
        const unsigned i = base + (unsigned)lane;
        unsigned off = 0;
        bool occ = false;
        if (i < steps) {
          const unsigned carries = (unsigned)(((unsigned long long)err0 + (unsigned long long)i * adb) / ada);
          off = start + (unsigned)((int)i * off_a) + (unsigned)((int)carries * off_b);
          occ = logodds[off] > 0.0f;                                             // data[offset] == 100  (:208)
        }
        const unsigned m = __ballot_sync(0xffffffffu, occ);
        if (m) {
          const int first = __ffs(m) - 1;
          const unsigned hoff = __shfl_sync(0xffffffffu, off, first);
          hit = make_int2((int)(hoff % (unsigned)sx), (int)(hoff / (unsigned)sx));   // :182
          const float fx = (float)(p0.x - hit.x), fy = (float)(p0.y - hit.y);
          dist = (float)(int)__fsqrt_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));  // int distMap = norm  (:184)
          break;
        }
      }
    }
    if (lane == 0) {
      out_dist[b] = dist;
      if (out_hit) out_hit[b] = hit;
    }
  }
}

__device__ __forceinline__ void claim_and_apply(const HsbUpdateLevelDev& L, unsigned off, float lf, float lo) {
  const uint32_t base = L.stamp_base;
  const uint32_t v = __ldcg(L.stamp + off);
  if (v - (base + 1u) < 2u) {                       // marked for this scan, not yet applied
    const uint32_t old = atomicMax(L.stamp + off, base + 3u);
    if (old - (base + 1u) < 2u) {                   // we are the owner
      float l = L.logodds[off];
      if (old == base + 1u) {
        l = __fadd_rn(l, lf);                       // updateSetFree, GridMapLogOdds.h:146-151
      } else if (l < 50.0f) {
        l = __fadd_rn(l, lo);                       // updateSetOccupied, GridMapLogOdds.h:135-140
      }
      L.logodds[off] = l;
      const float p = prob_from_logodds(l);
      L.prob[off] = p;
      if (L.surf) {
        const int y = (int)(off / (unsigned)L.sx), x = (int)(off - (unsigned)y * (unsigned)L.sx);
        surf2Dwrite(p, L.surf, x * (int)sizeof(float), y);
      }
    }
  }
}

// One warp per beam; blockIdx.y = level.  APPLY = false: mark phase, true: apply phase.
template <bool APPLY>
__global__ void __launch_bounds__(256) update_kernel(const __grid_constant__ HsbUpdateParams P) {
  const HsbUpdateLevelDev& L = P.lv[blockIdx.y];
  if (!L.active) return;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int warp0 = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int warp_stride = gridDim.x * warps_per_block;

  // pose in this level's cells, transform Translation(x,y)*Rotation(psi)   (OccGridMapBase.h:127-131)
  float mx, my;
  {
    const float* m = L.mtw;
    mx = __fadd_rn(__fmul_rn(m[0], P.pose_world[0]), __fadd_rn(__fmul_rn(m[1], P.pose_world[1]), m[2]));
    my = __fadd_rn(__fmul_rn(m[3], P.pose_world[0]), __fadd_rn(__fmul_rn(m[4], P.pose_world[1]), m[5]));
  }
  const float c = cosf_glibc(P.pose_world[2]), s = sinf_glibc(P.pose_world[2]);  // Rotation2Df, as glibc
  // beam start = (int)(T * origo + 0.5)                                   (:134-137)
  const float ox = L.origo_x * L.pt_scale, oy = L.origo_y * L.pt_scale;
  const float bxf = __fadd_rn(__fmul_rn(c, ox), __fadd_rn(__fmul_rn(-s, oy), mx));
  const float byf = __fadd_rn(__fmul_rn(s, ox), __fadd_rn(__fmul_rn(c, oy), my));
  const float bxh = __fadd_rn(bxf, 0.5f), byh = __fadd_rn(byf, 0.5f);
  if (!(fabsf(bxh) < 1.0e9f) || !(fabsf(byh) < 1.0e9f)) return;  // non-finite / absurd pose: every beam is dropped
  const int x0 = (int)bxh, y0 = (int)byh;
  if ((x0 < 0) || (x0 >= L.sx) || (y0 < 0) || (y0 >= L.sy)) return;  // :176 (same start for all beams)
  const unsigned start = (unsigned)y0 * (unsigned)L.sx + (unsigned)x0;
  const uint32_t free_s = L.stamp_base + 1u, occ_s = L.stamp_base + 2u;
  // dirty rectangle (for tile replication to map replicas): every written cell lies on a segment
  // between the start cell and an end cell, so the box of those end points covers them all
  int bx0 = x0, by0 = y0, bx1 = x0, by1 = y0;
  bool wrote = false;

  for (int b = warp0; b < L.n; b += warp_stride) {
    const float2 pt = L.pts[b];
    const float px = pt.x * L.pt_scale, py = pt.y * L.pt_scale;   // DataPointContainer.h:46-58 setFrom
    float exf = __fadd_rn(__fmul_rn(c, px), __fadd_rn(__fmul_rn(-s, py), mx));   // :148
    float eyf = __fadd_rn(__fmul_rn(s, px), __fadd_rn(__fmul_rn(c, py), my));
    exf = __fadd_rn(exf, 0.5f);                                                  // :152
    eyf = __fadd_rn(eyf, 0.5f);
    if (!(fabsf(exf) < 1.0e9f) || !(fabsf(eyf) < 1.0e9f)) continue;
    const int x1 = (int)exf, y1 = (int)eyf;                                      // :155
    if (x1 == x0 && y1 == y0) continue;                                          // :158
    if ((x1 < 0) || (x1 >= L.sx) || (y1 < 0) || (y1 >= L.sy)) continue;          // :186
    const int dx = x1 - x0, dy = y1 - y0;
    const unsigned adx = (unsigned)abs(dx), ady = (unsigned)abs(dy);
    const int off_dx = dx > 0 ? 1 : -1;                         // util::sign, UtilFunctions.h:56-59
    const int off_dy = (dy > 0 ? 1 : -1) * L.sx;
    unsigned ada, adb;
    int off_a, off_b;
    if (adx >= ady) { ada = adx; adb = ady; off_a = off_dx; off_b = off_dy; }   // :202-209
    else            { ada = ady; adb = adx; off_a = off_dy; off_b = off_dx; }
    const unsigned err0 = ada / 2u;
    // cell i (0 <= i < ada; start included, end excluded, :245-259): the serial walk adds adb per
    // step and carries when the error reaches ada, so after i steps it has carried
    // floor((err0 + i*adb) / ada) times.
    for (unsigned i = (unsigned)lane; i < ada; i += 32u) {
      const unsigned carries = (unsigned)(((unsigned long long)err0 + (unsigned long long)i * adb) / ada);
      const unsigned off = start + (unsigned)((int)i * off_a) + (unsigned)((int)carries * off_b);
      if (APPLY) {
        claim_and_apply(L, off, P.log_odds_free, P.log_odds_occ);
      } else {
        if (__ldcg(L.stamp + off) < free_s) atomicMax(L.stamp + off, free_s);   // bresenhamCellFree :216-224
      }
    }
    if (!APPLY) {
      bx0 = min(bx0, x1); by0 = min(by0, y1); bx1 = max(bx1, x1); by1 = max(by1, y1);
      wrote = true;
    }
    if (lane == 0) {
      const unsigned end = (unsigned)y1 * (unsigned)L.sx + (unsigned)x1;        // :211-212
      if (APPLY) {
        claim_and_apply(L, end, P.log_odds_free, P.log_odds_occ);
      } else {
        atomicMax(L.stamp + end, occ_s);                                        // bresenhamCellOcc :226-241
      }
    }
  }
  if (!APPLY && wrote && lane == 0 && L.dirty) {  // one set of atomics per warp
    atomicMin(L.dirty + 0, bx0);
    atomicMin(L.dirty + 1, by0);
    atomicMax(L.dirty + 2, bx1);
    atomicMax(L.dirty + 3, by1);
  }
}

// Dirty-rectangle transport: pack the log-odds of rect = {x0, y0, x1, y1} (inclusive) row by row
// into a contiguous buffer / write such a buffer back and refresh P (and the texture twin) there.
__global__ void pack_rect_kernel(const float* __restrict__ logodds, int sx, int x0, int y0, int w, int hgt,
                                 float* __restrict__ buf) {
  const size_t n = (size_t)w * hgt;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (size_t)w), c = (int)(i - (size_t)r * w);
    buf[i] = logodds[(size_t)(y0 + r) * sx + (x0 + c)];
  }
}
__global__ void unpack_rect_kernel(float* __restrict__ logodds, float* __restrict__ prob, cudaSurfaceObject_t surf, int sx,
                                   int x0, int y0, int w, int hgt, const float* __restrict__ buf) {
  const size_t n = (size_t)w * hgt;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (size_t)w), c = (int)(i - (size_t)r * w);
    const size_t off = (size_t)(y0 + r) * sx + (x0 + c);
    const float l = buf[i];
    logodds[off] = l;
    const float p = prob_from_logodds(l);
    prob[off] = p;
    if (surf) surf2Dwrite(p, surf, (x0 + c) * (int)sizeof(float), y0 + r);
  }
}

}  // namespace hsb
#endif
