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
// order-free, so the device does it in two grid-wide phases (cell i of a Bresenham line has a
// closed form, no serial walk):
//   MARK : a team of warps per beam, lanes striding along the line:
//          atomicMax(stamp[cell], base+1) along the line, atomicMax(stamp[end], base+2)
//   APPLY: a coalesced sweep over the bounding box of the scan's beams; every cell whose stamp is
//          base+1 / base+2 gets the log-odds update (free: l += lf ; occupied: if (l < 50) l += lo)
//          and its probability rewritten (and the texture twin through a surface store).
// A cell that the reference first frees and then hits ends as ((l + lf) - lf) + lo there and as
// l + lo here — equal to ~1 ulp; tests compare planes with abs tol 1e-5.
#ifndef HSB_UPDATE_KERNEL_CUH
#define HSB_UPDATE_KERNEL_CUH

#include <climits>

#include "hsb_internal.h"
#include "sincosf_glibc.h"

namespace hsb {

// P = e^l / (e^l + 1) in fp32 like the reference (GridMapLogOdds.h:165-166: expf, one addition, one division).  expf is
// glibc's, operation for operation (sincosf_glibc.h: expf_glibc) — a correctly rounded exp differs from it by one ulp on
// ~1 % of the arguments — so the probability plane equals the reference's bit for bit.
__device__ __forceinline__ float prob_from_logodds(float l) {
  const float odds = expf_glibc(l);
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

// N1 on a rectangle: the cells of {x0, y0, w, hgt} thresholded into a packed w x hgt buffer (dirty-rectangle publishing).
__global__ void occupancy_rect_kernel(const float* __restrict__ logodds, int sx, int x0, int y0, int w, int hgt,
                                      int8_t* __restrict__ out) {
  const size_t n = (size_t)w * hgt;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (size_t)w), c = (int)(i - (size_t)r * w);
    const float l = logodds[(size_t)(y0 + r) * sx + (x0 + c)];
    out[i] = l < 0.0f ? (int8_t)0 : (l > 0.0f ? (int8_t)100 : (int8_t)-1);
  }
}

// N4: DistanceMeasurementProvider::checkOccupancyBresenhami (HectorMapTools.h:148-214, bresenham2D :216-237), one warp
// per ray.  Same closed-form line as K2; 32 cells are tested per step and the first occupied one wins.
__device__ __forceinline__ float warp_raycast(const float* __restrict__ logodds, int sx, int sy, int2 p0, int2 p1, int lane,
                                              int2& hit) {
  float dist = -1.0f;
  hit = make_int2(-1, -1);
  const bool ok = p0.x >= 0 && p0.x < sx && p0.y >= 0 && p0.y < sy && p1.x >= 0 && p1.x < sx && p1.y >= 0 && p1.y < sy;  // :155-166
  if (ok) {
    const int dx = p1.x - p0.x, dy = p1.y - p0.y;
    const unsigned adx = (unsigned)abs(dx), ady = (unsigned)abs(dy);
    const int off_dx = dx > 0 ? 1 : -1, off_dy = (dy > 0 ? 1 : -1) * sx;   // :174-175
    unsigned ada, adb;
    int off_a, off_b;
    if (adx >= ady) { ada = adx; adb = ady; off_a = off_dx; off_b = off_dy; }   // :182-189
    else            { ada = ady; adb = adx; off_a = off_dy; off_b = off_dx; }
    const unsigned err0 = ada / 2u;
    const unsigned start = (unsigned)p0.y * (unsigned)sx + (unsigned)p0.x;
    const unsigned steps = min(5000u, ada);                                    // :218
    for (unsigned base = 0; base < steps; base += 32u) {
      const unsigned i = base + (unsigned)lane;
      unsigned off = 0;
      bool occ = false;
      if (i < steps) {
        const unsigned carries = (unsigned)(((unsigned long long)err0 + (unsigned long long)i * adb) / ada);
        off = start + (unsigned)((int)i * off_a) + (unsigned)((int)carries * off_b);
        occ = logodds[off] > 0.0f;                                             // data[offset] == 100  (:223)
      }
      const unsigned m = __ballot_sync(0xffffffffu, occ);
      if (m) {
        const int first = __ffs(m) - 1;
        const unsigned hoff = __shfl_sync(0xffffffffu, off, first);
        hit = make_int2((int)(hoff % (unsigned)sx), (int)(hoff / (unsigned)sx));   // :194
        const float fx = (float)(p0.x - hit.x), fy = (float)(p0.y - hit.y);
        dist = (float)(int)__fsqrt_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));  // int distMap = norm  (:196)
        break;
      }
    }
  }
  return dist;
}

__global__ void __launch_bounds__(256)
    raycast_kernel(const float* __restrict__ logodds, int sx, int sy, int B, const int2* __restrict__ begin,
                   const int2* __restrict__ end, float* __restrict__ out_dist, int2* __restrict__ out_hit) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b < B; b += nwarps) {
    int2 hit;
    const float dist = warp_raycast(logodds, sx, sy, begin[b], end[b], lane, hit);
    if (lane == 0) {
      out_dist[b] = dist;
      if (out_hit) out_hit[b] = hit;
    }
  }
}

// N4: DistanceMeasurementProvider::getDist (HectorMapTools.h:133-147) for B world-frame rays: world -> cell with
// CoordinateTransformer<float>::getC2Coords ((w - origo) * inv_scale, :93-96) truncated by cast<int> (:136-137; a
// non-finite or absurd coordinate is treated as outside the map), the ray cast, the hit cell back to the world with
// getC1Coords (origo + cell * scale, :88-91) and the distance scaled by getC1Scale (scale * dist, :98-101; nothing hit:
// scale * -1 as in the reference, hit_world = (0, 0) and found = 0 where the reference leaves hitCoords undefined).
__global__ void __launch_bounds__(256)
    getdist_kernel(const float* __restrict__ logodds, int sx, int sy, int B, float origo_x, float origo_y, float scale,
                   float inv_scale, const float2* __restrict__ begin_world, const float2* __restrict__ end_world,
                   float* __restrict__ out_dist, float2* __restrict__ out_hit_world, int* __restrict__ out_found) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b < B; b += nwarps) {
    const float2 bw = begin_world[b], ew = end_world[b];
    const float bx = __fmul_rn(__fsub_rn(bw.x, origo_x), inv_scale), by = __fmul_rn(__fsub_rn(bw.y, origo_y), inv_scale);
    const float ex = __fmul_rn(__fsub_rn(ew.x, origo_x), inv_scale), ey = __fmul_rn(__fsub_rn(ew.y, origo_y), inv_scale);
    const bool sane = fabsf(bx) < 1.0e9f && fabsf(by) < 1.0e9f && fabsf(ex) < 1.0e9f && fabsf(ey) < 1.0e9f;
    const int2 p0 = sane ? make_int2((int)bx, (int)by) : make_int2(-1, -1);
    const int2 p1 = sane ? make_int2((int)ex, (int)ey) : make_int2(-1, -1);
    int2 hit;
    const float dist = warp_raycast(logodds, sx, sy, p0, p1, lane, hit);
    if (lane == 0) {
      out_dist[b] = __fmul_rn(scale, dist);
      const bool found = dist >= 0.0f;
      if (out_found) out_found[b] = found ? 1 : 0;
      if (out_hit_world)
        out_hit_world[b] = found ? make_float2(__fadd_rn(origo_x, __fmul_rn((float)hit.x, scale)),
                                               __fadd_rn(origo_y, __fmul_rn((float)hit.y, scale)))
                                 : make_float2(0.0f, 0.0f);
    }
  }
}

__device__ __forceinline__ void pdl_launch_dependents_early() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- K2 -----------------------------------------------------------------------------------------

// HectorSlamProcessor::update's gate (slam_main/HectorSlamProcessor.h:83-95) evaluated on the device so that a
// fused step needs no host round trip between match and map write.
//   state: [0..2] lastMapUpdatePose, [3] out: 1.0f if the map is to be written by this step
//   in   : [0] minDist, [1] minAngle, [2] force (map_without_matching)
// pose_out (device) and pose_out_host (mapped host memory, may be null) receive the step's pose (the matched
// pose, or the hint when matching is skipped) and, in [3], a copy of the flag.
__global__ void slam_gate_kernel(float* __restrict__ state, const float* __restrict__ in, const float* pose_in,
                                 float* pose_out, float* pose_out_host, unsigned* seq_host, unsigned seq_value) {
  pdl_launch_dependents_early();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    slam_gate(state, in, pose_in[0], pose_in[1], pose_in[2], pose_out, pose_out_host);
    if (seq_host) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned*>(seq_host) = seq_value;
    }
  }
}

// map_without_matching step from a point cloud: the kept-endpoint count goes to mapped host memory, then the sequence number
__global__ void publish_int_kernel(const int* __restrict__ value_dev, int* value_host, unsigned* seq_host, unsigned seq_value) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    *value_host = *value_dev;
    if (seq_host) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned*>(seq_host) = seq_value;
    }
  }
}

// Per-level frame of one updateByScan call: the pose as Translation(x,y)*Rotation(psi) in this level's cells and
// the common start cell of all beams (OccGridMapBase.h:127-137).
struct BeamFrame {
  float c, s, mx, my;
  int x0, y0;
  bool ok;
};
__device__ __forceinline__ BeamFrame beam_frame(const HsbUpdateLevelDev& L, float wx, float wy, float wpsi) {
  BeamFrame f;
  const float* m = L.mtw;
  f.mx = __fadd_rn(__fmul_rn(m[0], wx), __fadd_rn(__fmul_rn(m[1], wy), m[2]));
  f.my = __fadd_rn(__fmul_rn(m[3], wx), __fadd_rn(__fmul_rn(m[4], wy), m[5]));
  f.c = cosf_glibc(wpsi);   // Rotation2Df, as glibc
  f.s = sinf_glibc(wpsi);
  // beam start = (int)(T * origo + 0.5)                                   (:134-137)
  const float ox = L.origo_x * L.pt_scale, oy = L.origo_y * L.pt_scale;
  const float bxf = __fadd_rn(__fmul_rn(f.c, ox), __fadd_rn(__fmul_rn(-f.s, oy), f.mx));
  const float byf = __fadd_rn(__fmul_rn(f.s, ox), __fadd_rn(__fmul_rn(f.c, oy), f.my));
  const float bxh = __fadd_rn(bxf, 0.5f), byh = __fadd_rn(byf, 0.5f);
  f.x0 = 0; f.y0 = 0;
  f.ok = (fabsf(bxh) < 1.0e9f) && (fabsf(byh) < 1.0e9f);   // non-finite / absurd pose: every beam is dropped
  if (f.ok) {
    f.x0 = (int)bxh; f.y0 = (int)byh;
    f.ok = !((f.x0 < 0) || (f.x0 >= L.sx) || (f.y0 < 0) || (f.y0 >= L.sy));   // :176 (same start for all beams)
  }
  return f;
}
// End cell of beam b; false when the reference drops the beam.
__device__ __forceinline__ bool beam_end(const HsbUpdateLevelDev& L, const BeamFrame& f, int b, int& x1, int& y1) {
  const float2 pt = L.pts[b];
  const float px = pt.x * L.pt_scale, py = pt.y * L.pt_scale;   // DataPointContainer.h:46-58 setFrom
  float exf = __fadd_rn(__fmul_rn(f.c, px), __fadd_rn(__fmul_rn(-f.s, py), f.mx));   // :148
  float eyf = __fadd_rn(__fmul_rn(f.s, px), __fadd_rn(__fmul_rn(f.c, py), f.my));
  exf = __fadd_rn(exf, 0.5f);                                                        // :152
  eyf = __fadd_rn(eyf, 0.5f);
  if (!(fabsf(exf) < 1.0e9f) || !(fabsf(eyf) < 1.0e9f)) return false;
  x1 = (int)exf; y1 = (int)eyf;                                                      // :155
  if (x1 == f.x0 && y1 == f.y0) return false;                                        // :158
  if ((x1 < 0) || (x1 >= L.sx) || (y1 < 0) || (y1 >= L.sy)) return false;            // :186
  return true;
}

// MARK: a team of TEAM warps per beam; blockIdx.y = level.  Team lanes stride along the line, four cells per
// lane in flight (the loop is bound by the L2 round trip of the stamp test, not by arithmetic); the stamps are raised
// with result-less atomics (RED.MAX: nothing waits for them).  The bounding box of the start cell and the beams' end
// cells goes to the scan's scratch slot (4 atomics per warp): the apply phase sweeps exactly that box and the dirty
// rectangles are fed from it.  (A variant that also compacted the first-marked cells into a list for the apply phase
// was measured in round 2: it needs the atomics' return values — two more dependent L2 round trips per iteration —
// and made mark + apply slower, 29 us against 23 us; at a scan's size the box sweep is launch- and latency-sized,
// not bandwidth-sized.)
// Programmatic dependent launch (sm_90+): mark and apply are launched with programmatic stream serialisation, i.e.
// their CTAs may become resident while the previous kernel of the stream (the match kernel of a fused SLAM step, or
// mark before apply) is still running; griddepcontrol.wait blocks until that kernel has completed and its writes are
// visible.  Everything a kernel reads from its predecessor (pose, gate flag, stamps, lists) comes after the wait.
// This takes the launch latency (2-3 us per dependent launch) off the critical path of the single-scan step.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <int TEAM>
__global__ void __launch_bounds__(256) update_mark_kernel(const __grid_constant__ HsbUpdateParams P) {
  pdl_launch_dependents();   // the apply kernel's CTAs may take the free slots now; they wait for this grid to finish
  pdl_wait();
  const HsbUpdateLevelDev& L = P.lv[blockIdx.y];
  if (!L.active) return;
  // gate flag, pose and beam count in ONE L2 round trip (independent loads, issued before the first dependent branch)
  // (plain loads: every thread reads these words, they have to hit L1)
  const float flag = P.gate_flag ? *P.gate_flag : 1.0f;
  const float* pw = P.pose_dev ? P.pose_dev : P.pose_world;
  const float wx = pw[0], wy = pw[1], wpsi = pw[2];
  const int n_dev = L.n_dev ? *L.n_dev : L.n;
  if (flag == 0.0f) return;
  constexpr int TL = 32 * TEAM;   // lanes per team
  constexpr int U = 4;
  const unsigned tlane = threadIdx.x % TL;
  const int team0 = (blockIdx.x * blockDim.x + threadIdx.x) / TL;
  const int team_stride = (gridDim.x * blockDim.x) / TL;
  const BeamFrame f = beam_frame(L, wx, wy, wpsi);
  if (!f.ok) return;
  const unsigned start = (unsigned)f.y0 * (unsigned)L.sx + (unsigned)f.x0;
  const uint32_t free_s = L.stamp_base + 1u, occ_s = L.stamp_base + 2u;
  int* slot = L.scratch + 8 * L.slot;
  int bx0 = INT_MAX, by0 = INT_MAX, bx1 = -1, by1 = -1;

  const int n_beams = min(n_dev, L.n);
  for (int b = team0; b < n_beams; b += team_stride) {
    int x1, y1;
    if (!beam_end(L, f, b, x1, y1)) continue;
    bx0 = min(bx0, x1); by0 = min(by0, y1); bx1 = max(bx1, x1); by1 = max(by1, y1);
    const int dx = x1 - f.x0, dy = y1 - f.y0;
    const unsigned adx = (unsigned)abs(dx), ady = (unsigned)abs(dy);
    const int off_dx = dx > 0 ? 1 : -1;                         // util::sign, UtilFunctions.h:56-59
    const int off_dy = (dy > 0 ? 1 : -1) * L.sx;
    unsigned ada, adb;
    int off_a, off_b;
    if (adx >= ady) { ada = adx; adb = ady; off_a = off_dx; off_b = off_dy; }   // :202-209
    else            { ada = ady; adb = adx; off_a = off_dy; off_b = off_dx; }
    const unsigned err0 = ada / 2u;
    // cell i (0 <= i < ada; start included, end excluded, :245-259): the serial walk adds adb per step and carries when
    // the error reaches ada, so after i steps it has carried q(i) = floor((err0 + i*adb) / ada) times.  A lane visits
    // i = tlane, tlane + TL, ...: q and the remainder r are advanced by the constant step (TL*adb = qT*ada + rT) —
    // two divisions per lane and beam instead of one per cell.
    unsigned q, r, qT, rT;
    if (ada < 32768u) {   // err0 + i*adb and TL*adb fit 32 bits
      const unsigned n0 = err0 + tlane * adb, nT = (unsigned)TL * adb;
      q = n0 / ada; r = n0 - q * ada;
      qT = nT / ada; rT = nT - qT * ada;
    } else {
      const unsigned long long n0 = (unsigned long long)err0 + (unsigned long long)tlane * adb, nT = (unsigned long long)TL * adb;
      q = (unsigned)(n0 / ada); r = (unsigned)(n0 - (unsigned long long)q * ada);
      qT = (unsigned)(nT / ada); rT = (unsigned)(nT - (unsigned long long)qT * ada);
    }
    for (unsigned i0 = tlane; i0 < ada; i0 += U * TL) {
      unsigned off[U];
      uint32_t v[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const unsigned i = i0 + (unsigned)(k * TL);
        v[k] = 0xffffffffu;
        if (i < ada) {
          off[k] = start + (unsigned)((int)i * off_a) + (unsigned)((int)q * off_b);
          v[k] = __ldcg(L.stamp + off[k]);
        }
        q += qT;
        r += rT;
        if (r >= ada) { r -= ada; ++q; }
      }
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (v[k] < free_s) atomicMax(L.stamp + off[k], free_s);   // bresenhamCellFree :216-224
    }
    if (tlane == 0) atomicMax(L.stamp + (unsigned)y1 * (unsigned)L.sx + (unsigned)x1, occ_s);   // bresenhamCellOcc :226-241, :211-212
  }
  // bounding box of this CTA's beams (+ the common start cell) -> scratch slot.  One reduction per CTA and only the
  // atomics that can still enlarge the box: with 4 atomics per WARP (6500 warps on the same four words) the drain of the
  // same-address atomics alone cost ~7 us of a 17 us kernel (profiles/r02_slam_step_launches.md).
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bx0 = min(bx0, __shfl_xor_sync(0xffffffffu, bx0, o));
    by0 = min(by0, __shfl_xor_sync(0xffffffffu, by0, o));
    bx1 = max(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
    by1 = max(by1, __shfl_xor_sync(0xffffffffu, by1, o));
  }
  __shared__ int box[4][8];
  const int wi = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { box[0][wi] = bx0; box[1][wi] = by0; box[2][wi] = bx1; box[3][wi] = by1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      bx0 = min(bx0, box[0][k]); by0 = min(by0, box[1][k]); bx1 = max(bx1, box[2][k]); by1 = max(by1, box[3][k]);
    }
    if (bx1 >= 0) {
      bx0 = min(bx0, f.x0); by0 = min(by0, f.y0); bx1 = max(bx1, f.x0); by1 = max(by1, f.y0);
      // the four current bounds in ONE L2 round trip (a stale value only costs an atomic that changes nothing; read one
      // by one through a volatile pointer they were four dependent round trips at the end of every CTA — a third of
      // this kernel's stall samples, profiles/r02_slam_step_ncu.md)
      const int c1 = __ldcg(slot + 1), c2 = __ldcg(slot + 2), c3 = __ldcg(slot + 3), c4 = __ldcg(slot + 4);
      if (bx0 < c1) atomicMin(slot + 1, bx0);
      if (by0 < c2) atomicMin(slot + 2, by0);
      if (bx1 > c3) atomicMax(slot + 3, bx1);
      if (by1 > c4) atomicMax(slot + 4, by1);
    }
  }
}

// APPLY: every marked cell lies in the bounding box of the start cell and the end cells, which the mark phase left in
// the scan's scratch slot, so the apply phase is a coalesced sweep of that box over the stamp plane (uint4 loads, four
// in flight per thread; a scalar variant for levels whose rows are not 16-byte multiples): stamp == base+1 -> l += lf,
// base+2 -> if (l < 50) l += lo, then P and the texture twin are rewritten.  One owner per cell by construction — no
// atomics, deterministic.
__device__ __forceinline__ void apply_cell(const HsbUpdateLevelDev& L, unsigned off, uint32_t v, float lf, float lo) {
  const uint32_t d = v - (L.stamp_base + 1u);
  if (d < 2u) {
    float l = L.logodds[off];
    if (d == 0u) {
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

__global__ void __launch_bounds__(256) update_apply_kernel(const __grid_constant__ HsbUpdateParams P) {
  pdl_wait();
  const HsbUpdateLevelDev& L = P.lv[blockIdx.y];
  if (!L.active) return;
  const int* cur = L.scratch + 8 * L.slot;
  if (blockIdx.x == 0 && threadIdx.x == 0) {   // the other slot is the next scan's: clear it (nobody reads it now)
    int* nxt = L.scratch + 8 * (L.slot ^ 1);
    nxt[0] = 0;
    nxt[1] = INT_MAX; nxt[2] = INT_MAX; nxt[3] = -1; nxt[4] = -1;
  }
  // gate flag and box in ONE round trip (independent loads, issued before the first dependent branch; plain loads:
  // every thread of the grid reads these words, they have to hit L1 — ld.cg here cost 2.5 us of L2 hot-spotting)
  const float flag = P.gate_flag ? *P.gate_flag : 1.0f;
  const int bx0 = cur[1], by0 = cur[2], bx1 = cur[3], by1 = cur[4];
  if (flag == 0.0f) return;
  if (bx1 < bx0) return;                       // nothing was marked (pose outside the map, every beam dropped)
  if (blockIdx.x == 0 && threadIdx.x == 0 && L.dirty) {   // both rectangles: replication [0..3] and host mirror [4..7]
    atomicMin(L.dirty + 0, bx0); atomicMin(L.dirty + 4, bx0);
    atomicMin(L.dirty + 1, by0); atomicMin(L.dirty + 5, by0);
    atomicMax(L.dirty + 2, bx1); atomicMax(L.dirty + 6, bx1);
    atomicMax(L.dirty + 3, by1); atomicMax(L.dirty + 7, by1);
  }
  const float lf = P.log_odds_free, lo = P.log_odds_occ;
  const unsigned nthreads = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned rows = (unsigned)(by1 - by0 + 1);
  constexpr int U = 4;
  if ((L.sx & 3) == 0) {   // rows are 16-byte aligned: four cells per load
    const unsigned xa = (unsigned)bx0 & ~3u;
    const unsigned w4 = (((unsigned)bx1 - xa) >> 2) + 1u;
    const unsigned total = w4 * rows;
    for (unsigned base = tid; base < total; base += U * nthreads) {
      unsigned off[U];
      uint4 v[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const unsigned idx = base + (unsigned)k * nthreads;
        v[k] = make_uint4(L.stamp_base, L.stamp_base, L.stamp_base, L.stamp_base);
        if (idx < total) {
          const unsigned r = idx / w4, c = idx - r * w4;
          off[k] = ((unsigned)by0 + r) * (unsigned)L.sx + xa + 4u * c;
          v[k] = __ldcg(reinterpret_cast<const uint4*>(L.stamp + off[k]));
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        apply_cell(L, off[k] + 0u, v[k].x, lf, lo);
        apply_cell(L, off[k] + 1u, v[k].y, lf, lo);
        apply_cell(L, off[k] + 2u, v[k].z, lf, lo);
        apply_cell(L, off[k] + 3u, v[k].w, lf, lo);
      }
    }
  } else {
    const unsigned wd = (unsigned)(bx1 - bx0 + 1);
    const unsigned total = wd * rows;
    for (unsigned idx = tid; idx < total; idx += nthreads) {
      const unsigned rr = idx / wd, c = idx - rr * wd;
      const unsigned off = ((unsigned)by0 + rr) * (unsigned)L.sx + (unsigned)bx0 + c;
      apply_cell(L, off, __ldcg(L.stamp + off), lf, lo);
    }
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
// (a replica's host mirror has to follow too: the rectangle is folded into the level's mirror rectangle)
__global__ void unpack_rect_kernel(float* __restrict__ logodds, float* __restrict__ prob, cudaSurfaceObject_t surf, int sx,
                                   int x0, int y0, int w, int hgt, const float* __restrict__ buf, int* mirror_dirty) {
  const size_t n = (size_t)w * hgt;
  if (mirror_dirty && blockIdx.x == 0 && threadIdx.x == 0) {
    atomicMin(mirror_dirty + 0, x0);
    atomicMin(mirror_dirty + 1, y0);
    atomicMax(mirror_dirty + 2, x0 + w - 1);
    atomicMax(mirror_dirty + 3, y0 + hgt - 1);
  }
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


// ---- one-shot tile transport: everything the host-synchronous protocol asked the host for stays on the device ----
// The owner's pack kernel reads the dirty rectangles of all levels from device memory, writes them into the buffer's
// header and the rectangles' log-odds rows behind it; the replicas' unpack kernel reads the header from the buffer the
// collective delivered.  No size ever travels through a host, so the whole replication step is stream-ordered: pack ->
// ONE fixed-size ncclBroadcast -> unpack.
struct TileLayout {
  int x0[HSB_MAX_LEVELS], y0[HSB_MAX_LEVELS], w[HSB_MAX_LEVELS], h[HSB_MAX_LEVELS];
  unsigned off[HSB_MAX_LEVELS + 1];   // cells before level l
};
__device__ __forceinline__ void tile_layout_from(const int* rects, int levels, TileLayout& t) {
  unsigned acc = 0;
  for (int l = 0; l < levels; ++l) {
    const int x0 = rects[4 * l], y0 = rects[4 * l + 1], x1 = rects[4 * l + 2], y1 = rects[4 * l + 3];
    const bool dirty = x1 >= x0 && y1 >= y0;
    t.x0[l] = x0; t.y0[l] = y0;
    t.w[l] = dirty ? x1 - x0 + 1 : 0;
    t.h[l] = dirty ? y1 - y0 + 1 : 0;
    t.off[l] = acc;
    acc += (unsigned)t.w[l] * (unsigned)t.h[l];
  }
  t.off[levels] = acc;
}

__global__ void __launch_bounds__(256) pack_dirty_kernel(const __grid_constant__ HsbTileParams P) {
  __shared__ int rects[4 * HSB_MAX_LEVELS];
  if (threadIdx.x < 4 * P.levels) rects[threadIdx.x] = P.lv[threadIdx.x >> 2].dirty[threadIdx.x & 3];
  __syncthreads();
  TileLayout t;
  tile_layout_from(rects, P.levels, t);
  const unsigned total = t.off[P.levels];
  const bool overflow = (unsigned long long)total + HSB_TILE_HEADER_WORDS > P.capacity_words;
  int* hdr = reinterpret_cast<int*>(P.buf);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    hdr[0] = HSB_TILE_MAGIC; hdr[1] = P.levels; hdr[2] = overflow ? 1 : 0; hdr[3] = (int)total;
  }
  if (blockIdx.x == 0 && threadIdx.x < 4 * P.levels) hdr[4 + threadIdx.x] = rects[threadIdx.x];
  if (overflow) return;
  float* out = P.buf + HSB_TILE_HEADER_WORDS;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int l = 0;
    while (i >= t.off[l + 1]) ++l;
    const unsigned k = i - t.off[l];
    const unsigned r = k / (unsigned)t.w[l], c = k - r * (unsigned)t.w[l];
    out[i] = P.lv[l].logodds[(size_t)(t.y0[l] + (int)r) * P.lv[l].sx + (t.x0[l] + (int)c)];
  }
}
// after pack (stream order): the shipped rectangles are forgotten — unless the buffer overflowed and nothing was shipped
__global__ void reset_dirty_kernel(const __grid_constant__ HsbTileParams P) {
  const int* hdr = reinterpret_cast<const int*>(P.buf);
  if (hdr[2] != 0) return;
  const int l = threadIdx.x >> 2, k = threadIdx.x & 3;
  if (l < P.levels) P.lv[l].dirty[k] = (k < 2) ? INT_MAX : -1;
}
__global__ void __launch_bounds__(256) unpack_dirty_kernel(const __grid_constant__ HsbTileParams P) {
  const int* hdr = reinterpret_cast<const int*>(P.buf);
  if (hdr[0] != HSB_TILE_MAGIC || hdr[1] != P.levels || hdr[2] != 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.error_count) atomicAdd(P.error_count, 1);
    return;
  }
  TileLayout t;
  tile_layout_from(hdr + 4, P.levels, t);
  const unsigned total = t.off[P.levels];
  if ((unsigned long long)total + HSB_TILE_HEADER_WORDS > P.capacity_words) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.error_count) atomicAdd(P.error_count, 1);
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x < P.levels && t.w[threadIdx.x] > 0) {   // the replica's host mirror has to follow
    const int l = threadIdx.x;
    int* md = P.lv[l].dirty + 4;
    atomicMin(md + 0, t.x0[l]); atomicMin(md + 1, t.y0[l]);
    atomicMax(md + 2, t.x0[l] + t.w[l] - 1); atomicMax(md + 3, t.y0[l] + t.h[l] - 1);
  }
  const float* in = P.buf + HSB_TILE_HEADER_WORDS;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int l = 0;
    while (i >= t.off[l + 1]) ++l;
    const unsigned k = i - t.off[l];
    const unsigned r = k / (unsigned)t.w[l], c = k - r * (unsigned)t.w[l];
    const int x = t.x0[l] + (int)c, y = t.y0[l] + (int)r;
    if (x < 0 || y < 0 || x >= P.lv[l].sx || y >= P.lv[l].sy) continue;   // a foreign geometry must not write out of bounds
    const size_t off = (size_t)y * P.lv[l].sx + x;
    const float v = in[i];
    P.lv[l].logodds[off] = v;
    const float p = prob_from_logodds(v);
    P.lv[l].prob[off] = p;
    if (P.lv[l].surf) surf2Dwrite(p, P.lv[l].surf, x * (int)sizeof(float), y);
  }
}

}  // namespace hsb
#endif
