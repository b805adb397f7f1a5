"""Seeded synthetic worlds and Hokuyo UTM-30LX scans (SURVEY.md §8d "synthetic world").

The reference ships no fixtures, bags or golden vectors, so tests and bench.py
generate their inputs here.  Nothing in this file is on the measured path: it produces host
arrays that are then handed to the C-ABI (and to the oracle).

World for a level-0 grid of S x S cells at 0.05 m: an R x R lattice of closed rectangular rooms,
R = S // 1024, each 20 m x 14 m with 8 circular pillars (radius U[0.3, 1.5] m) — i.e. overall
20R x 14R m and 8 R^2 pillars, centred on the world origin (= map centre for start coords 0.5).
Rooms are closed so that every beam returns inside 29.9 m and every scan keeps exactly 1081
valid endpoints (the node drops returns >= range_max - 0.1, HectorMappingRos.cpp:493-499).

Scan -> endpoints follows HectorMappingRos::rosLaserScanToDataContainer
(hector_mapping/src/HectorMappingRos.cpp:483-507) in fp32: the beam angle is accumulated
(`angle += angle_increment`), `dist *= scaleToMap`, point = (cos(angle)*dist, sin(angle)*dist),
origo = (0, 0).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

ROOM_W = 20.0
ROOM_D = 14.0
PILLARS_PER_ROOM = 8
N_BEAMS = 1081
ANGLE_MIN = np.float32(-135.0 * math.pi / 180.0)
ANGLE_INC = np.float32(0.25 * math.pi / 180.0)
RANGE_MIN = np.float32(0.1)
RANGE_MAX = np.float32(30.0)


def beam_angles(n: int = N_BEAMS) -> np.ndarray:
    """fp32 beam angles accumulated the way the node does (HectorMappingRos.cpp:487,505)."""
    out = np.empty(n, dtype=np.float32)
    a = np.float32(ANGLE_MIN)
    for i in range(n):
        out[i] = a
        a = np.float32(a + ANGLE_INC)
    return out


_ANGLES = beam_angles()


@dataclass
class World:
    """Lattice of closed rooms with pillars; analytic ray casting."""

    rooms_per_side: int
    seed: int = 1234

    def __post_init__(self):
        r = self.rooms_per_side
        rng = np.random.default_rng(self.seed)
        self.x0 = -0.5 * r * ROOM_W
        self.y0 = -0.5 * r * ROOM_D
        # pillars[i, j] -> (K, 3) array of (cx, cy, radius), world metres
        self.pillars = np.empty((r, r, PILLARS_PER_ROOM, 3), dtype=np.float64)
        for i in range(r):
            for j in range(r):
                rad = rng.uniform(0.3, 1.5, PILLARS_PER_ROOM)
                lo_x = self.x0 + i * ROOM_W
                lo_y = self.y0 + j * ROOM_D
                cx = lo_x + rad + 0.5 + rng.uniform(0, 1, PILLARS_PER_ROOM) * (ROOM_W - 2 * rad - 1.0)
                cy = lo_y + rad + 0.5 + rng.uniform(0, 1, PILLARS_PER_ROOM) * (ROOM_D - 2 * rad - 1.0)
                self.pillars[i, j, :, 0] = cx
                self.pillars[i, j, :, 1] = cy
                self.pillars[i, j, :, 2] = rad

    @classmethod
    def for_map_size(cls, map_size: int, seed: int = 1234) -> "World":
        return cls(max(1, map_size // 1024), seed)

    def room_index(self, x: float, y: float) -> tuple[int, int]:
        r = self.rooms_per_side
        i = min(r - 1, max(0, int((x - self.x0) // ROOM_W)))
        j = min(r - 1, max(0, int((y - self.y0) // ROOM_D)))
        return i, j

    def room_bounds(self, i: int, j: int) -> tuple[float, float, float, float]:
        lo_x = self.x0 + i * ROOM_W
        lo_y = self.y0 + j * ROOM_D
        return lo_x, lo_y, lo_x + ROOM_W, lo_y + ROOM_D

    def clearance(self, x: float, y: float) -> float:
        """Distance from (x, y) to the nearest wall or pillar surface of its room."""
        i, j = self.room_index(x, y)
        lo_x, lo_y, hi_x, hi_y = self.room_bounds(i, j)
        d = min(x - lo_x, hi_x - x, y - lo_y, hi_y - y)
        p = self.pillars[i, j]
        dp = np.hypot(p[:, 0] - x, p[:, 1] - y) - p[:, 2]
        return float(min(d, dp.min()))

    def cast(self, pose, angles: np.ndarray = _ANGLES) -> np.ndarray:
        """Exact ranges (float64) of beams leaving `pose` = (x, y, psi) at `angles` (robot frame)."""
        x, y, psi = float(pose[0]), float(pose[1]), float(pose[2])
        i, j = self.room_index(x, y)
        lo_x, lo_y, hi_x, hi_y = self.room_bounds(i, j)
        a = angles.astype(np.float64) + psi
        dx, dy = np.cos(a), np.sin(a)
        with np.errstate(divide="ignore", invalid="ignore"):
            tx = np.where(dx > 0, (hi_x - x) / dx, np.where(dx < 0, (lo_x - x) / dx, np.inf))
            ty = np.where(dy > 0, (hi_y - y) / dy, np.where(dy < 0, (lo_y - y) / dy, np.inf))
        t = np.minimum(tx, ty)
        for cx, cy, rad in self.pillars[i, j]:
            ox, oy = x - cx, y - cy
            b = ox * dx + oy * dy
            c = ox * ox + oy * oy - rad * rad
            disc = b * b - c
            hit = disc >= 0
            root = np.sqrt(np.where(hit, disc, 0.0))
            t0 = -b - root
            t = np.where(hit & (t0 > 1e-9), np.minimum(t, t0), t)
        return t

    def sample_free_poses(self, n: int, rng: np.random.Generator, margin: float = 0.6) -> np.ndarray:
        """n poses uniform over the free space of all rooms (clearance >= margin), psi U[-pi, pi)."""
        r = self.rooms_per_side
        out = np.empty((n, 3), dtype=np.float64)
        k = 0
        while k < n:
            x = self.x0 + rng.uniform(0, r * ROOM_W)
            y = self.y0 + rng.uniform(0, r * ROOM_D)
            psi = rng.uniform(-math.pi, math.pi)
            if self.clearance(x, y) >= margin:
                out[k] = (x, y, psi)
                k += 1
        return out

    def mapping_poses(self) -> np.ndarray:
        """Deterministic poses used to build a map of every room: a 4 x 3 lattice of positions per
        room (nudged off pillars), two opposite headings each (270 deg FOV -> full coverage)."""
        poses = []
        r = self.rooms_per_side
        for i in range(r):
            for j in range(r):
                lo_x, lo_y, hi_x, hi_y = self.room_bounds(i, j)
                for u in range(4):
                    for v in range(3):
                        x = lo_x + (u + 0.5) * ROOM_W / 4
                        y = lo_y + (v + 0.5) * ROOM_D / 3
                        # walk away from pillars deterministically until there is clearance
                        step = 0
                        while self.clearance(x, y) < 0.6 and step < 200:
                            x += 0.37 * math.cos(step * 2.399963)
                            y += 0.37 * math.sin(step * 2.399963)
                            x = min(max(x, lo_x + 0.7), hi_x - 0.7)
                            y = min(max(y, lo_y + 0.7), hi_y - 0.7)
                            step += 1
                        if self.clearance(x, y) < 0.6:
                            continue
                        for psi in (0.3, 0.3 + math.pi):
                            poses.append((x, y, math.atan2(math.sin(psi), math.cos(psi))))
        return np.asarray(poses, dtype=np.float64)


def ranges_to_points(ranges: np.ndarray, scale_to_map: float, angles: np.ndarray = _ANGLES) -> np.ndarray:
    """fp32 restatement of rosLaserScanToDataContainer (HectorMappingRos.cpp:483-507): keeps
    returns with range_min < r < range_max - 0.1, scales by scaleToMap, returns (n, 2) float32."""
    r = ranges.astype(np.float32)
    keep = (r > RANGE_MIN) & (r < np.float32(RANGE_MAX - np.float32(0.1)))
    d = (r[keep] * np.float32(scale_to_map)).astype(np.float32)
    a = angles[keep]
    pts = np.empty((d.shape[0], 2), dtype=np.float32)
    pts[:, 0] = (np.cos(a).astype(np.float32) * d).astype(np.float32)
    pts[:, 1] = (np.sin(a).astype(np.float32) * d).astype(np.float32)
    return pts


def make_scan(world: World, pose, rng: np.random.Generator | None, scale_to_map: float = 20.0,
              sigma: float = 0.01) -> np.ndarray:
    """One 1081-beam scan from `pose`; Gaussian range noise sigma (m) if rng is given."""
    r = world.cast(pose)
    if rng is not None and sigma > 0:
        r = r + rng.normal(0.0, sigma, r.shape)
    return ranges_to_points(r, scale_to_map)


def make_range_batch(world: World, poses: np.ndarray, noise_seed: int = 7, sigma: float = 0.01) -> np.ndarray:
    """Raw ranges (B, 1081) float32 — what sensor_msgs/LaserScan carries — same noise stream as
    make_scan_batch, so ranges_to_points(make_range_batch(..)[b]) == make_scan_batch(..) scan b."""
    rng = np.random.default_rng(noise_seed)
    out = np.empty((len(poses), N_BEAMS), dtype=np.float32)
    for b, p in enumerate(poses):
        r = world.cast(p)
        if sigma > 0:
            r = r + rng.normal(0.0, sigma, r.shape)
        out[b] = r.astype(np.float32)
    return out


def ranges_to_cloud(ranges: np.ndarray, angles: np.ndarray = _ANGLES, cutoff: float = 30.0) -> np.ndarray:
    """A sensor_msgs/PointCloud in the LASER frame from one scan's ranges, the way the node obtains its default input
    (laser_geometry projectLaser(scan, cloud, 30.0), HectorMappingRos.cpp:274): one Point32 (x, y, z = 0) per return
    with range_min <= r <= min(range_max, cutoff), beam order.  (n, 3) float32 — input generator, not product code."""
    r = ranges.astype(np.float32)
    keep = (r >= RANGE_MIN) & (r <= np.float32(min(float(RANGE_MAX), cutoff)))
    a = angles[keep].astype(np.float32)
    pts = np.zeros((int(keep.sum()), 3), dtype=np.float32)
    pts[:, 0] = (np.cos(a).astype(np.float32) * r[keep]).astype(np.float32)
    pts[:, 1] = (np.sin(a).astype(np.float32) * r[keep]).astype(np.float32)
    return pts


def laser_transform(xyz=(0.12, -0.03, 0.31), rpy=(0.02, -0.035, 0.05)) -> np.ndarray:
    """tf base_frame <- laser frame as 12 float64, rows of [R | t] (R = Rz(yaw) Ry(pitch) Rx(roll))."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    R = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                  [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                  [-sp, cp * sr, cp * cr]], dtype=np.float64)
    return np.ascontiguousarray(np.concatenate([R, np.asarray(xyz, np.float64)[:, None]], axis=1).reshape(12))


# HectorMappingRos.cpp:98-108 defaults: laser_min_dist 0.4 m, laser_max_dist 30 m (squared, as floats), z window +-1 m
CLOUD_FORMAT = dict(sqr_laser_min_dist=float(np.float32(0.4 * 0.4)), sqr_laser_max_dist=float(np.float32(30.0 * 30.0)),
                    laser_z_min_value=-1.0, laser_z_max_value=1.0)

SCAN_FORMAT = dict(n_beams=N_BEAMS, angle_min=float(ANGLE_MIN), angle_increment=float(ANGLE_INC),
                   range_min=float(RANGE_MIN), range_max=float(RANGE_MAX))


def make_scan_batch(world: World, poses: np.ndarray, noise_seed: int = 7, scale_to_map: float = 20.0,
                    sigma: float = 0.01):
    """Scans for every pose. Returns (pts (sum_n, 2) f32, offsets (B+1,) i32)."""
    rng = np.random.default_rng(noise_seed)
    chunks, offsets = [], [0]
    for p in poses:
        s = make_scan(world, p, rng, scale_to_map, sigma)
        chunks.append(s)
        offsets.append(offsets[-1] + s.shape[0])
    pts = np.concatenate(chunks, axis=0) if chunks else np.zeros((0, 2), np.float32)
    return np.ascontiguousarray(pts, dtype=np.float32), np.asarray(offsets, dtype=np.int32)


def perturb_hints(poses: np.ndarray, seed: int = 1, dxy: float = 0.1, dpsi: float = 0.05) -> np.ndarray:
    """hint = truth + U[-dxy, dxy] m, U[-dpsi, dpsi] rad, float32."""
    rng = np.random.default_rng(seed)
    h = poses.copy()
    h[:, 0] += rng.uniform(-dxy, dxy, len(poses))
    h[:, 1] += rng.uniform(-dxy, dxy, len(poses))
    h[:, 2] += rng.uniform(-dpsi, dpsi, len(poses))
    return np.ascontiguousarray(h, dtype=np.float32)
