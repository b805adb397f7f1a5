"""The C++ host façade (hector_slam_b200/host/SlamProcessorB200.hpp) through its ctypes shim.
CPU part: the map-update gate util::poseDifferenceLargerThan (UtilFunctions.h:73-92) against the
oracle's behaviour.  GPU part: HectorSlamProcessor::update semantics over the golden SLAM run and
with the gate active."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, golden_planes, load_golden, pose_err

HOST_LIB = os.path.join(ROOT, "hector_slam_b200", "lib", "libhsb200_host.so")


def host(hsb_lib):
    L = C.CDLL(HOST_LIB)
    L.hsbp_create.restype = C.c_void_p
    L.hsbp_create.argtypes = [C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]
    L.hsbp_destroy.argtypes = [C.c_void_p]
    L.hsbp_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.hsbp_reset.argtypes = [C.c_void_p]
    L.hsbp_set_update_factors.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.hsbp_set_map_update_thresholds.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.hsbp_get_grid_map.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.hsbp_pose_difference_larger_than.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float]
    return L


def test_pose_difference_gate(hsb_lib):
    L = host(hsb_lib)
    f = np.float32

    def gate(a, b, d, an):
        a, b = np.asarray(a, f), np.asarray(b, f)
        return bool(L.hsbp_pose_difference_larger_than(a.ctypes.data, b.ctypes.data, d, an))

    assert not gate([0, 0, 0], [0.3, 0, 0], 0.4, 0.9)
    assert gate([0, 0, 0], [0.5, 0, 0], 0.4, 0.9)
    assert gate([0, 0, 0.0], [0, 0, 1.0], 0.4, 0.9)
    assert not gate([0, 0, 3.1], [0, 0, -3.1], 0.4, 0.9)          # wraps: 6.2 - 2*pi = -0.083
    assert gate([0, 0, 0], [3.4e38, 3.4e38, 3.4e38], 0.4, 0.9)    # FLT_MAX after reset: always writes
    assert not gate([1, 1, 0.05], [1, 1, 0.0], 0.4, 0.06)
    assert gate([1, 1, 0.07], [1, 1, 0.0], 0.4, 0.06)


@pytest.mark.gpu
def test_facade_slam_run(hsb_lib, pyoracle):
    """update() through the C++ façade == HectorSlamProcessor::update of the oracle, with the
    map-update gate active (only some scans write the map)."""
    L = host(hsb_lib)
    g = load_golden("slam3.npz")
    res, size = float(g["res"]), int(g["size"])
    p = L.hsbp_create(res, size, size, 0.5, 0.5, 3, 0)
    assert p
    orc = pyoracle.Oracle("port", res, size, 3)
    for dist, ang in ((0.0, 0.0), (0.35, 0.05)):
        L.hsbp_reset(p)
        orc.reset()
        L.hsbp_set_update_factors(p, 0.4, 0.9)
        orc.set_update_factors(0.4, 0.9)
        L.hsbp_set_map_update_thresholds(p, dist, ang)
        orc.set_map_update_thresholds(dist, ang)
        hint_g = g["first_hint"].copy()
        hint_o = g["first_hint"].copy()
        origo = np.zeros(2, np.float32)
        for k in range(g["scans"].shape[0]):
            scan = np.ascontiguousarray(g["scans"][k])
            pose, cov = np.zeros(3, np.float32), np.zeros(9, np.float32)
            assert L.hsbp_update(p, scan.ctypes.data, scan.shape[0], origo.ctypes.data, hint_g.ctypes.data, 0,
                                 pose.ctypes.data, cov.ctypes.data) == 0
            want, _ = orc.update(scan, hint_o)
            ex, ey, ea = pose_err(pose, want)
            assert max(ex, ey) <= 1e-4 and ea <= 1e-4, (dist, k, ex, ey, ea)
            if dist == 0.0:
                assert np.abs(pose - g["est"][k]).max() <= 1e-4
            hint_g, hint_o = pose, want
        for l in range(3):
            sz = size >> l
            got = np.zeros((sz, sz), np.float32)
            assert L.hsbp_get_grid_map(p, l, got.ctypes.data) == 0
            diff = np.abs(got - orc.get_logodds(l))
            assert (diff > 1e-5).sum() <= max(3, int(2e-3 * (got != 0).sum())), (dist, l, int((diff > 1e-5).sum()))
    L.hsbp_destroy(p)
    orc.close()
