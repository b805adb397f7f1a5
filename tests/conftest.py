"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI export checks.
`-m gpu` runs on a B200: CUDA path vs the oracle / golden vectors, through the C-ABI.
/root/reference is never read by a test: where the compiled reference (oracle/_ref/libhsref.so)
is present it is used as a second oracle, otherwise the plain-C port alone.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_port():
    from oracle import pyoracle

    if not pyoracle.available("port"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
    return pyoracle


@pytest.fixture(scope="session")
def pyoracle():
    return _ensure_port()


@pytest.fixture(scope="session")
def oracle_kinds(pyoracle):
    kinds = ["port"]
    if pyoracle.available("reference"):
        kinds.append("reference")
    return kinds


@pytest.fixture(scope="session")
def hsb_lib():
    """Build (if needed) and load the CUDA library; never falls back to anything else."""
    from hector_slam_b200 import build, capi

    build.build_all()
    return capi.load_library()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def densify(g, prefix, level, size):
    plane = np.zeros(size * size, np.float32)
    plane[g[f"{prefix}_idx{level}"]] = g[f"{prefix}_val{level}"]
    return plane.reshape(size, size)


def golden_planes(g, prefix="map"):
    size, levels = int(g["size"]), int(g["levels"])
    return [densify(g, prefix, l, size >> l) for l in range(levels)]


def apply_diff(planes, g, prefix):
    out = []
    for l, p in enumerate(planes):
        q = p.copy().reshape(-1)
        q[g[f"{prefix}_idx{l}"]] = g[f"{prefix}_val{l}"]
        out.append(q.reshape(p.shape))
    return out


def angle_diff(a, b):
    d = np.asarray(a, np.float64) - np.asarray(b, np.float64)
    return np.abs((d + np.pi) % (2 * np.pi) - np.pi)


def pose_err(a, b):
    a = np.asarray(a, np.float64).reshape(-1, 3)
    b = np.asarray(b, np.float64).reshape(-1, 3)
    return np.abs(a[:, 0] - b[:, 0]).max(), np.abs(a[:, 1] - b[:, 1]).max(), angle_diff(a[:, 2], b[:, 2]).max()


def report(line: str):
    """Append a line to gpurun_out/parity_report.log (observed parity figures: differing cells, worst pose
    differences) — pytest -q swallows stdout, and the tolerances in the tests are pinned from these."""
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.log"), "a") as f:
            f.write(line.rstrip() + "\n")
    except OSError:
        pass
    print(line)
