"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol that
include/hector_slam_b200.h declares.  No compute call is made (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hector_slam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(hsb_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_and_binding_agree():
    from hector_slam_b200 import capi

    assert sorted(capi.EXPORTED) == declared_symbols()


def test_library_exports_every_declared_symbol(hsb_lib):
    for name in declared_symbols():
        assert hasattr(hsb_lib, name), name
    assert b"sm_100a" in hsb_lib.hsb_version()
    assert hsb_lib.hsb_status_string(0) == b"ok"
    assert hsb_lib.hsb_status_string(-4) == b"no CUDA device"


def test_library_contains_sm100a_code_only(hsb_lib):
    from hector_slam_b200 import capi

    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        import pytest

        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_create_without_gpu_fails_loudly(hsb_lib):
    """No silent CPU fallback: without a device hsb_create reports HSB_ERR_NO_DEVICE."""
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is visible")
    from hector_slam_b200 import capi

    try:
        capi.MapRepB200(0.05, 256, levels=1)
    except capi.HsbError as e:
        assert e.status == -4
    else:
        raise AssertionError("hsb_create succeeded without a GPU")
