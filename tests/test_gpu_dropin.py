"""GPU: the UNMODIFIED reference façade (HectorSlamProcessor::update) with MapRepB200 plugged in,
against the same façade on the reference's own CPU maps — compiled from the reference headers in
the build container (oracle/Makefile `dropin`), run here as a prebuilt binary."""
import os
import subprocess

import pytest

from conftest import ROOT, report

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_check")


def test_reference_facade_with_b200_maprep(hsb_lib):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dropin_check not built (needs /root/reference at build time)")
    out = subprocess.run([BIN, "60"], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    for line in out.stdout.splitlines():
        if 'cells touched' in line or 'max' in line:
            report('dropin_check: ' + line)
    assert out.returncode == 0 and "DROPIN OK" in out.stdout, out.stdout + out.stderr
