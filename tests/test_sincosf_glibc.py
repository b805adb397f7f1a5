"""CPU: hector_slam_b200/csrc/sincosf_glibc.h (the sinf/cosf the kernels use for the pose angle)
against the running libm, bit for bit.  The header transcribes glibc 2.39's FMA build, which libm
selects through ifunc on every AVX2+FMA host; on a host without FMA libm takes another code path
and the comparison is skipped."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

SRC = r'''
#include <cstdio>
#include <cmath>
#include "sincosf_glibc.h"
int main() {
  long bad = 0, n = 0; unsigned s = 2463534242u;
  for (long i = 0; i < 12000000; ++i) {
    s = s * 1664525u + 1013904223u;
    float u = (float)(s >> 8) / 16777216.0f * 2.f - 1.f;
    float x = (i % 3 == 0) ? u * 3.1415927f : (i % 3 == 1) ? u * 119.9f : ldexpf(u, -(int)(i % 40));
    ++n;
    if (hsb::sinf_glibc(x) != sinf(x)) ++bad;
    if (hsb::cosf_glibc(x) != cosf(x)) ++bad;
    float sc, cc; hsb::sincosf_glibc(x, &sc, &cc);
    if (sc != sinf(x) || cc != cosf(x)) ++bad;
  }
  const float edge[] = {0.f, -0.f, 0.78539816f, 0.785398185f, 0.78539822f, 1.5707963f, 3.1415927f, -3.1415927f, 6.2831855f,
                        2.4414062e-4f, 2.4414065e-4f, 119.99999f};
  for (float x : edge) {
    if (hsb::sinf_glibc(x) != sinf(x)) ++bad; if (hsb::cosf_glibc(x) != cosf(x)) ++bad;
    float sc, cc; hsb::sincosf_glibc(x, &sc, &cc);
    if (sc != sinf(x) || cc != cosf(x)) ++bad;
  }
  std::printf("%ld %ld\n", n, bad);
  return bad ? 1 : 0;
}
'''


def test_sincosf_matches_libm(tmp_path):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    flags = open("/proc/cpuinfo").read()
    if " fma" not in flags or " avx2" not in flags:
        pytest.skip("host CPU without FMA/AVX2: libm does not run the code path the header transcribes")
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run([cxx, "-O2", "-I", os.path.join(ROOT, "hector_slam_b200", "csrc"), str(src), "-o", str(exe), "-lm"],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    n, bad = (int(v) for v in out.stdout.split())
    assert bad == 0 and n == 12000000, out.stdout
