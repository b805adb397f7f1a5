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


EXP_SRC = r'''
#include <cstdio>
#include <cmath>
#include <cstring>
#include <cstdint>
#include "sincosf_glibc.h"
static bool same(float a, float b) { uint32_t x, y; memcpy(&x, &a, 4); memcpy(&y, &b, 4); return x == y || (a != a && b != b); }
int main() {
  long bad = 0, n = 0; unsigned s = 2463534242u;
  // the log-odds a grid can hold: sums of log(0.4/0.6) and log(0.9/0.1) steps, anything in (-inf, 52.2]; plus the whole
  // float range around the special-case thresholds
  for (long i = 0; i < 16000000; ++i) {
    s = s * 1664525u + 1013904223u;
    float u = (float)(s >> 8) / 16777216.0f * 2.f - 1.f;
    float x = (i % 4 == 0) ? u * 52.2f : (i % 4 == 1) ? u * 110.f : (i % 4 == 2) ? ldexpf(u, -(int)(i % 60)) : -103.f - u;
    ++n;
    if (!same(hsb::expf_glibc(x), expf(x))) ++bad;
  }
  // every float between the underflow thresholds and a little beyond, and the overflow edge
  for (float x = -104.2f; x < -103.0f; x = nextafterf(x, 0.f)) { ++n; if (!same(hsb::expf_glibc(x), expf(x))) ++bad; }
  for (float x = 88.5f; x < 89.0f; x = nextafterf(x, 1e9f)) { ++n; if (!same(hsb::expf_glibc(x), expf(x))) ++bad; }
  const float edge[] = {0.f, -0.f, 1.f, -1.f, 50.f, 52.197224f, -0.4054651f, 2.1972246f, 87.99999f, 88.f, -87.99999f, -88.f,
                        INFINITY, -INFINITY, NAN, 1e-30f, -1e-30f, 3.4e38f, -3.4e38f};
  for (float x : edge) { ++n; if (!same(hsb::expf_glibc(x), expf(x))) ++bad; }
  // cumulative log-odds exactly as a cell accumulates them
  const float lf = logf(0.4f / 0.6f), lo = logf(0.9f / 0.1f);
  for (int occ = 0; occ < 40; ++occ) {
    float l = 0.f;
    for (int k = 0; k < occ; ++k) if (l < 50.f) l += lo;
    for (int k = 0; k < 400; ++k) { l += lf; ++n; if (!same(hsb::expf_glibc(l), expf(l))) ++bad; }
  }
  std::printf("%ld %ld\n", n, bad);
  return bad ? 1 : 0;
}
'''


def test_expf_matches_libm(tmp_path):
    """hsb::expf_glibc (the exp of getGridProbability, GridMapLogOdds.h:165) against the running libm, bit for bit."""
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    flags = open("/proc/cpuinfo").read()
    if " fma" not in flags or " avx2" not in flags:
        pytest.skip("host CPU without FMA/AVX2: libm does not run the code path the header transcribes")
    src = tmp_path / "e.cpp"
    src.write_text(EXP_SRC)
    exe = tmp_path / "e"
    subprocess.run([cxx, "-O2", "-I", os.path.join(ROOT, "hector_slam_b200", "csrc"), str(src), "-o", str(exe), "-lm"],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    n, bad = (int(v) for v in out.stdout.split())
    assert bad == 0 and n > 16000000, out.stdout
