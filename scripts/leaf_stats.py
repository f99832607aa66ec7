#!/usr/bin/env python
"""Where the leaf walk's work goes: host build of omm_hier.cuh with its counters compiled in, run over a slice
of BASELINE config 3.  No GPU needed.   usage: python scripts/leaf_stats.py [triangles=3000]"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_hier_host as T  # noqa: E402
from omm_b200 import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hier_host"), "libhier_host_stats.so"])
lib = ctypes.CDLL(os.path.join(ROOT, "tests", "hier_host", "libhier_host_stats.so"))
wl = W.config3(num_tris=n)
st = T.check(lib, wl.mips[0], wl.texcoords.reshape(-1, 6), np.full(n, 6))
out = (ctypes.c_ulonglong * 16)()
lib.hier_host_stats(out)
o = list(out)
print(f"{st.microTriangles} micro-triangles of {n} work items; region tests (64/16/4): {list(st.tests)[:3]}, passes {list(st.passes)[:3]}")
print(f"leaves: {o[10]} = {100 * o[10] / st.microTriangles:.2f} % of the micro-triangles; crossed by the level line: {100 * o[11] / max(o[10], 1):.1f} % of the leaves")
print(f"leaf cells: {o[0]} ({o[0] / max(o[10], 1):.2f} per leaf); closed by the edge filter (D): {100 * o[2] / o[0]:.1f} %; skipped by (E): {100 * o[3] / o[0]:.1f} %; "
      f"vertex values of both signs: {100 * (o[0] - o[1]) / o[0]:.1f} %")
print(f"cells that run the three edge tests: {o[8]} ({100 * o[8] / o[0]:.1f} %), with a hit: {100 * o[9] / max(o[8], 1):.1f} %")
print(f"(D) failed although the vertex values have one sign: near the line {o[12]}, d is rounding noise (< 1e-6) {o[13]}, other {o[14]} "
      f"({100 * o[13] / o[0]:.1f} % of the cells are the rounding-noise case)")
