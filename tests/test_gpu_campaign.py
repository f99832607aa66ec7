"""Randomized GPU-vs-oracle campaign (-m gpu): the PRODUCT (nvcc device build: -fmad=false, device sqrtf / division, the hierarchical kernels) against
the strongest CPU checker (oracle/_ref = the unmodified SDK build when it travelled with the repo, else the plain-C port) on random bakes, byte for byte.
The fixed parity cases pin known corners; this pins the space between them -- the one real parity bug of round 1 (SAT x Mirror) was found by
randomization.  The reference's analogous pin is running itself on every bake (support/tests/test_omm_bake_cpu.cpp:323-344).

Configuration space (tests/campaign.py::random_bake): 5 address modes x pow2 / npot / tiny textures x FP32 / UNORM8 x 1-4 mips x SAT on / off / other
cutoff x Linear / Nearest x 3 promotions x 2 formats x state mappings x per-triangle levels 0-8 (13 = global) x dynamic levels x degenerate / NaN /
reused triangles x 8/16/32-bit indices x bake flags; a second leg (10 bakes) adds work items of level 9-12.  OMM_CAMPAIGN_BAKES scales the first leg (default 260)."""
import os

import numpy as np
import pytest

import campaign
from omm_b200 import Baker
from omm_b200 import workloads as W
from omm_b200.baker import OmmError

pytestmark = pytest.mark.gpu


def _bake(lib, wl):
    with Baker(lib) as b:
        inp, tex = W.make_input(b, wl)
        try:
            return b.bake(inp)
        except OmmError as e:
            return e.result
        finally:
            tex.destroy()


def _compare(product_lib, checker_lib, wl, kw, run):
    want = _bake(checker_lib, wl)
    got = _bake(product_lib, wl)
    if isinstance(want, int) or isinstance(got, int):
        assert want == got, f"run {run}: result codes differ (checker {want}, product {got}); config {kw}"
        return 0
    d = got.diff(want)
    assert d == [], f"run {run}: CUDA result differs from {checker_lib.path}: {d}; config {kw}"
    return int(got.timings.microTriangles) if got.timings is not None else 0


def test_random_bakes_match_the_checker(product_lib, checker_lib):
    bakes = int(os.environ.get("OMM_CAMPAIGN_BAKES", "260"))
    rng = np.random.default_rng(int(os.environ.get("OMM_CAMPAIGN_SEED", "20261017")))
    total = 0
    for run in range(bakes):
        wl, kw = campaign.random_bake(rng)
        total += _compare(product_lib, checker_lib, wl, kw, run)
    assert total > 1_000_000, total


def test_random_bakes_with_big_levels_match_the_checker(product_lib, checker_lib):
    """Work items of level 9-12 (up to 16.7 M micro-triangles in one block) among small ones: several classifier chunks per item, XXH64 of
    megabyte blocks, special-index promotion of huge uniform blocks."""
    rng = np.random.default_rng(99)
    slow_checker = not checker_lib.path.endswith("libomm-lib.so")   # the scalar port: keep it short
    for run in range(4 if slow_checker else 10):
        wl, kw = campaign.random_bake(rng, big_levels=True)
        if slow_checker:
            wl.subdivision_levels = np.minimum(wl.subdivision_levels, 9)
        _compare(product_lib, checker_lib, wl, kw, run)
