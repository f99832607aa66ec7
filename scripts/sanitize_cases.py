#!/usr/bin/env python
"""Small bakes for compute-sanitizer: BASELINE config 1, a slice of config 3 (hierarchical classifier, exact dedup), a small mixed-level
config 5 (big footprints, constant-area tables), the flat kernels (Nearest), SAT, 2-state packing, and the optional passes (near-duplicate merge, budget
compression).  Every result is compared with the CPU checker, so a sanitizer run is also a parity run.
usage: compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python scripts/sanitize_cases.py [case name ...]"""
import os
import sys

os.environ.setdefault("OMM_B200_SCRATCH_BLOCKS", "0")   # one allocation per array: the sanitizer sees every bound

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from omm_b200 import capi, load_product_library  # noqa: E402
from omm_b200 import workloads as W  # noqa: E402
import parity_cases as PC  # noqa: E402

lib = load_product_library()
ref_path = os.path.join(ROOT, "oracle", "_ref", "libomm-lib.so")
checker = capi.OmmLib(ref_path if os.path.exists(ref_path) else os.path.join(ROOT, "oracle", "liboracle_port.so"))
cases = {
    "C1": (W.config1(), {}),
    "C3 slice": (W.config3(num_tris=400, tex_size=256, level=5), {}),
    "C5 small": (W.config5(num_tris=1200, tex_size=256, distinct=200, flat_tris=300, max_level=7), {}),
    "nearest": (W.random_mesh(31, 120, filter=capi.FILTER_NEAREST, unknown_state_promotion=capi.PROMOTE_NEAREST, uv_lo=-0.5, uv_hi=1.5), {}),
    "sat": (W.random_mesh(44, 150, tex_kind="blocky", tex_alpha_cutoff=0.5, addressing_mode=capi.ADDR_CLAMP, tri_texels=20, max_subdivision_level=4), {}),
    "2-state mips": (W.random_mesh(43, 120, mips=3, format=capi.FORMAT_2_STATE), {}),
    "big blocks": (W.config3(num_tris=3, tex_size=256, level=9), {}),   # digest kernel with a producer and a chain warp (shared-memory double buffer)
}
if checker.path.endswith("libomm-lib.so"):
    cases["near-duplicates (LSH)"] = (W.random_mesh(101, 250, tex_kind="blocky", tri_texels=14, max_subdivision_level=3, reuse_frac=0.1, bake_flags=capi.BAKE_ENABLE_NEAR_DUPLICATE_DETECTION), {})
    cases["budget compression"] = (W.random_mesh(104, 150, tri_texels=16, max_subdivision_level=4, max_array_data_size=4000), {})
if len(sys.argv) > 1:   # only the named cases
    cases = {k: v for k, v in cases.items() if k in sys.argv[1:]}
bad = 0
for name, (wl, over) in cases.items():
    got = PC.run_bake(lib, wl, **over)
    want = PC.run_bake(checker, wl, **over)
    d = got.diff(want)
    print(f"{name}: {'identical' if not d else d}", flush=True)
    bad += bool(d)
sys.exit(1 if bad else 0)
