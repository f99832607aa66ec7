#!/usr/bin/env python
"""One process, config 3 staged once: device-resident bakes under a list of tuning settings (chunk lanes, chunk size, grid shapes -- all read per
bake from the environment), then the drop-in call and the result digest for each.  usage: python scripts/sweep_lanes.py [steps=4] [set ...]
where a set is  name:VAR=val,VAR=val  (default: the built-in list).  OMM_SWEEP_CONFIG=C5 runs BASELINE config 5 instead of config 3."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from omm_b200 import Baker, capi  # noqa: E402
from omm_b200 import workloads as W  # noqa: E402
from omm_b200.baker import _copy_result, result_sha256  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
M = 1 << 20
default_sets = [
    ("default (1 lane, 32M regions, 128 blocks/SM)", {}),
    ("2 lanes, 32M", {"OMM_B200_CHUNK_LANES": 2}),
    ("2 lanes, 8M", {"OMM_B200_CHUNK_LANES": 2, "OMM_B200_CHUNK_REGIONS": 8 * M}),
    ("2 lanes, 8M, 32 blocks/SM", {"OMM_B200_CHUNK_LANES": 2, "OMM_B200_CHUNK_REGIONS": 8 * M, "OMM_B200_LIST_GRID_MULT": 32}),
    ("2 lanes, 4M, 32 blocks/SM", {"OMM_B200_CHUNK_LANES": 2, "OMM_B200_CHUNK_REGIONS": 4 * M, "OMM_B200_LIST_GRID_MULT": 32}),
    ("2 lanes, 4M, 16 blocks/SM", {"OMM_B200_CHUNK_LANES": 2, "OMM_B200_CHUNK_REGIONS": 4 * M, "OMM_B200_LIST_GRID_MULT": 16}),
    ("3 lanes, 4M, 16 blocks/SM", {"OMM_B200_CHUNK_LANES": 3, "OMM_B200_CHUNK_REGIONS": 4 * M, "OMM_B200_LIST_GRID_MULT": 16}),
    ("4 lanes, 8M, 32 blocks/SM", {"OMM_B200_CHUNK_LANES": 4, "OMM_B200_CHUNK_REGIONS": 8 * M, "OMM_B200_LIST_GRID_MULT": 32}),
    ("4 lanes, 2M, 16 blocks/SM", {"OMM_B200_CHUNK_LANES": 4, "OMM_B200_CHUNK_REGIONS": 2 * M, "OMM_B200_LIST_GRID_MULT": 16}),
    ("1 lane, 8M, 32 blocks/SM", {"OMM_B200_CHUNK_REGIONS": 8 * M, "OMM_B200_LIST_GRID_MULT": 32}),
    ("default, slow grid 8", {"OMM_B200_SLOW_GRID_MULT": 8}),
    ("default again", {}),
]
sets = default_sets
if len(sys.argv) > 2:
    sets = []
    for a in sys.argv[2:]:
        name, _, kv = a.partition(":")
        sets.append((name, dict(x.split("=") for x in kv.split(",") if x)))
TUNING = ["OMM_B200_CHUNK_LANES", "OMM_B200_CHUNK_REGIONS", "OMM_B200_LIST_GRID_MULT", "OMM_B200_INIT_GRID_MULT", "OMM_B200_LEAF_GRID_MULT", "OMM_B200_SLOW_GRID_MULT",
          "OMM_B200_BIG_HASH"]

torch.cuda.set_device(0)
lib = capi.load_product_library()
assert lib.dll.ommB200SetDevice(0) == capi.SUCCESS
baker = Baker(lib)
CONFIG = os.environ.get("OMM_SWEEP_CONFIG", "C3")
wl = W.config5() if CONFIG == "C5" else W.config3()
wl.indices, _k1 = bench.pinned_like(wl.indices)
wl.texcoords, _k2 = bench.pinned_like(wl.texcoords)
inp, tex = W.make_input(baker, wl)
desc = inp.to_desc()
golden = bench.golden_digests().get(CONFIG, {}).get("sha256")
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tm = capi.B200BakeTimings()
staged = C.c_void_p()
assert lib.dll.ommB200StageInputs(baker.handle, C.byref(desc), C.byref(staged)) == capi.SUCCESS
out = []
for name, env in sets:
    for k in TUNING:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    ms, cl, ip, post = [], [], [], []
    for it in range(2 + steps):
        flush.fill_(it & 0xFF)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h = C.c_void_p()
        rc = lib.dll.ommB200BakeResident(baker.handle, staged, sp, C.byref(h))
        e1.record(stream)
        assert rc == capi.SUCCESS, rc
        torch.cuda.synchronize()
        lib.dll.ommB200GetLastBakeTimings(baker.handle, C.byref(tm))
        if it >= 2:
            ms.append(e0.elapsed_time(e1)); cl.append(tm.classifyMs); ip.append(tm.itemPostMs); post.append(tm.postMs)
        lib.dll.ommCpuDestroyBakeResult(h)
    e2e, sha = [], None
    for it in range(3):
        t0 = time.perf_counter()
        h = C.c_void_p()
        rc = lib.dll.ommCpuBake(baker.handle, C.byref(desc), C.byref(h))
        pdesc = C.POINTER(capi.CpuBakeResultDesc)()
        rc2 = lib.dll.ommCpuGetBakeResultDesc(h, C.byref(pdesc))
        dt = time.perf_counter() - t0
        assert rc == capi.SUCCESS and rc2 == capi.SUCCESS, (rc, rc2)
        if it >= 1:
            e2e.append(dt * 1e3)
        if it == 2:
            sha = result_sha256(_copy_result(pdesc.contents))
        lib.dll.ommCpuDestroyBakeResult(h)
    row = dict(name=name, env=env, step_ms=round(sum(ms) / len(ms), 3), min_ms=round(min(ms), 3), classify_ms=round(sum(cl) / len(cl), 3), item_post_ms=round(sum(ip) / len(ip), 3),
               post_ms=round(sum(post) / len(post), 3), e2e_ms=round(min(e2e), 3), launches=int(tm.kernelLaunches), matches_golden=(sha == golden))
    out.append(row)
    print(json.dumps(row), flush=True)
lib.dll.ommB200DestroyStagedInputs(staged)
