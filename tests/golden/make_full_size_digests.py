#!/usr/bin/env python
"""Generates tests/golden/full_size_digests.json: sha256 digests of the result arrays the UNMODIFIED SDK build (oracle/_ref/libomm-lib.so, compiled from
/root/reference by oracle/Makefile) produces for BASELINE configs 2, 3 and 5 at full size.  bench.py prints the same digest of the GPU result at every
N and compares it with these (and, at N=1, with a full SDK bake made in the same run); tests use them where oracle/_ref is absent.
Digest = sha256(arrayData | descArray | descArrayHistogram | indexBuffer | indexHistogram | indexFormat byte), see omm_b200/baker.py::result_sha256.
usage: python tests/golden/make_full_size_digests.py      (build container; config 3 takes ~7 minutes on 8 cores and ~10 GB)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
from omm_b200 import Baker, capi  # noqa: E402
from omm_b200 import workloads as W  # noqa: E402
from omm_b200.baker import result_sha256  # noqa: E402

ref = capi.OmmLib(os.path.join(ROOT, "oracle", "_ref", "libomm-lib.so"))
out = {}
for name, wl in (("C2", W.config2()), ("C5", W.config5()), ("C3", W.config3())):
    t0 = time.time()
    with Baker(ref) as b:
        inp, tex = W.make_input(b, wl, bake_flags=capi.BAKE_ENABLE_INTERNAL_THREADS)
        res = b.bake(inp)
        tex.destroy()
    out[name] = {"workload": wl.name, "sha256": result_sha256(res), "array_data_bytes": int(res.array_data.size), "desc_count": int(res.desc_array.size),
                 "index_count": int(res.index_buffer.size), "sdk_seconds": round(time.time() - t0, 1), "cores": os.cpu_count()}
    print(name, out[name], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "full_size_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
