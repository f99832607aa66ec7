"""Seeded slice of scripts/oracle_campaign.py in the CPU suite: random bakes through the plain-C oracle port and the SDK build (oracle/_ref),
byte for byte.  The same generator (tests/campaign.py::random_bake) drives the GPU campaign (tests/test_gpu_campaign.py)."""
import numpy as np

import campaign
import parity_cases as PC


def test_port_matches_sdk_on_random_bakes(ref_lib, port_lib):
    rng = np.random.default_rng(4242)
    for run in range(40):
        wl, kw = campaign.random_bake(rng)
        res = []
        for lib in (ref_lib, port_lib):
            try:
                res.append(PC.run_bake(lib, wl))
            except RuntimeError as e:
                res.append(str(e).split(";")[0])
        if isinstance(res[0], str) or isinstance(res[1], str):
            assert res[0] == res[1], (run, kw)
        else:
            assert res[0].diff(res[1]) == [], (run, kw)
