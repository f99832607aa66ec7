"""Re-entrancy of ommCpuBake (ref: docs/integration_guide.md:434 -- a baker is stateless after creation, so concurrent bakes on one baker and one
texture are legal; SURVEY 8b "Threading").  Several host threads bake through ONE baker and ONE texture with different alpha cutoffs, promotions and
formats at the same time; every result must be byte-identical with the same bake run alone.  Different cutoffs on one texture is the case that used
to share a single per-texture table of the hierarchical classifier (round-1 advisor finding): the tables are now per (texture, cutoff)."""
import threading

import numpy as np
import pytest

from omm_b200 import Baker, capi
from omm_b200 import workloads as W
from omm_b200.baker import BakeInput

pytestmark = pytest.mark.gpu

VARIANTS = [
    dict(alpha_cutoff=0.5),
    dict(alpha_cutoff=0.3, unknown_state_promotion=capi.PROMOTE_NEAREST),
    dict(alpha_cutoff=0.7, format=capi.FORMAT_2_STATE),
    dict(alpha_cutoff=0.4, unknown_state_promotion=capi.PROMOTE_FORCE_TRANSPARENT),
    dict(alpha_cutoff=0.6, bake_flags=capi.BAKE_DISABLE_SPECIAL_INDICES),
    dict(alpha_cutoff=0.2, filter=capi.FILTER_NEAREST),
    dict(alpha_cutoff=0.55, bake_flags=capi.BAKE_DISABLE_DUPLICATE_DETECTION, max_subdivision_level=4),
    dict(alpha_cutoff=0.45, alpha_cutoff_gt=capi.STATE_T, alpha_cutoff_le=capi.STATE_O),
]


def _input(tex, wl, over):
    kw = dict(wl.desc)
    kw.update(over)
    return BakeInput(texture=tex, indices=wl.indices, texcoords=wl.texcoords, texcoord_format=wl.texcoord_format, **kw)


@pytest.mark.parametrize("rounds", [3])
def test_concurrent_bakes_on_one_baker_and_texture(product_lib, checker_lib, rounds):
    wl = W.config5(num_tris=6000, tex_size=512, distinct=1500, flat_tris=800, max_level=6)   # constant areas: the (H) tables are really used
    wl.subdivision_levels = None
    wl.desc["max_subdivision_level"] = 5
    with Baker(product_lib) as b:
        tex = b.create_texture(wl.mips)
        # the same bakes alone, in order
        serial = [b.bake(_input(tex, wl, v)) for v in VARIANTS]
        results = [[None] * rounds for _ in VARIANTS]
        errors = []
        start = threading.Barrier(len(VARIANTS))

        def worker(i):
            try:
                start.wait()
                for r in range(rounds):
                    results[i][r] = b.bake(_input(tex, wl, VARIANTS[i]))
            except Exception as e:  # noqa: BLE001
                errors.append((i, repr(e)))

        threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(VARIANTS))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        for i, v in enumerate(VARIANTS):
            for r in range(rounds):
                assert results[i][r].diff(serial[i]) == [], f"variant {i} {v}, round {r}: concurrent result differs from the serial one"
        # and the serial results are the SDK's (one texture object per cutoff on the checker is not needed: the SDK texture is cutoff-free here)
        with Baker(checker_lib) as cb:
            ctex = cb.create_texture(wl.mips)
            for i, v in enumerate(VARIANTS):
                assert serial[i].diff(cb.bake(_input(ctex, wl, v))) == [], f"variant {i} {v}: differs from {checker_lib.path}"
            ctex.destroy()
        tex.destroy()


def test_more_cutoffs_than_cached_table_sets(product_lib):
    """Cycling through more cutoffs than the texture keeps idle table sets for: every bake still sees the tables of ITS cutoff."""
    wl = W.config5(num_tris=2000, tex_size=256, distinct=500, flat_tris=500, max_level=5)
    with Baker(product_lib) as b:
        tex = b.create_texture(wl.mips)
        cutoffs = [0.1, 0.3, 0.5, 0.7, 0.9, 0.3, 0.1, 0.9]
        first = {}
        for c in cutoffs:
            inp, _ = None, None
            kw = dict(wl.desc)
            kw["alpha_cutoff"] = c
            res = b.bake(BakeInput(texture=tex, indices=wl.indices, texcoords=wl.texcoords, subdivision_levels=wl.subdivision_levels, **kw))
            if c in first:
                assert res.diff(first[c]) == []
            first[c] = res
        assert len({r.array_data.tobytes() for r in first.values()}) > 1
        tex.destroy()
