"""Parity cases shared by the CPU (port vs SDK) and GPU (product vs checker) suites."""
import numpy as np

from omm_b200 import capi
from omm_b200 import workloads as W

A = capi


def _levels(seed, n, lo, hi):
    return (lo + (W._unit(seed, n, 77) * (hi - lo + 1)).astype(np.int64)).astype(np.uint8)


def cases():
    """name -> (workload factory, BakeInput overrides)"""
    c = {}
    c["c1_quad_checker_l3_2state"] = (lambda: W.config1(), {})
    c["c2_small_l4"] = (lambda: W.config2(num_quads=300, tex_size=256, level=4), {})
    c["c2_small_unorm8"] = (lambda: W.config2(num_quads=200, tex_size=256, level=4, unorm8=True), {})
    c["c3_small_l5"] = (lambda: W.config3(num_tris=600, tex_size=256, level=5), {})
    c["c3_small_l6_nearest_promo"] = (lambda: W.config3(num_tris=300, tex_size=256, level=6, promotion=A.PROMOTE_NEAREST), {})
    c["c3_small_sat"] = (lambda: W.config3(num_tris=500, tex_size=256, level=5, tex_alpha_cutoff=0.5), {})
    c["c5_small_mixed_levels"] = (lambda: W.config5(num_tris=3000, tex_size=256, distinct=256, flat_tris=500, max_level=7), {})
    # address modes x pow2 / non-pow2, UVs reaching outside [0,1]
    for mode, mname in ((A.ADDR_WRAP, "wrap"), (A.ADDR_MIRROR, "mirror"), (A.ADDR_CLAMP, "clamp"), (A.ADDR_MIRROR_ONCE, "mirroronce")):
        for size, sname in (((128, 128), "pow2"), ((100, 75), "npot")):
            c[f"addr_{mname}_{sname}"] = (lambda mode=mode, size=size: W.random_mesh(11, 400, tex_size=size, uv_lo=-1.3, uv_hi=2.3,
                                                                                     addressing_mode=mode, max_subdivision_level=3), {})
    # Border: the SDK's point sample reads out of bounds when the footprint leaves the texture -> stay inside
    c["addr_border_inside"] = (lambda: W.random_mesh(12, 400, tex_size=(128, 128), uv_lo=0.2, uv_hi=0.8, tri_texels=8,
                                                     addressing_mode=A.ADDR_BORDER, border_alpha=0.7, max_subdivision_level=3), {})
    for promo, pname in ((A.PROMOTE_NEAREST, "nearest"), (A.PROMOTE_FORCE_OPAQUE, "fo"), (A.PROMOTE_FORCE_TRANSPARENT, "ft")):
        c[f"promo_{pname}_4state"] = (lambda promo=promo: W.random_mesh(21, 300, unknown_state_promotion=promo), {})
        c[f"promo_{pname}_2state"] = (lambda promo=promo: W.random_mesh(22, 300, unknown_state_promotion=promo, format=A.FORMAT_2_STATE), {})
    c["filter_nearest_wrap"] = (lambda: W.random_mesh(31, 300, filter=A.FILTER_NEAREST, unknown_state_promotion=A.PROMOTE_NEAREST, uv_lo=-0.5, uv_hi=1.5), {})
    c["filter_nearest_fo_npot"] = (lambda: W.random_mesh(32, 300, tex_size=(90, 120), filter=A.FILTER_NEAREST, addressing_mode=A.ADDR_MIRROR), {})
    c["filter_nearest_mips"] = (lambda: W.random_mesh(33, 200, mips=4, filter=A.FILTER_NEAREST, unknown_state_promotion=A.PROMOTE_NEAREST), {})
    c["unorm8_blocky"] = (lambda: W.random_mesh(41, 300, tex_kind="blocky", unorm8=True, unknown_state_promotion=A.PROMOTE_NEAREST), {})
    c["mips_linear"] = (lambda: W.random_mesh(42, 300, mips=5, tex_kind="circle", unknown_state_promotion=A.PROMOTE_NEAREST), {})
    c["mips_linear_2state_fo"] = (lambda: W.random_mesh(43, 300, mips=3, format=A.FORMAT_2_STATE), {})
    c["sat_clamp"] = (lambda: W.random_mesh(44, 400, tex_kind="blocky", tex_alpha_cutoff=0.5, addressing_mode=A.ADDR_CLAMP, tri_texels=20, max_subdivision_level=5), {})
    c["sat_wrap_outside"] = (lambda: W.random_mesh(45, 400, tex_kind="blocky", tex_alpha_cutoff=0.5, uv_lo=-1.0, uv_hi=2.0, tri_texels=20, max_subdivision_level=5), {})
    # SAT pass where the address mapping folds inside a micro-triangle's texel range (Mirror / MirrorOnce across a mirror axis): the SDK's
    # SAT rectangle is then not the set of texels the micro-triangle uses, and only its own evaluation order reproduces the result
    # (omm_hier.cuh (S); found by the host campaign, see tests/test_hier_host.py::test_sat_pass_where_the_address_mapping_folds)
    c["sat_mirror_once_outside"] = (lambda: W.random_mesh(47, 500, tex_kind="noise", unorm8=True, tex_alpha_cutoff=0.5, uv_lo=-0.6, uv_hi=1.6, tri_texels=14,
                                                          addressing_mode=A.ADDR_MIRROR_ONCE, max_subdivision_level=5), {})
    c["sat_mirror_outside"] = (lambda: W.random_mesh(48, 500, tex_kind="noise", tex_alpha_cutoff=0.5, uv_lo=-1.2, uv_hi=2.2, tri_texels=14,
                                                     addressing_mode=A.ADDR_MIRROR, max_subdivision_level=5, unknown_state_promotion=A.PROMOTE_NEAREST), {})
    c["sat_mirror_once_axis"] = (lambda: W.random_mesh(49, 300, tex_size=(128, 128), tex_kind="noise", unorm8=True, tex_alpha_cutoff=0.5, uv_lo=-0.08, uv_hi=0.08,
                                                       tri_texels=9, addressing_mode=A.ADDR_MIRROR_ONCE, max_subdivision_level=4, format=A.FORMAT_2_STATE), {})
    c["sat_wrap_tiny_texture"] = (lambda: W.random_mesh(50, 200, tex_size=(8, 8), tex_kind="noise", tex_alpha_cutoff=0.5, uv_lo=-1.0, uv_hi=2.0, tri_texels=6,
                                                        max_subdivision_level=2), {})
    c["sat_disable_zorder"] = (lambda: W.random_mesh(46, 200, tex_kind="circle", tex_alpha_cutoff=0.5, tex_flags=A.TEXFLAG_DISABLE_ZORDER), {})
    c["degenerate_and_nan"] = (lambda: W.random_mesh(51, 400, degenerate_frac=0.3, nan_frac=0.1, tri_texels=30), {})
    c["all_triangles_invalid"] = (lambda: W.random_mesh(59, 64, nan_frac=1.0, bake_flags=A.BAKE_DISABLE_SPECIAL_INDICES | A.BAKE_DISABLE_DUPLICATE_DETECTION), {})
    c["degenerate_nearest_promo"] = (lambda: W.random_mesh(52, 300, degenerate_frac=0.5, tri_texels=40, unknown_state_promotion=A.PROMOTE_NEAREST,
                                                           unresolved_tri_state=A.SPECIAL_FUT, nan_frac=0.05), {})
    c["degenerate_dynamic_levels"] = (lambda: W.random_mesh(53, 300, degenerate_frac=0.4, tri_texels=50, dynamic_subdivision_scale=2.0,
                                                            max_subdivision_level=6), {})
    c["dynamic_levels_area"] = (lambda: W.random_mesh(54, 400, tri_texels=40, dynamic_subdivision_scale=1.5, max_subdivision_level=7), {})
    c["per_triangle_levels"] = (lambda: W.random_mesh(55, 500, subdivision_levels=_levels(55, 500, 0, 6), unknown_state_promotion=A.PROMOTE_NEAREST), {})
    c["per_triangle_levels_use_global"] = (lambda: W.random_mesh(56, 300, subdivision_levels=np.where(np.arange(300) % 3 == 0, 13, _levels(56, 300, 0, 5)).astype(np.uint8),
                                                                 max_subdivision_level=2), {})
    c["big_microtriangles_l0_l2"] = (lambda: W.random_mesh(57, 60, tri_texels=120, subdivision_levels=_levels(57, 60, 0, 2),
                                                           unknown_state_promotion=A.PROMOTE_NEAREST, uv_lo=0.3, uv_hi=0.7), {})
    c["reuse_uv_and_content"] = (lambda: W.random_mesh(58, 600, reuse_frac=0.5, tex_kind="blocky", tri_texels=6, max_subdivision_level=2), {})
    c["flag_disable_special"] = (lambda: W.random_mesh(61, 300, tex_kind="blocky", tri_texels=5, bake_flags=A.BAKE_DISABLE_SPECIAL_INDICES), {})
    c["flag_disable_special_levels"] = (lambda: W.random_mesh(74, 400, tex_kind="blocky", tri_texels=4, subdivision_levels=_levels(74, 400, 0, 6),
                                                              bake_flags=A.BAKE_DISABLE_SPECIAL_INDICES), {})
    # large triangles at low levels over a texture with big constant areas (constant-cell table (H) of the hierarchical classifier)
    c["flat_areas_big_triangles"] = (lambda: W.config5(num_tris=1500, tex_size=256, distinct=128, flat_tris=700, max_level=6), {})
    c["flat_areas_clamp_outside"] = (lambda: W.random_mesh(75, 200, tex_kind="blocky", tri_texels=90, uv_lo=-0.4, uv_hi=1.4, addressing_mode=A.ADDR_CLAMP,
                                                           subdivision_levels=_levels(75, 200, 0, 4)), {})
    c["flag_disable_dup"] = (lambda: W.random_mesh(62, 300, reuse_frac=0.5, tex_kind="blocky", bake_flags=A.BAKE_DISABLE_DUPLICATE_DETECTION), {})
    c["flag_force32"] = (lambda: W.random_mesh(63, 300, bake_flags=A.BAKE_FORCE_32BIT_INDICES), {})
    c["flag_allow8"] = (lambda: W.random_mesh(64, 100, bake_flags=A.BAKE_ALLOW_8BIT_INDICES), {})
    c["flag_allow8_too_many"] = (lambda: W.random_mesh(65, 200, bake_flags=A.BAKE_ALLOW_8BIT_INDICES), {})
    c["rejection_threshold"] = (lambda: W.random_mesh(66, 300, rejection_threshold=0.6), {})
    c["states_swapped"] = (lambda: W.random_mesh(67, 300, alpha_cutoff_gt=A.STATE_T, alpha_cutoff_le=A.STATE_O, unknown_state_promotion=A.PROMOTE_NEAREST), {})
    c["states_le_unknown_opaque"] = (lambda: W.random_mesh(68, 300, alpha_cutoff_le=A.STATE_UO, alpha_cutoff_gt=A.STATE_T, mips=3), {})
    c["cutoff_03"] = (lambda: W.random_mesh(69, 300, alpha_cutoff=0.3, tex_kind="circle"), {})
    c["index_u16"] = (lambda: W.random_mesh(71, 300, index_dtype=np.uint16), {})
    c["index_u8"] = (lambda: W.random_mesh(72, 80, index_dtype=np.uint8), {})
    c["tri_count_33000_u16_output_32"] = (lambda: W.random_mesh(73, 33000, tri_texels=3, max_subdivision_level=1), {})
    c["internal_disable_level_line"] = (lambda: W.random_mesh(81, 200, bake_flags=A.BAKE_INT_DISABLE_LEVEL_LINE, degenerate_frac=0.2), {})
    c["internal_aabb_testing"] = (lambda: W.random_mesh(82, 200, bake_flags=A.BAKE_INT_DISABLE_LEVEL_LINE | A.BAKE_INT_AABB_TESTING), {})
    c["internal_disable_fine_sat"] = (lambda: W.random_mesh(83, 200, tex_kind="blocky", tex_alpha_cutoff=0.5, bake_flags=A.BAKE_INT_DISABLE_FINE, tri_texels=20), {})
    c["internal_edge_heuristic"] = (lambda: W.random_mesh(84, 200, tri_texels=40, dynamic_subdivision_scale=2.0, max_subdivision_level=6,
                                                          bake_flags=A.BAKE_INT_EDGE_HEURISTIC), {})
    return c


def sdk_only_cases():
    """Cases for the optional passes the plain-C port does not restate (near-duplicate merge a17, Compress a18): checked
    against the SDK build only (and its golden digests)."""
    c = {}
    hexa = lambda **kw: W.random_mesh(101, 700, tex_kind="blocky", tri_texels=14, max_subdivision_level=3, reuse_frac=0.1, **kw)
    c["neardup_lsh_blocky"] = (lambda: hexa(bake_flags=A.BAKE_ENABLE_NEAR_DUPLICATE_DETECTION), {})
    c["neardup_lsh_noise_l4"] = (lambda: W.random_mesh(102, 500, tri_texels=10, max_subdivision_level=4, bake_flags=A.BAKE_ENABLE_NEAR_DUPLICATE_DETECTION,
                                                       near_duplicate_factor=0.25), {})
    c["neardup_lsh_mixed_levels"] = (lambda: W.random_mesh(103, 600, tex_kind="circle", tri_texels=30, subdivision_levels=_levels(103, 600, 0, 5),
                                                           bake_flags=A.BAKE_ENABLE_NEAR_DUPLICATE_DETECTION, unknown_state_promotion=A.PROMOTE_NEAREST), {})
    c["neardup_bruteforce"] = (lambda: hexa(bake_flags=A.BAKE_ENABLE_NEAR_DUPLICATE_DETECTION | A.BAKE_INT_NEAR_DUP_BRUTE_FORCE), {})
    c["neardup_lsh_rejection"] = (lambda: hexa(bake_flags=A.BAKE_ENABLE_NEAR_DUPLICATE_DETECTION, rejection_threshold=0.5), {})
    c["compress_budget_half"] = (lambda: W.random_mesh(104, 300, tri_texels=16, max_subdivision_level=4, max_array_data_size=8000), {})
    c["compress_budget_tiny"] = (lambda: W.random_mesh(105, 200, tri_texels=16, max_subdivision_level=5, max_array_data_size=600), {})
    c["compress_budget_not_binding"] = (lambda: W.random_mesh(106, 100, max_subdivision_level=3, max_array_data_size=1 << 20), {})
    c["compress_and_neardup"] = (lambda: W.random_mesh(107, 400, tex_kind="blocky", tri_texels=14, max_subdivision_level=4, max_array_data_size=5000,
                                                       bake_flags=A.BAKE_ENABLE_NEAR_DUPLICATE_DETECTION), {})
    c["compress_disable_dup"] = (lambda: W.random_mesh(108, 200, tri_texels=16, max_subdivision_level=4, max_array_data_size=4000,
                                                       bake_flags=A.BAKE_DISABLE_DUPLICATE_DETECTION), {})
    return c


def uv_format_cases():
    """UV16_UNORM / UV16_FLOAT / strided UV32 variants of one mesh."""
    out = {}
    base = W.random_mesh(91, 300, uv_lo=0.05, uv_hi=0.95)
    uv = base.texcoords
    wl = W.random_mesh(91, 300, uv_lo=0.05, uv_hi=0.95)
    wl.texcoords, wl.texcoord_format = W.pack_unorm16(uv), A.UV16_UNORM
    out["uv16_unorm"] = wl
    wl = W.random_mesh(91, 300, uv_lo=0.05, uv_hi=0.95)
    wl.texcoords, wl.texcoord_format = W.pack_half(uv), A.UV16_FLOAT
    out["uv16_float"] = wl
    wl = W.random_mesh(91, 300, uv_lo=0.05, uv_hi=0.95)
    padded = np.zeros((uv.shape[0], 5), dtype=np.float32)
    padded[:, :2] = uv
    wl.texcoords = padded
    wl.desc["texcoord_stride"] = 20
    out["uv32_stride20"] = wl
    return out


def run_bake(lib, wl, **overrides):
    from omm_b200 import Baker
    msgs = []
    with Baker(lib, on_message=lambda sev, m: msgs.append((sev, m))) as b:
        inp, tex = W.make_input(b, wl, **overrides)
        try:
            return b.bake(inp)
        except Exception as e:
            raise RuntimeError(f"{e}; messages: {msgs}") from e
        finally:
            tex.destroy()
