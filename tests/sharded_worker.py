"""Worker for the sharded-bake parity test (launched with torch.distributed.run, one rank per GPU): every rank bakes
the same inputs with work items sharded over the ranks and must end up with the byte-identical full result."""
import ctypes as C
import hashlib
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from omm_b200 import Baker, capi, load_product_library  # noqa: E402
from omm_b200 import workloads as W  # noqa: E402
import parity_cases as PC  # noqa: E402


def digest(res):
    h = hashlib.sha256()
    for k in ("array_data", "desc_array", "desc_histogram", "index_buffer", "index_histogram"):
        h.update(getattr(res, k).tobytes())
    h.update(bytes([res.index_format]))
    return h.hexdigest()


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = load_product_library()
    assert lib.dll.ommB200SetDevice(local) == capi.SUCCESS
    cases = {
        "c3_l6": W.config3(num_tris=3000, tex_size=512, level=6),
        "c5_mixed": W.config5(num_tris=4000, tex_size=256, distinct=300, flat_tris=600, max_level=8),
        "c2": W.config2(num_quads=400, tex_size=256, level=4),
        "tiny_levels": PC.cases()["per_triangle_levels"][0](),
        # the flat classifier (Nearest filter) and the host passes (every state block is exchanged) over several shards per rank
        "c3_nearest": (W.config3(num_tris=1500, tex_size=256, level=5), dict(filter=capi.FILTER_NEAREST)),
        "c3_near_duplicates": (W.config3(num_tris=600, tex_size=128, level=4), dict(bake_flags=capi.BAKE_ENABLE_NEAR_DUPLICATE_DETECTION)),
        "three_items": W.config3(num_tris=3, tex_size=64, level=3),  # fewer work items than shards
        # the complete array on rank 0 only: gathered over NVLink (resident) / assembled in the shared page-locked window (ommCpuBake, > 1 MiB)
        "c3_on_rank0": (W.config3(num_tris=9000, tex_size=1024, level=6), dict(), "rank0"),
    }
    out = {}
    for name, wl in cases.items():
        over, mode = {}, "replicated"
        if isinstance(wl, tuple):
            mode = wl[2] if len(wl) > 2 else "replicated"
            wl, over = wl[0], wl[1]
        # single-GPU result on this rank
        with Baker(lib) as b:
            inp, tex = W.make_input(b, wl, **over)
            single = b.bake(inp)
            tex.destroy()
        # sharded result
        b = Baker(lib)
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (C.c_uint8 * 128)()
            assert lib.dll.ommB200GetNcclUniqueId(raw, 128) == capi.SUCCESS
            idbuf = torch.tensor(list(raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(idbuf, 0)
        raw = (C.c_uint8 * 128)(*idbuf.cpu().tolist())
        assert lib.dll.ommB200InitSharding(b.handle, rank, world, raw, 128) == capi.SUCCESS
        inp, tex = W.make_input(b, wl, **over)
        if mode == "rank0":
            assert lib.dll.ommB200SetShardedResultMode(b.handle, capi.SHARDED_RESULT_ON_RANK0) == capi.SUCCESS
            desc = inp.to_desc()
            # ommCpuBake: host copy assembled by all ranks; complete on rank 0, refused elsewhere
            rc, h = b.bake_raw(desc)
            assert rc == capi.SUCCESS
            pdesc = C.POINTER(capi.CpuBakeResultDesc)()
            rc = lib.dll.ommCpuGetBakeResultDesc(h, C.byref(pdesc))
            if rank == 0:
                assert rc == capi.SUCCESS
                from omm_b200.baker import _copy_result
                assert _copy_result(pdesc.contents).diff(single) == [], f"{name}: host copy on rank 0 differs from the single-GPU result"
            else:
                assert rc == capi.INVALID_ARGUMENT
            lib.dll.ommCpuDestroyBakeResult(h)
            # ommB200BakeResident: gathered into rank 0's HBM
            staged = C.c_void_p()
            assert lib.dll.ommB200StageInputs(b.handle, C.byref(desc), C.byref(staged)) == capi.SUCCESS
            h = C.c_void_p()
            assert lib.dll.ommB200BakeResident(b.handle, staged, None, C.byref(h)) == capi.SUCCESS
            rc = lib.dll.ommB200DownloadResult(h)
            assert rc == (capi.SUCCESS if rank == 0 else capi.INVALID_ARGUMENT)
            if rank == 0:
                assert lib.dll.ommCpuGetBakeResultDesc(h, C.byref(pdesc)) == capi.SUCCESS
                sharded = _copy_result(pdesc.contents)
                assert sharded.diff(single) == [], f"{name}: resident result gathered on rank 0 differs from the single-GPU result"
            else:
                sharded = single
            t2 = capi.B200BakeTimings()
            lib.dll.ommB200GetLastBakeTimings(b.handle, C.byref(t2))
            mine = t2.microTriangles
            lib.dll.ommCpuDestroyBakeResult(h)
            lib.dll.ommB200DestroyStagedInputs(staged)
        else:
            sharded = b.bake(inp)
            mine = sharded.timings.microTriangles
        tex.destroy()
        b.destroy()
        assert sharded.diff(single) == [], f"rank {rank} {name}: sharded result differs from the single-GPU result"
        t = torch.tensor([mine], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        out[name] = {"digest": digest(sharded), "micro_triangles_all_ranks": int(t.item()), "micro_triangles_single": int(single.timings.microTriangles)}
        assert out[name]["micro_triangles_all_ranks"] == out[name]["micro_triangles_single"], out[name]
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    for g in gathered[1:]:
        assert g == gathered[0], "ranks disagree"
    if rank == 0:
        print("SHARDED_OK " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
