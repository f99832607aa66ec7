#!/usr/bin/env python
"""bench.py -- micro-triangles classified per second on BASELINE config 3 (1 M triangles, 4096^2 alpha, level 6, 4-state).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one complete bake of the synthetic config-3 input: UV pre-dedup, classification of
4.096e9 micro-triangles, special-index scan, XXH64 dedup, spatial sort, bit-pack, index buffer.

  value   device-resident: inputs staged in HBM once (ommB200StageInputs), each step = ommB200BakeResident on the
          current CUDA stream, timed with CUDA events on that stream, result left in HBM.
  e2e     the drop-in call: ommCpuBake() with host input buffers (pinned) -> host result arrays; wall clock around the C
          call, so the host->device and device->host copies are inside the timed region.
  N > 1   strong scaling of the same 1 M-triangle bake: work items sharded over ranks, one NCCL all-gather of the state
          blocks, merge replicated (launched with torch.distributed.run, one rank per GPU).  In the e2e arm every rank stages
          the inputs and takes part in the bake; the host copy of the result is read on rank 0.
  --impl reference   the SDK's own CPU baker (oracle/_ref/libomm-lib.so, OpenMP, all host cores) on a bounded slice of
          the same workload per step (rank 0 only), sized so that the K + W steps take about --ref-budget-s seconds.
  parity  the metric says "bit-exact vs CPU baker", so the run proves it: at N=1 the cpu_baseline leg bakes the FULL config once with the
          SDK build and the five result arrays of the GPU bake are compared with it byte for byte; at every N the sha256 of the result
          (all ranks) is printed and compared with the committed digest of the SDK's result (tests/golden/full_size_digests.json), so
          the N = 2 / 4 / 8 lines show the same digest as N = 1.
  secondary  BASELINE configs 2 and 5 (config.secondary): device-resident and end-to-end times, digest against the SDK's.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from omm_b200 import capi  # noqa: E402
from omm_b200 import workloads as W  # noqa: E402
from omm_b200.baker import Baker  # noqa: E402

METRIC = "micro-triangles classified/sec at 1M tris subdiv-6; bit-exact vs CPU baker"
UNIT = "micro-triangles/s"
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libomm-lib.so")
PORT_LIB = os.path.join(ROOT, "oracle", "liboracle_port.so")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tris", type=int, default=1_000_000)
    ap.add_argument("--tex", type=int, default=4096)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--cpu-sample-tris", type=int, default=0, help="triangles in the bounded CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=float(os.environ.get("OMM_BENCH_CPU_BUDGET_S", "330")),
                    help="the full-config SDK bake of the cpu_baseline leg is replaced by a slice when a probe predicts it would take longer than this")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="--impl reference: target duration of all K + W steps together")
    ap.add_argument("--no-secondary", action="store_true", help="skip BASELINE configs 2 and 5")
    ap.add_argument("--result-mode", default="rank0", choices=["rank0", "replicated"],
                    help="N > 1: where the complete arrayData ends up (ommB200SetShardedResultMode); rank0 = in one GPU's HBM / one host buffer, like at N = 1")
    return ap.parse_args()


def workload_name(a):
    return f"C3: {a.tris} unindexed triangles on a jittered {max(2, round(708 * a.tex / 4096))}^2 cell grid, {a.tex}x{a.tex} FP32 2-octave value noise, " \
           f"Linear/Wrap, cutoff 0.5, level {a.level}, OC1_4_State, ForceOpaque"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi sampler (B200_PROFILING.md recipe) running for the duration of the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines, self.skip = gpu_index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first(self, timeout_s: float = 2.0):
        """Block until nvidia-smi delivers its first line: its start-up (NVML enumerates every GPU of the box) was seen to stall CUDA calls
        of all ranks for ~20 ms, so it is started during the last warm-up step and only loops (-lms) through the timed region."""
        t0 = time.time()
        while self.proc and not self.lines and time.time() - t0 < timeout_s:
            time.sleep(0.01)

    def mark(self):
        """Samples delivered before this call (warm-up) are not reported."""
        self.skip = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[self.skip:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def pinned_like(arr: np.ndarray):
    """Copy a numpy array into page-locked host memory (torch is only the allocator here)."""
    import torch
    t = torch.empty(arr.nbytes, dtype=torch.uint8, pin_memory=True)
    view = t.numpy().view(arr.dtype).reshape(arr.shape)
    view[...] = arr
    return view, t


# ---------------------------------------------------------------------------------------------------------------------
def golden_digests():
    p = os.path.join(ROOT, "tests", "golden", "full_size_digests.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def is_full_c3(a):
    return a.tris == 1_000_000 and a.tex == 4096 and a.level == 6


def cpu_library():
    """(lib, kind, threads) of the CPU checker: the unmodified SDK build when it travelled with the repo, else the scalar port."""
    cores = os.cpu_count() or 1
    if os.path.exists(REF_LIB):
        os.environ["OMP_NUM_THREADS"] = str(cores)
        os.environ.setdefault("OMP_PROC_BIND", "spread")
        return capi.OmmLib(REF_LIB), "reference", cores
    if not os.path.exists(PORT_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    return capi.OmmLib(PORT_LIB), "port", 1


def cpu_bake(lib, wl, keep_result=False):
    """One ommCpuBake on the CPU library; returns (seconds of the ommCpuBake call, BakeResult or None)."""
    from omm_b200.baker import _copy_result
    with Baker(lib) as b:
        inp, tex = W.make_input(b, wl, bake_flags=capi.BAKE_ENABLE_INTERNAL_THREADS)
        desc = inp.to_desc()
        t0 = time.perf_counter()
        rc, h = b.bake_raw(desc)
        dt = time.perf_counter() - t0
        assert rc == capi.SUCCESS, rc
        res = None
        if keep_result:
            pdesc = C.POINTER(capi.CpuBakeResultDesc)()
            assert lib.dll.ommCpuGetBakeResultDesc(h, C.byref(pdesc)) == capi.SUCCESS
            res = _copy_result(pdesc.contents)
        lib.dll.ommCpuDestroyBakeResult(h)
        tex.destroy()
    return dt, res


def run_reference(a):
    """The SDK's CPU baker on a bounded slice of the same workload per step, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib, kind, used = cpu_library()
    # probe: a small slice gives the rate; the per-step sample is sized so that all K + W steps fit the budget
    probe = min(a.tris, 2048 if kind == "reference" else 128)
    dt_probe, _ = cpu_bake(lib, W.config3(num_tris=probe, tex_size=a.tex, level=a.level))
    per_tri = dt_probe / probe
    sample = a.cpu_sample_tris or int(a.ref_budget_s / ((a.steps + a.warmup) * per_tri))
    sample = max(256, min(sample, a.tris))
    wl = W.config3(num_tris=sample, tex_size=a.tex, level=a.level)
    utris = sample * 4 ** a.level
    times = []
    with Baker(lib) as b:
        inp, tex = W.make_input(b, wl, bake_flags=capi.BAKE_ENABLE_INTERNAL_THREADS)
        desc = inp.to_desc()
        for it in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            rc, h = b.bake_raw(desc)
            dt = time.perf_counter() - t0
            assert rc == capi.SUCCESS, rc
            lib.dll.ommCpuDestroyBakeResult(h)
            if it >= a.warmup:
                times.append(dt)
        tex.destroy()
    total = sum(times)
    value = utris * len(times) / total
    sample_txt = (f"all {a.tris} triangles" if sample == a.tris else f"first {sample} of {a.tris} triangles of the same grid/texture") + \
                 f" ({utris:.3e} micro-triangles) per step, one ommCpuBake call each"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample_txt,
                   "note": "rate per micro-triangle; the b200 arm's cpu_baseline leg times the FULL config once on the same cores (and compares the bytes)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def cpu_baseline(a, gpu_result):
    """The SDK build on the FULL config (one ommCpuBake call on all host cores), and the byte comparison with the GPU result.  When a probe predicts
    more than --cpu-budget-s the largest slice that fits is timed instead and parity rests on the committed digest of the SDK's full result."""
    from omm_b200.baker import result_sha256
    lib, kind, used = cpu_library()
    probe = min(a.tris, 4096 if kind == "reference" else 128)
    dt_probe, _ = cpu_bake(lib, W.config3(num_tris=probe, tex_size=a.tex, level=a.level))
    predicted = dt_probe / probe * a.tris
    sample = a.cpu_sample_tris or (a.tris if predicted <= a.cpu_budget_s else max(256, int(a.tris * a.cpu_budget_s / predicted)))
    sample = min(sample, a.tris)
    full = sample == a.tris
    dt, res = cpu_bake(lib, W.config3(num_tris=sample, tex_size=a.tex, level=a.level), keep_result=full)
    utris = sample * 4 ** a.level
    base = {"value": utris / dt, "unit": UNIT, "cores": used, "kind": kind, "seconds": dt,
            "sample": (f"FULL config: all {a.tris} triangles" if full else f"first {sample} of {a.tris} triangles of the same grid/texture") +
                      f", one ommCpuBake call ({utris:.3e} micro-triangles)"}
    parity = None
    if full and gpu_result is not None:
        d = gpu_result.diff(res)
        parity = {"arrays_identical": d == [], "compared": "arrayData, descArray, descArrayHistogram, indexBuffer (+ format), indexHistogram of the GPU ommCpuBake result "
                  f"vs one full ommCpuBake of the {kind} library in this run", "oracle_sha256": result_sha256(res), "diff": d}
    return base, parity


# ---------------------------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    from omm_b200.baker import _copy_result, result_sha256

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    lib = capi.load_product_library()
    assert lib.dll.ommB200GetDeviceCount() > 0, "no CUDA device visible"
    assert lib.dll.ommB200SetDevice(local) == capi.SUCCESS

    baker = Baker(lib, on_message=lambda sev, msg: print(f"[omm-b200 message, rank {rank}, severity {sev}] {msg}", file=sys.stderr, flush=True))
    if world > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_uint8 * 128)()
            assert lib.dll.ommB200GetNcclUniqueId(raw, 128) == capi.SUCCESS
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        idbuf = idbuf.cuda()
        dist.broadcast(idbuf, 0)
        raw = (C.c_uint8 * 128)(*idbuf.cpu().tolist())
        assert lib.dll.ommB200InitSharding(baker.handle, rank, world, raw, 128) == capi.SUCCESS
        assert lib.dll.ommB200SetShardedResultMode(baker.handle, capi.SHARDED_RESULT_ON_RANK0 if a.result_mode == "rank0" else capi.SHARDED_RESULT_REPLICATED) == capi.SUCCESS
    everywhere = world == 1 or a.result_mode == "replicated"   # the complete result exists on every rank
    stream = torch.cuda.current_stream()
    stream_ptr = C.c_void_p(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    tm = capi.B200BakeTimings()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def resident_steps(desc, warmup, steps, sampler=None):
        """W + K device-resident bakes of staged inputs; returns per-step ms (max over ranks), per-step classify ms, launches, timings of the last."""
        staged = C.c_void_p()
        assert lib.dll.ommB200StageInputs(baker.handle, C.byref(desc), C.byref(staged)) == capi.SUCCESS
        step_ms, classify_ms, launches, last = [], [], 0, None
        for it in range(warmup + steps):
            flush.fill_(it & 0xFF)  # evict L2 between steps (outside the timed events)
            barrier()
            if sampler is not None and rank == 0 and it == 0:
                # started with the FIRST warm-up step: nvidia-smi's start-up (NVML enumerates every GPU of the box) stalls CUDA calls for tens of
                # milliseconds and was seen to leak into the first timed step when it was started during the last warm-up step
                sampler.start()
                sampler.wait_first(5.0)
            if sampler is not None and it == warmup:
                sampler.mark()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            h = C.c_void_p()
            rc = lib.dll.ommB200BakeResident(baker.handle, staged, stream_ptr, C.byref(h))
            e1.record(stream)
            assert rc == capi.SUCCESS, f"ommB200BakeResident -> {rc}"
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1))
            lib.dll.ommB200GetLastBakeTimings(baker.handle, C.byref(tm))
            if it >= warmup:
                step_ms.append(ms)
                classify_ms.append(tm.classifyMs)
                launches += tm.kernelLaunches
            last = dict(work_items=tm.workItems, array_bytes=tm.arrayDataBytes, desc_count=tm.descCount, my_utris=tm.microTriangles, setup_ms=tm.setupMs,
                        post_ms=tm.postMs, item_post_ms=tm.itemPostMs, gather_ms=tm.gatherMs)
            lib.dll.ommCpuDestroyBakeResult(h)
        lib.dll.ommB200DestroyStagedInputs(staged)
        return step_ms, classify_ms, launches, last

    def e2e_steps(desc, warmup, steps, keep_last=False, every_rank_downloads=False):
        """W + K drop-in calls ommCpuBake + ommCpuGetBakeResultDesc (host buffers in, host arrays out); wall seconds per step (max over ranks)."""
        secs, launches, info, kept = [], 0, None, None
        for it in range(warmup + steps):
            barrier()
            t0 = time.perf_counter()
            h = C.c_void_p()
            rc = lib.dll.ommCpuBake(baker.handle, C.byref(desc), C.byref(h))
            pdesc = C.POINTER(capi.CpuBakeResultDesc)()
            # the host copy of the result is materialised where it is consumed: on rank 0
            wants = rank == 0 or (every_rank_downloads and everywhere)
            rc2 = lib.dll.ommCpuGetBakeResultDesc(h, C.byref(pdesc)) if wants else capi.SUCCESS
            dt = time.perf_counter() - t0
            assert rc == capi.SUCCESS and rc2 == capi.SUCCESS, (rc, rc2)
            dt = max_over_ranks(dt)
            lib.dll.ommB200GetLastBakeTimings(baker.handle, C.byref(tm))
            info = {"h2d": int(tm.h2dBytes), "d2h": int(tm.d2hBytes),
                    "breakdown": {"stage_ms": tm.hostStageMs, "bake_ms": tm.hostBakeMs, "download_ms": tm.hostDownloadMs, "h2d_ms": tm.h2dMs, "d2h_ms": tm.d2hMs,
                                  "device_total_ms": tm.totalDeviceMs}}
            if it >= warmup:
                secs.append(dt)
                launches += tm.kernelLaunches
            if keep_last and it == warmup + steps - 1 and wants:
                kept = _copy_result(pdesc.contents)
            lib.dll.ommCpuDestroyBakeResult(h)
        return secs, launches, info, kept

    # ================= headline: BASELINE config 3 =================
    wl = W.config3(num_tris=a.tris, tex_size=a.tex, level=a.level)
    utris_total = a.tris * 4 ** a.level
    pageable_idx, pageable_uv = wl.indices, wl.texcoords
    idx_pinned, _k1 = pinned_like(wl.indices)
    uv_pinned, _k2 = pinned_like(wl.texcoords)
    wl.indices, wl.texcoords = idx_pinned, uv_pinned
    inp, tex = W.make_input(baker, wl)
    desc = inp.to_desc()

    sampler = ClockSampler(local)
    step_ms, classify_ms, launches, last = resident_steps(desc, a.warmup, a.steps, sampler)
    clocks = sampler.stop()
    total_ms = sum(step_ms)
    value = utris_total * len(step_ms) / (total_ms * 1e-3)

    # ---- end-to-end arm: the drop-in ommCpuBake with host buffers (page-locked by the caller) ----
    e2e_s, l2, e2e_info, _ = e2e_steps(desc, a.warmup, a.steps)
    launches += l2
    e2e_value = utris_total * len(e2e_s) / sum(e2e_s)
    # ---- the same call with PAGEABLE caller buffers (what a stock SDK caller passes) ----
    wl.indices, wl.texcoords = pageable_idx, pageable_uv
    inp_pg = BakeInputFrom(inp, pageable_idx, pageable_uv)
    desc_pg = inp_pg.to_desc()
    pg_s, _, _, _ = e2e_steps(desc_pg, 1, max(2, min(3, a.steps)))
    # ---- verification bake (outside every timed region): every rank downloads its copy of the result; digests must agree ----
    _, _, _, mine = e2e_steps(desc_pg, 0, 1, keep_last=True, every_rank_downloads=True)
    my_sha = result_sha256(mine) if mine is not None else None
    shas = [my_sha]
    if dist is not None:
        shas = [None] * world
        dist.all_gather_object(shas, my_sha)
        shas = [x for x in shas if x is not None]
    golden = golden_digests()
    parity = {"config": "C3 full" if is_full_c3(a) else f"C3 variant ({a.tris} triangles, {a.tex}^2, level {a.level})", "result_sha256": shas[0],
              "ranks_identical": len(set(shas)) == 1, "ranks": world, "ranks_holding_the_result": len(shas)}
    if is_full_c3(a) and "C3" in golden:
        parity["golden_sha256"] = golden["C3"]["sha256"]
        parity["matches_golden"] = shas[0] == golden["C3"]["sha256"] and len(set(shas)) == 1
        parity["golden_source"] = "tests/golden/full_size_digests.json: sha256 of the unmodified SDK build's result for this config (tests/golden/make_full_size_digests.py)"

    # ================= secondary: BASELINE configs 2 and 5 =================
    secondary = {}
    if not a.no_secondary and is_full_c3(a):
        tex.destroy()
        tex = None
        for name, swl in (("C2", W.config2()), ("C5", W.config5())):
            sinp, stex = W.make_input(baker, swl)
            sdesc = sinp.to_desc()
            sms, scl, sl, slast = resident_steps(sdesc, 2, 3)
            ss, sl2, _, sres = e2e_steps(sdesc, 1, 3, keep_last=True, every_rank_downloads=True)
            launches += sl + sl2
            sha = result_sha256(sres) if sres is not None else None
            sshas = [sha]
            if dist is not None:
                sshas = [None] * world
                dist.all_gather_object(sshas, sha)
                sshas = [x for x in sshas if x is not None]
            sutris = int(slast["my_utris"]) if world == 1 else None
            entry = {"workload": swl.name, "ms_per_step": statistics.median(sms), "classify_ms": statistics.median(scl), "item_post_ms": slast["item_post_ms"],
                     "post_ms": slast["post_ms"], "setup_ms": slast["setup_ms"], "gather_ms": slast["gather_ms"], "e2e_ms_per_step": 1e3 * statistics.median(ss),
                     "work_items": slast["work_items"], "array_data_bytes": slast["array_bytes"], "result_sha256": sshas[0], "ranks_identical": len(set(sshas)) == 1}
            if sutris:
                entry["micro_triangles"] = sutris
                entry["micro_triangles_per_s"] = sutris / (statistics.median(sms) * 1e-3)
            if name in golden:
                entry["matches_golden"] = sshas[0] == golden[name]["sha256"] and len(set(sshas)) == 1
            if name == "C2" and rank == 0 and world == 1 and not a.no_cpu_baseline:
                clib, ckind, _ = cpu_library()
                cdt, cres = cpu_bake(clib, swl, keep_result=True)
                entry["arrays_identical"] = sres.diff(cres) == []
                entry["cpu_seconds"] = cdt
            secondary[name] = entry
            stex.destroy()

    # ---- roofline: B_alg = texture + geometry + the outputs the result really holds; duration = the whole device-resident bake ----
    peaks, peak_kind = measured_peaks()
    tex_bytes = a.tex * a.tex * 4
    b_alg = tex_bytes + wl.indices.nbytes + wl.texcoords.nbytes + int(last["array_bytes"]) + 8 * int(last["desc_count"]) + 4 * a.tris
    ms_step = total_ms / len(step_ms)
    cls_ms = sum(classify_ms) / len(classify_ms)
    achieved = b_alg / (ms_step * 1e-3) / 1e9
    traffic, issue = None, None
    if world == 1 and is_full_c3(a):
        tj = latest_profile_json("stage_traffic")
        if tj is not None:
            traffic = tj["dram_bytes_read_per_bake"] + tj["dram_bytes_written_per_bake"]
        issue = latest_profile_json("issue")
    roofline = {"bound": "hbm", "kernel": "whole bake (classification stage = the Hier* kernels is %.0f %% of it)" % (100.0 * cls_ms / ms_step),
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "algorithmic_bytes": b_alg, "algorithmic_bytes_def": "SURVEY 8d B_alg = texture + index buffer + UV buffer + arrayData + 8*descs + 4*triangles",
                "kernel_ms": ms_step, "classify_stage_ms": cls_ms, "frac_classify_stage_only": b_alg / (cls_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "traffic": traffic, "traffic_ratio": (traffic / b_alg) if traffic else None, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                "issue": issue,
                "note": "the path is instruction-issue bound, not HBM bound: B_alg is 0.1 byte per micro-triangle against the bit-exact level-line arithmetic (IEEE divisions, "
                        "square roots) of the micro-triangles the level line touches; `issue` holds the ncu counters of the dominant kernels of this build (profiles/)"}

    if rank == 0:
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            cpu, full_parity = cpu_baseline(a, mine)
            if full_parity is not None:
                parity.update(full_parity)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "l2": "256 MiB buffer written between steps (outside the timed events)",
                       "work_items": last["work_items"], "array_data_bytes": last["array_bytes"], "desc_count": last["desc_count"],
                       "step_ms": [round(x, 3) for x in step_ms], "setup_ms": last["setup_ms"], "classify_ms": cls_ms, "post_ms": last["post_ms"],
                       "item_post_ms": last["item_post_ms"], "gather_ms": last["gather_ms"],
                       "sharding": "none" if world == 1 else f"work items split over {world} ranks; one NCCL all-gather of 12-byte per-item records; every rank packs its shards to their final byte range; "
                                   + ("ranges gathered into rank 0's HBM (resident) / written by every rank into one page-locked host window (ommCpuBake)" if a.result_mode == "rank0"
                                      else "ranges broadcast in place to every rank"),
                       "secondary": secondary},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_info["h2d"], "d2h_bytes_per_step": e2e_info["d2h"],
                    "ms_per_step": 1e3 * sum(e2e_s) / len(e2e_s),
                    "call": "ommCpuBake + ommCpuGetBakeResultDesc, page-locked caller buffers -> host result" + ("" if world == 1 else " (read on rank 0)"),
                    "pageable_ms_per_step": 1e3 * sum(pg_s) / len(pg_s), "pageable_value": utris_total * len(pg_s) / sum(pg_s),
                    "last_step_breakdown": e2e_info["breakdown"]},
            "gpu_launches": launches,
            "roofline": roofline,
            "parity": parity,
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    if tex is not None:
        tex.destroy()
    baker.destroy()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def BakeInputFrom(inp, indices, texcoords):
    """A copy of a BakeInput with other host buffers."""
    import copy
    c = copy.copy(inp)
    c.indices, c.texcoords = indices, texcoords
    return c


def latest_profile_json(kind):
    """profiles/r<round><letter>_<kind>.json of the newest round present (written by scripts/ncu_summary.py from an ncu capture of this build)."""
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{kind}.json")))
    if not cands:
        return None
    with open(cands[-1]) as f:
        j = json.load(f)
    j["file"] = os.path.relpath(cands[-1], ROOT)
    return j


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
