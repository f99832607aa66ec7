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
          the same workload per step (rank 0 only).

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
def run_reference(a):
    """The SDK's CPU baker on a bounded slice of the same workload, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if os.path.exists(REF_LIB):
        os.environ["OMP_NUM_THREADS"] = str(cores)
        os.environ.setdefault("OMP_PROC_BIND", "spread")
        lib, kind, used = capi.OmmLib(REF_LIB), "reference", cores
        sample = a.cpu_sample_tris or max(1024, a.tris // 64)
    else:
        if not os.path.exists(PORT_LIB):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
        lib, kind, used = capi.OmmLib(PORT_LIB), "port", 1
        sample = a.cpu_sample_tris or max(256, a.tris // 1024)
    sample = min(sample, a.tris)
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
    sample_txt = f"first {sample} of {a.tris} triangles of the same grid/texture ({utris:.3e} micro-triangles) per step"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample_txt},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def cpu_baseline(a):
    cores = os.cpu_count() or 1
    if os.path.exists(REF_LIB):
        os.environ["OMP_NUM_THREADS"] = str(cores)
        os.environ.setdefault("OMP_PROC_BIND", "spread")
        lib, kind, used = capi.OmmLib(REF_LIB), "reference", cores
        sample = a.cpu_sample_tris or max(1024, a.tris // 32)
    elif os.path.exists(PORT_LIB):
        lib, kind, used = capi.OmmLib(PORT_LIB), "port", 1
        sample = a.cpu_sample_tris or max(256, a.tris // 512)
    else:
        return None
    sample = min(sample, a.tris)
    wl = W.config3(num_tris=sample, tex_size=a.tex, level=a.level)
    with Baker(lib) as b:
        inp, tex = W.make_input(b, wl, bake_flags=capi.BAKE_ENABLE_INTERNAL_THREADS)
        desc = inp.to_desc()
        t0 = time.perf_counter()
        rc, h = b.bake_raw(desc)
        dt = time.perf_counter() - t0
        assert rc == capi.SUCCESS
        lib.dll.ommCpuDestroyBakeResult(h)
        tex.destroy()
    utris = sample * 4 ** a.level
    return {"value": utris / dt, "unit": UNIT, "cores": used, "kind": kind, "seconds": dt,
            "sample": f"first {sample} of {a.tris} triangles of the same grid/texture, one ommCpuBake call ({utris:.3e} micro-triangles)"}


# ---------------------------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch

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

    wl = W.config3(num_tris=a.tris, tex_size=a.tex, level=a.level)
    utris_total = a.tris * 4 ** a.level
    idx_pinned, _k1 = pinned_like(wl.indices)
    uv_pinned, _k2 = pinned_like(wl.texcoords)
    wl.indices, wl.texcoords = idx_pinned, uv_pinned

    baker = Baker(lib)
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
    inp, tex = W.make_input(baker, wl)
    desc = inp.to_desc()
    stream = torch.cuda.current_stream()
    stream_ptr = C.c_void_p(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ----
    staged = C.c_void_p()
    assert lib.dll.ommB200StageInputs(baker.handle, C.byref(desc), C.byref(staged)) == capi.SUCCESS
    tm = capi.B200BakeTimings()
    step_ms, classify_ms, launches, last = [], [], 0, None
    sampler = ClockSampler(local)
    for it in range(a.warmup + a.steps):
        flush.fill_(it & 0xFF)  # evict L2 between steps (outside the timed events)
        barrier()
        if rank == 0 and it == max(a.warmup - 1, 0):
            sampler.start()
            sampler.wait_first()
        if it == a.warmup:
            sampler.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h = C.c_void_p()
        rc = lib.dll.ommB200BakeResident(baker.handle, staged, stream_ptr, C.byref(h))
        e1.record(stream)
        assert rc == capi.SUCCESS, f"ommB200BakeResident -> {rc}"
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        lib.dll.ommB200GetLastBakeTimings(baker.handle, C.byref(tm))
        if it >= a.warmup:
            step_ms.append(ms)
            classify_ms.append(tm.classifyMs)
            launches += tm.kernelLaunches
        last = (tm.workItems, tm.arrayDataBytes, tm.descCount, tm.microTriangles, tm.setupMs, tm.postMs, tm.itemPostMs, tm.gatherMs)
        lib.dll.ommCpuDestroyBakeResult(h)
    clocks = sampler.stop()
    lib.dll.ommB200DestroyStagedInputs(staged)
    total_ms = sum(step_ms)
    value = utris_total * len(step_ms) / (total_ms * 1e-3)

    # ---- end-to-end arm: the drop-in ommCpuBake with host buffers ----
    e2e_s, h2d, d2h = [], 0, 0
    for it in range(a.warmup + a.steps):
        barrier()
        t0 = time.perf_counter()
        h = C.c_void_p()
        rc = lib.dll.ommCpuBake(baker.handle, C.byref(desc), C.byref(h))
        pdesc = C.POINTER(capi.CpuBakeResultDesc)()
        # the host copy of the result is materialised where it is consumed: on rank 0 (every rank holds the complete result in HBM and
        # could download it; N simultaneous 290 MB downloads through one host only measure the host's memory system)
        rc2 = lib.dll.ommCpuGetBakeResultDesc(h, C.byref(pdesc)) if rank == 0 else capi.SUCCESS
        dt = time.perf_counter() - t0
        assert rc == capi.SUCCESS and rc2 == capi.SUCCESS
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        lib.dll.ommB200GetLastBakeTimings(baker.handle, C.byref(tm))
        h2d, d2h = int(tm.h2dBytes), int(tm.d2hBytes)
        host_break = {"stage_ms": tm.hostStageMs, "bake_ms": tm.hostBakeMs, "download_ms": tm.hostDownloadMs, "h2d_ms": tm.h2dMs, "d2h_ms": tm.d2hMs,
                      "device_total_ms": tm.totalDeviceMs}
        if it >= a.warmup:
            e2e_s.append(dt)
            launches += tm.kernelLaunches
        lib.dll.ommCpuDestroyBakeResult(h)
    e2e_value = utris_total * len(e2e_s) / sum(e2e_s)

    # ---- roofline of the dominant stage: classification (the Hier* kernels, >= 80 % of the device time of a step) ----
    peaks, peak_kind = measured_peaks()
    work_items, array_bytes, desc_count, my_utris, setup_ms, post_ms, item_post_ms, gather_ms = last
    tex_bytes = a.tex * a.tex * 4
    # algorithmic bytes of one classification pass on this rank: the texture once, one 32-byte item record per work item,
    # 2 bits written per micro-triangle (DESIGN.md "Kernels").  Duration: CUDA events recorded by the library on the bake's
    # stream around the stage (ommB200BakeTimings.classifyMs), averaged over the timed steps.
    classify_bytes = tex_bytes + 32 * (work_items // world) + my_utris // 4
    cls_ms = sum(classify_ms) / len(classify_ms)
    achieved = classify_bytes / (cls_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1d_stage_traffic.json")
    if world == 1 and a.tris == 1_000_000 and a.level == 6 and a.tex == 4096 and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes_read_per_bake"] + tj["dram_bytes_written_per_bake"]  # ncu capture of this very command, see the file
    roofline = {"bound": "hbm", "kernel": "classification stage = HierTestInitial + HierTestUnresolved + 2x HierTestList + HierLeaves per chunk (HierLeaves ~55 % of it)",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                "kernel_ms": cls_ms, "algorithmic_bytes": classify_bytes,
                "note": "instruction-issue bound, not HBM bound: 0.27 algorithmic bytes per micro-triangle against the bit-exact level-line arithmetic "
                        "(IEEE divisions and square roots) of every micro-triangle the level line touches; the hierarchical classifier removes the "
                        "arithmetic of provably uniform regions (97 % of the micro-triangles), profiles/r1c_HierLeaves_* and r1d_HierTestList_* hold issue utilisation and pipe mix"}
    # whole-path algorithmic bytes per SURVEY 8d: texture + geometry + outputs
    path_bytes = tex_bytes + wl.indices.nbytes + wl.texcoords.nbytes + array_bytes + 8 * desc_count + 4 * a.tris

    if rank == 0:
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            cpu = cpu_baseline(a)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / len(step_ms), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "l2": "256 MiB buffer written between steps (outside the timed events)",
                       "work_items": work_items, "array_data_bytes": array_bytes, "desc_count": desc_count,
                       "path_algorithmic_bytes": path_bytes, "step_ms": [round(x, 3) for x in step_ms], "setup_ms": setup_ms, "classify_ms": cls_ms, "post_ms": post_ms, "item_post_ms": item_post_ms, "gather_ms": gather_ms,
                       "sharding": "none" if world == 1 else f"work items split over {world} ranks, 1 NCCL all-gather of state blocks"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * sum(e2e_s) / len(e2e_s),
                    "call": "ommCpuBake + ommCpuGetBakeResultDesc, pinned host inputs -> host result" + ("" if world == 1 else " (downloaded on rank 0)"),
                    "last_step_breakdown": host_break},
            "gpu_launches": launches,
            "roofline": roofline,
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    tex.destroy()
    baker.destroy()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
