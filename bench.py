#!/usr/bin/env python
"""bench.py — Mrays/s of the render hot path on N B200s (one process per GPU), with the reference
CPU renderer timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one frame of the workload (BASELINE.json configs[1]: car_only, 1920x1080, 1 spp, reference
camera, 4 bounces, heuristic-6 tree), rendered through the C-ABI (include/rt_b200.h).  With N > 1 the
SAME frame is cut into interleaved 16x8 tiles, rank r renders the tiles of part r (scene replicated per
GPU) and the frame is assembled on rank 0 — fused (every rank's kernel stores finished pixels straight
into rank 0's frame over NVLink, mapped with CUDA IPC) or unfused (`--gather nccl`: packed tiles,
NCCL all-gather, unpack kernel).  Total work is fixed as N grows: "scaling": "strong".

  value  rays per frame x K / (sum of per-frame device times, CUDA events on the launching stream,
         max over ranks); L2 is flushed between timed frames (a 512 MiB buffer is overwritten)
  e2e    same metric by wall clock through the public API with HOST buffers: per step the camera /
         render parameters go host->device (kernel arguments), the frame is assembled and the 8-bit
         frame is copied device->host into pinned memory; frames are queued as a sequence, so frame
         k+1 renders while frame k is copied (rt_render_async / rt_download_async / rt_frame_wait)
  rays   one ray = one closest-hit or one shadow traversal; counted by the kernel and
         checked against the oracle's count in tests/test_gpu_parity.py

`--impl reference` times the reference's own CPU renderer (oracle/_ref, the unmodified reference
sources compiled by oracle/Makefile) on the host cores, same workload and metric.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # BASELINE.json configs[1]
    "car_only_1080p": dict(scene="car_only", width=1920, height=1080, spp=1),
    # configs[2] (two_cars geometry is missing from the reference tree -> car_boxed)
    "car_boxed_4k": dict(scene="car_boxed", width=3840, height=2160, spp=1),
    # configs[0] (dragon missing -> the reference's default scene)
    "car_boxed_1080p": dict(scene="car_boxed", width=1920, height=1080, spp=1),
}
METRIC = "Mrays/s"


def config_of(wl_name: str, wl: dict) -> dict:
    """The workload both arms are measured on — the SAME dict in both JSON lines (how each arm ran it goes under "run")."""
    return {"workload": wl_name, **wl, "bounces": 4, "bvh": "heuristic 6 (reference GPU default, gpu/include/options.cuh:50)",
            "camera": "reference default (cpu/src/main.c:105-106)"}


def scene_file(name: str) -> Path:
    return ROOT / "tests" / "golden" / "scenes" / f"{name}.rtsc"


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}  # B200_PROFILING.md


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.split(",") for l in Path(self.f.name).read_text().splitlines() if l.count(",") >= 8]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def reference_arm(args, wl_name: str, wl: dict) -> dict:
    """The reference's CPU renderer on this box's host cores: K timed frames after W warm-up frames."""
    import oracle as O
    ref = O.RefCpu(6)
    cores = os.cpu_count() or 1
    base = {"impl": "reference", "metric": METRIC, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl_name, wl)}
    if not ref.available:
        base["unavailable"] = "oracle/_ref reference binary missing or not runnable on this host"
        return base
    rays = rays_per_frame(wl)
    # bounded sample: at most ~60 s of CPU work
    probe = ref.run(rtsc=scene_file(wl["scene"]), width=wl["width"], height=wl["height"], spp=wl["spp"], threads=cores, frames=1, aov=False)
    per_frame = probe["frame_ms"][0]
    k = max(1, min(args.steps, int(60000 / max(per_frame, 1e-3))))
    w = max(0, min(args.warmup, int(20000 / max(per_frame, 1e-3))))
    r = ref.run(rtsc=scene_file(wl["scene"]), width=wl["width"], height=wl["height"], spp=wl["spp"], threads=cores, frames=k, warmup=w, aov=False)
    ms = sum(r["frame_ms"]) / len(r["frame_ms"])
    val = rays / ms / 1e3
    base.update({"value": val, "ms_per_step": ms, "steps": k, "warmup": w,
                 "cpu_baseline": {"value": val, "unit": METRIC, "cores": cores, "kind": "reference",
                                  "sample": f"{k} full frames of {wl_name} after {w} warm-up frames, {cores} pthreads, "
                                            f"reference flags -O3 -ffast-math -flto ({ref.isa})"},
                 "e2e": {"value": val, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0, "rays_per_frame": rays})
    return base


_RAYS = {}


def rays_per_frame(wl: dict) -> int:
    """Rays per frame of a workload: counted by the strict oracle on the CPU (bounded: the count is per
    pixel and resolution independent to first order, so large frames are counted at 1/4 scale per axis
    and scaled; exact counts are used on the GPU arm, from the kernel's own counters)."""
    key = (wl["scene"], wl["width"], wl["height"], wl["spp"])
    if key in _RAYS:
        return _RAYS[key]
    known = {("car_only", 1920, 1080, 1): 2978532, ("car_boxed", 1920, 1080, 1): 13247876}  # SURVEY §8(d), tests/test_gpu_parity.py
    if key in known:
        _RAYS[key] = known[key]
        return known[key]
    import oracle as O
    s = O.Oracle().scene(O.load_rtsc(scene_file(wl["scene"])))
    s.build_bvh(6 | 0x100)
    w, h = wl["width"] // 4, wl["height"] // 4
    r = s.render(w, h, spp=wl["spp"])
    _RAYS[key] = int((r["rays_closest"] + r["rays_shadow"]) * (wl["width"] * wl["height"]) / (w * h))
    return _RAYS[key]


# ------------------------------------------------------------------------------------------
class CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can view library-owned device memory."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def gpu_arm(args, wl_name: str, wl: dict) -> dict:
    import numpy as np
    import torch
    import torch.distributed as dist
    import parallel_ray_tracer_b200 as rt

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    if rt.device_count() < 1:
        raise SystemExit("no CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    W, H, SPP = wl["width"], wl["height"], wl["spp"]
    sc = rt.Scene.load_rtsc(scene_file(wl["scene"])).build_bvh(6)
    t0 = time.perf_counter()
    ctx = rt.Context(sc, [local])
    create_ms = (time.perf_counter() - t0) * 1e3
    base = dict(width=W, height=H, spp=SPP, part_index=rank, part_count=world)
    if args.ctas_per_sm: base["ctas_per_sm"] = args.ctas_per_sm
    if args.block: base["block_threads"] = args.block
    if args.refill: base["refill_threshold"] = args.refill
    params = rt.default_params(**base)

    # ---- frame assembly set-up ----
    gather = args.gather if world > 1 else "none"
    if gather == "ipc":
        try:
            for slot in range(rt.RT_FRAME_SLOTS):   # both frame slots of rank 0 become peer-store targets
                h = torch.zeros(64, dtype=torch.uint8, device=dev)
                if rank == 0:
                    h.copy_(torch.frombuffer(bytearray(ctx.frame_ipc_export(W, H, slot)), dtype=torch.uint8))
                dist.broadcast(h, 0)
                if rank != 0:
                    ctx.frame_ipc_import(bytes(h.cpu().numpy().tobytes()), W, H, slot)
            ok = torch.ones(1, device=dev)
        except rt.RtError as e:
            print(f"[rank {rank}] CUDA IPC mapping failed ({e}); falling back to --gather nccl", file=sys.stderr)
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            gather = "nccl"
    n_local = rt.part_tile_count(W, H, rank, world)
    stride = max(rt.part_tile_count(W, H, p, world) for p in range(world)) * 128 * 4
    if gather == "nccl":
        send = torch.zeros(stride, dtype=torch.uint8, device=dev)
        recv = torch.zeros(stride * world, dtype=torch.uint8, device=dev)

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    pinned = [torch.empty(W * H * 4, dtype=torch.uint8).pin_memory() for _ in range(rt.RT_FRAME_SLOTS)]

    def assemble() -> float:
        """Unfused path: packed tiles -> NCCL all-gather -> unpack on rank 0.  Returns device ms."""
        if gather != "nccl":
            return 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ptr, nbytes = ctx.packed_tiles()
        e0.record()
        if nbytes:
            send[:nbytes].copy_(torch.as_tensor(CudaArray(ptr, nbytes), device=dev))
        dist.all_gather_into_tensor(recv, send)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if rank == 0:
            t = time.perf_counter()
            ctx.unpack_tiles(recv.data_ptr(), stride, world)
            ms += (time.perf_counter() - t) * 1e3
        return ms

    def step(flush_l2: bool):
        if flush_l2:
            flush.fill_(1)
            torch.cuda.synchronize()
        tm = ctx.render_frame(params)
        g = assemble()
        return tm, tm.kernel_ms[0] + g

    # ---- warm-up, then EXACTLY K timed steps ----
    sampler = ClockSampler(local) if rank == 0 else None   # started before the warm-up: nvidia-smi needs ~0.1 s to deliver its first sample
    for _ in range(max(args.warmup, 3)):
        step(True)
    barrier(); torch.cuda.synchronize()
    dev_ms, kern_ms, launches = [], [], 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        tm, ms = step(True)
        dev_ms.append(ms); kern_ms.append(tm.kernel_ms[0])
        launches += tm.launches + (2 if gather == "nccl" else 0)
    torch.cuda.synchronize(); barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop() if sampler else None
    rays_local = tm.rays_closest + tm.rays_shadow

    # L2-warm variant (no flush) for reference
    warm = []
    for _ in range(args.steps):
        tm2, ms = step(False)
        warm.append(ms)

    # ---- e2e: public API, host buffers, every frame's D2H inside the timed region ----
    # Frame sequence as the reference runs it (ITERATIONS frames, gpu/src/main.cu:111-114), through the sequence API:
    # frame k renders into frame slot k % 2 while frame k-1 is copied device->host on a second stream
    # (rt_render_async / rt_download_async / rt_frame_wait).  Per step: render parameters H2D (kernel arguments),
    # the assembled 8-bit frame D2H into pinned memory.  With the unfused NCCL gather the steps stay synchronous.
    slot_params = []
    for slot in range(rt.RT_FRAME_SLOTS):
        slot_params.append(rt.default_params(**base, frame_slot=slot))
    barrier_s = [0.0]   # host time inside dist.barrier() of the e2e loop (includes waiting for the slowest rank's kernel)

    def e2e_loop(steps):
        if gather == "nccl":
            for _ in range(steps):
                ctx.render_frame(params)
                assemble()
                barrier()
                if rank == 0:
                    ctx.download_into(pinned[0].data_ptr())
        elif world == 1:
            for k in range(steps):
                s = k % rt.RT_FRAME_SLOTS
                if k >= rt.RT_FRAME_SLOTS:
                    ctx.frame_wait(s)           # frame k-2 is on the host (pinned[s] is consumed here)
                ctx.render_frame_async(slot_params[s])
                ctx.download_async(s, pinned[s].data_ptr())
            for s in range(min(steps, rt.RT_FRAME_SLOTS)):
                ctx.frame_wait(s)
        else:
            for k in range(steps):
                s = k % rt.RT_FRAME_SLOTS
                ctx.render_frame_async(slot_params[s])
                ctx.frame_wait(s)               # this rank's tiles of frame k are stored in rank 0's slot s
                if rank == 0 and k >= 1:
                    ctx.frame_wait(1 - s)       # D2H of frame k-1 (ran during this render) is complete: slot 1-s is free again
                tb = time.perf_counter()
                barrier()                       # frame k on rank 0 is complete once every rank has stored its tiles
                barrier_s[0] += time.perf_counter() - tb
                if rank == 0:
                    ctx.download_async(s, pinned[s].data_ptr())
            if rank == 0:
                for s in range(min(steps, rt.RT_FRAME_SLOTS)):
                    ctx.frame_wait(s)

    e2e_loop(max(args.warmup, 4))   # untimed: second frame slot allocation, lazy IPC peer mapping, first NCCL barriers
    barrier(); torch.cuda.synchronize()
    barrier_s[0] = 0.0
    t_e0 = time.perf_counter()
    e2e_loop(args.steps)
    torch.cuda.synchronize(); barrier()
    e2e_ms = (time.perf_counter() - t_e0) * 1e3 / args.steps
    barrier_ms = barrier_s[0] * 1e3 / args.steps
    # unpipelined reference point: render, wait, copy, wait — one frame at a time
    barrier(); torch.cuda.synchronize()
    t_s0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.render_frame(params)
        assemble()
        barrier()
        if rank == 0:
            ctx.download_into(pinned[0].data_ptr())
    torch.cuda.synchronize(); barrier()
    e2e_sync_ms = (time.perf_counter() - t_s0) * 1e3 / args.steps

    # ---- frame sequences across GPUs: alternate-frame rendering (every rank renders WHOLE frames of the sequence, its own
    # copy stream, no frame assembly) — the weak-scaling counterpart of the tile split above, reported beside it ----
    afr = None
    if world > 1:
        full = [rt.default_params(width=W, height=H, spp=SPP, frame_slot=s) for s in range(rt.RT_FRAME_SLOTS)]
        ctx_full = rt.Context(sc, [local])   # a context without the imported frame: renders into its own slots

        def afr_loop(steps):
            for k in range(steps):
                s = k % rt.RT_FRAME_SLOTS
                if k >= rt.RT_FRAME_SLOTS:
                    ctx_full.frame_wait(s)
                ctx_full.render_frame_async(full[s])
                ctx_full.download_async(s, pinned[s].data_ptr())
            for s in range(min(steps, rt.RT_FRAME_SLOTS)):
                ctx_full.frame_wait(s)
        afr_loop(4)
        barrier(); torch.cuda.synchronize()
        t_a0 = time.perf_counter()
        afr_loop(args.steps)
        torch.cuda.synchronize(); barrier()
        afr_ms = (time.perf_counter() - t_a0) * 1e3
        afr = (afr_ms,)
        ctx_full.close()

    # ---- multi-GPU correctness: the assembled N-part frame must be byte-identical to the frame one
    # GPU renders alone.  Every frame slot is poisoned first, so a stale frame cannot pass. ----
    def frame_check(cx, scene, w, h, spp, slots, asm):
        if world == 1:
            return None
        full = None
        if rank == 0:
            c1 = rt.Context(scene, [local])
            c1.render_frame(rt.default_params(width=w, height=h, spp=spp))
            full = c1.load_from_gpu()["bgra"].copy()
            c1.close()
        ok = True
        for slot in slots:
            kw = dict(width=w, height=h, spp=spp, part_index=rank, part_count=world, frame_slot=slot)
            cx.render_frame(rt.default_params(**kw))          # makes `slot` the context's current frame
            if rank == 0:
                ptr, nbytes = cx.frame_device_ptr()
                torch.as_tensor(CudaArray(ptr, nbytes), device=dev).fill_(0x5a)
            torch.cuda.synchronize(); barrier()
            cx.render_frame(rt.default_params(**kw))
            if asm:
                assemble()
            torch.cuda.synchronize(); barrier()
            if rank == 0:
                ok = ok and bool(np.array_equal(cx.load_from_gpu()["bgra"], full))
        tk = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(tk, op=dist.ReduceOp.MIN)
        return bool(tk.item() > 0)

    frame_ok = frame_check(ctx, sc, W, H, SPP, range(rt.RT_FRAME_SLOTS) if gather == "ipc" else [0], gather == "nccl")

    # ---- reductions over ranks ----
    def allmax(x):
        if world == 1: return x
        t = torch.tensor([x], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return t.item()

    def allsum(x):
        if world == 1: return x
        t = torch.tensor([x], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.SUM); return t.item()

    def sum_of_step_max(v):
        """A strong-scaled frame is complete when its SLOWEST rank is: max over ranks per step, then the sum over steps."""
        if world == 1: return float(sum(v))
        tv = torch.tensor(v, dtype=torch.float64, device=dev); dist.all_reduce(tv, op=dist.ReduceOp.MAX); return float(tv.sum().item())

    total_dev_ms = sum_of_step_max(dev_ms)
    total_kern_ms = sum_of_step_max(kern_ms)
    total_warm_ms = sum_of_step_max(warm)
    wall_ms = allmax(wall_ms)
    e2e_ms = allmax(e2e_ms)
    e2e_sync_ms = allmax(e2e_sync_ms)
    afr_ms = allmax(afr[0]) if afr else None
    rays = int(allsum(rays_local))
    launches = int(allsum(launches))

    # ---- work counters (separate, untimed pass with the counting build) for the roofline ----
    # (strict build: it walks the reference's visit order, so the counts are the algorithmic work of SURVEY §8d)
    pw = rt.default_params(**base, aov_mask=rt.RT_AOV_WORK, mode=rt.RT_MODE_STRICT)
    tmw = ctx.render_frame(pw)
    inner = int(allsum(tmw.inner_visits)); tris = int(allsum(tmw.tri_tests))

    # ---- measured gather rooflines of this device (node/triangle bytes over L1/L2/HBM bandwidth) ----
    gather_peaks = None
    if rank == 0:
        gather_peaks = {"l1_resident_64KB": rt.gather_bandwidth(64 << 10, local), "l2_resident_4MB": rt.gather_bandwidth(4 << 20, local),
                        "hbm_8GB": rt.gather_bandwidth(8 << 30, local)}

    # ---- the multi-GPU configuration of BASELINE.json (configs[2]) measured the same way, fewer frames ----
    also = None
    if args.also and args.also != wl_name:
        wl2 = WORKLOADS[args.also]
        W2, H2 = wl2["width"], wl2["height"]
        sc2 = rt.Scene.load_rtsc(scene_file(wl2["scene"])).build_bvh(6)
        ctx2 = rt.Context(sc2, [local])
        p2 = rt.default_params(width=W2, height=H2, spp=wl2["spp"], part_index=rank, part_count=world)
        ok2 = True
        if world > 1:
            try:
                h2 = torch.zeros(64, dtype=torch.uint8, device=dev)
                if rank == 0:
                    h2.copy_(torch.frombuffer(bytearray(ctx2.frame_ipc_export(W2, H2)), dtype=torch.uint8))
                dist.broadcast(h2, 0)
                if rank != 0:
                    ctx2.frame_ipc_import(bytes(h2.cpu().numpy().tobytes()), W2, H2)
            except rt.RtError:
                ok2 = False
            okt = torch.tensor([1.0 if ok2 else 0.0], device=dev); dist.all_reduce(okt, op=dist.ReduceOp.MIN); ok2 = okt.item() > 0
        if ok2:
            k2 = max(5, min(args.steps, 15))
            for _ in range(3):
                flush.fill_(1); torch.cuda.synchronize(); ctx2.render_frame(p2)
            barrier(); torch.cuda.synchronize()
            ms2 = []
            for _ in range(k2):
                flush.fill_(1); torch.cuda.synchronize()
                tm2 = ctx2.render_frame(p2)
                ms2.append(tm2.kernel_ms[0])
            torch.cuda.synchronize(); barrier()
            tot2 = sum_of_step_max(ms2); rays2 = int(allsum(tm2.rays_closest + tm2.rays_shadow))
            tmw2 = ctx2.render_frame(rt.default_params(width=W2, height=H2, spp=wl2["spp"], part_index=rank, part_count=world,
                                                       aov_mask=rt.RT_AOV_WORK, mode=rt.RT_MODE_STRICT))
            inner2 = int(allsum(tmw2.inner_visits)); tris2 = int(allsum(tmw2.tri_tests))
            ok_also = frame_check(ctx2, sc2, W2, H2, wl2["spp"], [0], False)
            if frame_ok is not None:
                frame_ok = bool(frame_ok and ok_also)
            also = {"workload": args.also, **wl2, "frame_equals_1gpu": ok_also, "steps": k2, "ms_per_step": tot2 / k2, "value": rays2 / (tot2 / k2) / 1e3, "unit": METRIC,
                    "rays_per_frame": rays2, "note": "same timing rules as value (CUDA events, max over ranks, L2 flushed), fused peer stores"}
            if rank == 0:
                ach2 = (64 * inner2 + 40 * tris2) / world / (tot2 / k2 * 1e-3) / 1e9
                also["roofline_gather"] = {"unit": "GB/s", "achieved_algorithmic": ach2, "frac_of_l1": ach2 / gather_peaks["l1_resident_64KB"],
                                           "frac_of_l2": ach2 / gather_peaks["l2_resident_4MB"], "frac_of_hbm_stream_peak": ach2 / peaks()["hbm_gbs"]}
        ctx2.close()

    out = None
    if rank == 0:
        pk = peaks()
        ms_per_step = total_dev_ms / args.steps
        value = rays / ms_per_step / 1e3
        # algorithmic bytes / flops per launch: 64 B per inner visit (two 32 B child boxes),
        # 40 B per triangle test (36 B vertices + 4 B index); 48 flops per inner visit, 54 per triangle test.
        # The dominant kernel is the render kernel; with N ranks one launch handles 1/N of the frame.
        bytes_alg = (64 * inner + 40 * tris) / world
        flops_alg = (48 * inner + 54 * tris) / world
        k_ms = total_kern_ms / args.steps
        ach = bytes_alg / (k_ms * 1e-3) / 1e9
        fp_peak = 148 * 128 * pk["sm_max_mhz"] * 1e6 / 1e12  # T lane-ops/s at max clock
        fp_ach = flops_alg / (k_ms * 1e-3) / 1e12
        traffic = None
        prof = ROOT / "profiles" / "traffic.json"
        if prof.exists():
            traffic = json.loads(prof.read_text()).get(f"{wl_name}@{world}")
        out = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": config_of(wl_name, wl),
               "run": {"mode": "RT_MODE_FAST", "bvh_built": "on the host as the reference does (csrc/bvh_build.cpp)",
                       "partition": f"{world} x interleaved 16x8 tiles" if world > 1 else "single GPU",
                       "gather": {"ipc": "fused peer stores over NVLink (CUDA IPC)", "nccl": "packed tiles + NCCL all-gather + unpack",
                                  "none": "none"}[gather],
                       "l2": "flushed between timed frames (512 MiB overwrite)",
                       "scene_note": "car_only as shipped by the reference; substitutions for missing scenes in DESIGN.md"},
               "frame_equals_1gpu": frame_ok,
               "clocks": clocks,
               "e2e": {"value": rays / e2e_ms / 1e3, "unit": METRIC, "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": C.sizeof(rt.rt_render_params) * world, "d2h_bytes_per_step": W * H * 4,
                       "pipeline": ("synchronous (NCCL gather)" if gather == "nccl" else
                                    "frame k+1 renders while frame k is copied device->host (2 frame slots, rt_render_async / rt_download_async)"),
                       "barrier_ms_per_step": barrier_ms,
                       "value_unpipelined": rays / e2e_sync_ms / 1e3, "ms_per_step_unpipelined": e2e_sync_ms},
               "gpu_launches": launches,
               "rays_per_frame": rays, "primary_mrays_s": W * H * SPP / ms_per_step / 1e3,
               "value_l2_warm": rays / (total_warm_ms / args.steps) / 1e3,
               "wall_ms_per_step_incl_flush": wall_ms / args.steps, "kernel_ms_per_step": k_ms,
               "scene_upload_ms": create_ms,
               "roofline": {"bound": "l1_gather", "achieved": ach, "peak": gather_peaks["l1_resident_64KB"], "unit": "GB/s",
                            "frac": ach / gather_peaks["l1_resident_64KB"], "traffic": traffic,
                            "peak_source": "measured live on this device by rt_debug_gather_bandwidth (random 64-byte records, working set in L1)",
                            "frac_l2_gather": ach / gather_peaks["l2_resident_4MB"],
                            "frac_hbm_stream": ach / pk["hbm_gbs"], "hbm_stream_peak": pk["hbm_gbs"], "hbm_peak_source": pk["source"],
                            "note": "algorithmic bytes = 64 B x inner visits + 40 B x triangle tests per launch (SURVEY 8d, reference visit "
                                    "order); the 4-5 MB scene is L1/L2 resident (DRAM traffic = `traffic`, ~0.1 % of the algorithmic bytes), so "
                                    "the bound that applies is the L1 record-gather rate, not HBM; HBM only binds the 50 M-triangle scene "
                                    "(profiles/r02_config5_*)"},
               "roofline_gather": {"unit": "GB/s", "achieved_algorithmic": ach, **{k: round(v, 1) for k, v in gather_peaks.items()},
                                   "frac_of_l1": ach / gather_peaks["l1_resident_64KB"], "frac_of_l2": ach / gather_peaks["l2_resident_4MB"],
                                   "note": "peaks measured live: random 64-byte-record gathers, one record per lane (the shape of a node fetch), working "
                                           "set resident in L1 / L2 / HBM (rt_debug_gather_bandwidth); the shipped scenes (4-5 MB) are L1/L2 resident, "
                                           "L1 hit rate 95-98 % (profiles/), so the L1 figure is the bound that applies"},
               "roofline_fp32": {"achieved_tlaneops": fp_ach, "peak_tlaneops": fp_peak, "frac": fp_ach / fp_peak,
                                 "note": "48 flops x inner visits + 54 x triangle tests vs 148 SM x 128 lanes x max SM clock"},
               "sequence_afr": (None if afr_ms is None else {
                   "value": rays * world * args.steps / afr_ms / 1e3, "unit": METRIC, "frames": world * args.steps, "ms_total": afr_ms, "scaling": "weak",
                   "note": "alternate-frame rendering of a frame sequence: every rank renders whole frames (K each) and copies them to its own "
                           "pinned host buffer; no frame assembly.  Reported beside the tile-split numbers, not instead of them"}),
               "also": also,
               "work": {"inner_visits": inner, "tri_tests": tris,
                        "note": "reference visit order (strict build counters == oracle counters, tests/test_gpu_parity.py)"}}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return out


def ref_gpu_baseline(wl: dict, rays: int) -> dict:
    """The reference's own GPU program ("fuse", the only stage in its tree; staged + patched for gcc-hosted nvcc by
    scripts/stage_ref_gpu.py into git-ignored baseline/_ref/gpu, -arch=sm_100) on this B200, at the best block shape of its
    .bat sweeps (8x8, profiles/r01_reference_gpu_fuse_b200.jsonl), timed by ITS OWN protocol: CUDA events around the launch,
    50 warm-up + 100 timed frames, median (gpu/src/main.cu:111-127).  It culls with round-to-nearest FP16 boxes
    (gpu/src/gpu.cu:176-185): not parity-equivalent to the CPU renderer, a speed baseline only."""
    import re
    g = ROOT / "baseline" / "_ref" / "gpu"
    exe = g / f"raytracer_{wl['scene']}_{wl['width']}x{wl['height']}"
    if wl["spp"] != 1 or not exe.exists():
        return {"unavailable": f"{exe.name} not staged (scripts/stage_ref_gpu.py builds it in the build container)"}
    try:
        r = subprocess.run([str(exe), "8", "8"], cwd=g, capture_output=True, text=True, timeout=600)
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{exe.name}: {e}"}
    m = re.search(r"Frame time \(median\): ([0-9.]+) ms", r.stdout)
    if r.returncode != 0 or not m:
        return {"unavailable": f"{exe.name} failed: {(r.stderr or r.stdout)[-200:]}"}
    ms = float(m.group(1))
    return {"ms": ms, "mrays_s": rays / ms / 1e3, "block": "8x8", "protocol": "50 warm-up + 100 timed frames, CUDA events, median",
            "note": "reference gpu/ kernel recompiled for sm_100; FP16 non-conservative culling; rays = this frame's oracle count"}


def cpu_baseline(wl_name: str, wl: dict) -> dict:
    """Bounded sample of the reference CPU renderer on this box's host cores (rank 0, N = 1 only)."""
    import oracle as O
    cores = os.cpu_count() or 1
    ref = O.RefCpu(6)
    rays = rays_per_frame(wl)
    if ref.available:
        r = ref.run(rtsc=scene_file(wl["scene"]), width=wl["width"], height=wl["height"], spp=wl["spp"], threads=cores, frames=3, warmup=1, aov=False)
        ms = statistics.median(r["frame_ms"])
        out = {"value": rays / ms / 1e3, "unit": METRIC, "cores": cores, "kind": "reference", "frame_ms": ms,
               "sample": f"median of 3 full frames of {wl_name} (+1 warm-up), {cores} pthreads, oracle/_ref (unmodified reference "
                         f"sources, -O3 -ffast-math -flto, {ref.isa}), heuristic-6 tree"}
        # the reference CPU program's own default tree (BVH_HEURISTIC 3, cpu/include/options.h:34): same image, slower walk
        ref3 = O.RefCpu(3)
        if ref3.available:
            r3 = ref3.run(rtsc=scene_file(wl["scene"]), width=wl["width"], height=wl["height"], spp=wl["spp"], threads=cores, frames=2, warmup=1, aov=False)
            ms3 = statistics.median(r3["frame_ms"])
            out["heuristic3_tree"] = {"value": rays / ms3 / 1e3, "frame_ms": ms3, "note": "rays counted on the heuristic-6 tree; the image is the same"}
        return out
    s = O.Oracle().scene(O.load_rtsc(scene_file(wl["scene"])))
    s.build_bvh(6 | 0x100)
    t = time.perf_counter(); s.render(wl["width"], wl["height"], spp=wl["spp"], threads=cores); ms = (time.perf_counter() - t) * 1e3
    return {"value": rays / ms / 1e3, "unit": METRIC, "cores": cores, "kind": "port", "frame_ms": ms,
            "sample": f"1 full frame of {wl_name}, {cores} pthreads, oracle/rt_oracle.c (-O2, strict IEEE)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="car_only_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--gather", default="ipc", choices=["ipc", "nccl"])
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--refill", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference GPU program (baseline/_ref/gpu)")
    ap.add_argument("--also", default="car_boxed_4k", help="second workload reported under \"also\" ('' to skip)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args, args.workload, wl)), flush=True)
        return
    out = gpu_arm(args, args.workload, wl)
    if rank == 0:
        if args.gpus == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.workload, wl)
        if args.gpus == 1 and not args.no_ref_gpu:
            out["ref_gpu_baseline"] = ref_gpu_baseline(wl, out["rays_per_frame"])
            if "ms" in out["ref_gpu_baseline"]:
                out["ref_gpu_baseline"]["speedup_kernel"] = out["ref_gpu_baseline"]["ms"] / out["kernel_ms_per_step"]
            if out.get("also"):
                rg = ref_gpu_baseline(WORKLOADS[out["also"]["workload"]], out["also"]["rays_per_frame"])
                if "ms" in rg:
                    rg["speedup_kernel"] = rg["ms"] / out["also"]["ms_per_step"]
                out["also"]["ref_gpu_baseline"] = rg
        print(json.dumps(out), flush=True)
        if out.get("frame_equals_1gpu") is False:
            print("bench.py: the assembled multi-GPU frame differs from the single-GPU frame", file=sys.stderr)
            sys.exit(3)


if __name__ == "__main__":
    main()
