#!/usr/bin/env python
"""Run the five BASELINE.json configurations once on one B200 (substitutions per SURVEY.md §8d) and print one JSON
line each: kernel ms (median, CUDA events), Mrays/s, parity of the fast build against the strict build at full
size, scene bytes in HBM.  Results are copied to profiles/ by hand.

  1  cpu/raytracer on dragon            -> car_boxed 1920x1080 (the reference's default scene), + reference CPU time
  2  car_only 1920x1080 1 spp           -> as is
  3  two_cars 3840x2160                 -> car_boxed 3840x2160
  4  sportscar 7680x4320, spp 1..64     -> car_boxed 7680x4320, spp in {1,2,4,8,16,32,64}
  5  dragon x N ~ 50 M triangles, 4K    -> car_only instanced 39 x 40 = 1560 copies (50.1 M triangles), 3840x2160
"""
import json
import os
import statistics
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import parallel_ray_tracer_b200 as rt  # noqa: E402
import oracle as O  # noqa: E402

GOLD = ROOT / "tests" / "golden" / "scenes"


def measure(ctx, frames, **kw):
    p = rt.default_params(**kw)
    t_end = time.perf_counter() + 0.2
    n = 0
    while time.perf_counter() < t_end or n < 2:
        tm = ctx.render_frame(p); n += 1
    ms = [ctx.render_frame(p).kernel_ms[0] for _ in range(frames)]
    tm = ctx.render_frame(p)
    return statistics.median(ms), tm.rays_closest + tm.rays_shadow


def parity(ctx, w, h, **kw):
    aov = rt.RT_AOV_TRI_ID | rt.RT_AOV_DEPTH
    ctx.render_frame(rt.default_params(width=w, height=h, mode=rt.RT_MODE_FAST, aov_mask=aov, **kw))
    a = ctx.load_from_gpu(tri_id=True, depth=True)
    ctx.render_frame(rt.default_params(width=w, height=h, mode=rt.RT_MODE_STRICT, aov_mask=aov, **kw))
    b = ctx.load_from_gpu(tri_id=True, depth=True)
    return {k: round(v, 6) if isinstance(v, float) else v for k, v in O.compare_aovs(a, b).items()}


def main():
    which = set(sys.argv[1:]) or {"1", "2", "3", "4", "5"}
    out = lambda d: print(json.dumps(d), flush=True)
    scenes = {}

    def ctx_of(name):
        if name not in scenes:
            sc = rt.Scene.load_rtsc(GOLD / f"{name}.rtsc").build_bvh(6)
            scenes[name] = (sc, rt.Context(sc, [0]))
        return scenes[name][1]

    if "1" in which:
        ctx = ctx_of("car_boxed")
        ms, rays = measure(ctx, 30, width=1920, height=1080)
        rec = {"config": 1, "workload": "car_boxed 1920x1080 (substitute for dragon)", "kernel_ms": ms, "rays": rays, "mrays_s": rays / ms / 1e3,
               "fast_vs_strict": parity(ctx, 1920, 1080)}
        ref = O.RefCpu(6)
        if ref.available:
            r = ref.run(rtsc=GOLD / "car_boxed.rtsc", width=1920, height=1080, frames=3, warmup=1, aov=False)
            rec["reference_cpu"] = {"frame_ms": statistics.median(r["frame_ms"]), "threads": os.cpu_count(), "mrays_s": rays / statistics.median(r["frame_ms"]) / 1e3}
        out(rec)
    if "2" in which:
        ctx = ctx_of("car_only")
        ms, rays = measure(ctx, 50, width=1920, height=1080)
        out({"config": 2, "workload": "car_only 1920x1080", "kernel_ms": ms, "rays": rays, "mrays_s": rays / ms / 1e3, "fast_vs_strict": parity(ctx, 1920, 1080)})
    if "3" in which:
        ctx = ctx_of("car_boxed")
        ms, rays = measure(ctx, 20, width=3840, height=2160)
        out({"config": 3, "workload": "car_boxed 3840x2160 (substitute for two_cars)", "kernel_ms": ms, "rays": rays, "mrays_s": rays / ms / 1e3,
             "fast_vs_strict": parity(ctx, 3840, 2160)})
    if "4" in which:
        ctx = ctx_of("car_boxed")
        for spp in (1, 2, 4, 8, 16, 32, 64):
            ms, rays = measure(ctx, 3 if spp >= 16 else 6, width=7680, height=4320, spp=spp, seed=1)
            out({"config": 4, "workload": f"car_boxed 7680x4320 spp {spp} (substitute for sportscar)", "kernel_ms": ms, "rays": rays, "mrays_s": rays / ms / 1e3})
    if "5" in which:
        t0 = time.perf_counter()
        base = rt.Scene.load_rtsc(GOLD / "car_only.rtsc")
        big = base.instance_grid(39, 40, 1, (11.5, 6.5, 3.0))
        t1 = time.perf_counter()
        big.build_bvh(6)
        t2 = time.perf_counter()
        # the same tree built on the GPU (csrc/bvh_build_gpu.cu): must equal the host build entry for entry
        import hashlib
        def tree_hash(sc):
            a = sc.arrays()
            return hashlib.sha256(a["bvh_nodes"].tobytes() + a["tri_idx"].tobytes()).hexdigest()
        h_host = tree_hash(big)
        tg0 = time.perf_counter()
        gst = big.build_bvh_gpu(6)
        tg1 = time.perf_counter()
        gpu_build = {"wall_s": tg1 - tg0, "equal_to_host_build": tree_hash(big) == h_host,
                     **{k: getattr(gst, k) for k, _ in gst._fields_}}
        t2b = time.perf_counter()
        ctx = rt.Context(big, [0])
        t3 = time.perf_counter()
        v = big.view()
        ms, rays = measure(ctx, 5, width=3840, height=2160)
        tmw = ctx.render_frame(rt.default_params(width=3840, height=2160, mode=rt.RT_MODE_STRICT, aov_mask=rt.RT_AOV_WORK))
        out({"config": 5, "workload": "car_only x 1560 instances = 50.1 M triangles, 3840x2160 (substitute for dragon x N)", "triangles": v.n_tris,
             "bvh_nodes": v.bvh_len, "instance_s": t1 - t0, "bvh_build_s": t2 - t1, "bvh_build_gpu": gpu_build, "flatten_upload_s": t3 - t2b, "threads": os.cpu_count(),
             "hbm_scene_bytes": 64 * (v.bvh_len // 2) + 64 * v.n_tris + 16 * v.n_tris,
             "kernel_ms": ms, "rays": rays, "mrays_s": rays / ms / 1e3,
             "algorithmic_bytes": 64 * tmw.inner_visits + 40 * tmw.tri_tests, "algorithmic_gbs": (64 * tmw.inner_visits + 40 * tmw.tri_tests) / ms / 1e6,
             "fast_vs_strict": parity(ctx, 3840, 2160)})
        # triangles -> render-ready context entirely on the device (rt_create_gpu: GPU build + GPU flatten, no host round trip)
        ctx.render_frame(rt.default_params(width=3840, height=2160))
        want = ctx.load_from_gpu()["bgra"].copy()
        ctx.close(); big.close()
        big2 = base.instance_grid(39, 40, 1, (11.5, 6.5, 3.0))
        tp0 = time.perf_counter()
        ctx2 = rt.Context.build_on_gpu(big2, [0])
        tp1 = time.perf_counter()
        ctx2.render_frame(rt.default_params(width=3840, height=2160))
        same = bool(np.array_equal(ctx2.load_from_gpu()["bgra"], want))
        out({"config": 5, "device_pipeline": {"triangles_to_context_s": tp1 - tp0, "build_ms": ctx2.build_stats.total_ms, "frame_equals_host_path": same}})
        ctx2.close(); big2.close()


if __name__ == "__main__":
    main()
