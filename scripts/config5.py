#!/usr/bin/env python
"""BASELINE.json configs[4] on one B200: car_only instanced 39 x 40 = 1 560 times = 50.1 M triangles (scale kept), 3840x2160.
Triangles -> render-ready context entirely on the device (rt_create_gpu: staged pinned upload, GPU BVH build, device-side
flatten incl. the compressed 8-wide tree), then every fast traversal variant against the strict build, kernel ms (median),
the byte counts of the scene arrays, and one strict RT_AOV_WORK pass for the algorithmic bytes.  JSON lines on stdout.
usage: python scripts/config5.py [nx ny] [frames]"""
import json, os, statistics, sys, time
os.environ.setdefault("RT_TIMING", "1")
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import parallel_ray_tracer_b200 as rt
import oracle as O


def main():
    nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (39, 40)
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    W, H = 3840, 2160
    base = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / "car_only.rtsc")
    warm = rt.Context.build_on_gpu(rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / "car_only.rtsc"), [0])  # CUDA context, module load
    warm.close()
    print(json.dumps({"h2d_gbs_1GB": {"pageable_cudaMemcpy": round(rt.copy_bandwidth(1 << 30, 0), 2), "staged_pinned_ring": round(rt.copy_bandwidth(1 << 30, 1), 2),
                                       "pinned_cudaMemcpy": round(rt.copy_bandwidth(1 << 30, 2), 2)}}), flush=True)
    t0 = time.perf_counter()
    big = base.instance_grid(nx, ny, 1, (11.5, 6.5, 3.0))
    t1 = time.perf_counter()
    ctx = rt.Context.build_on_gpu(big, [0])
    t2 = time.perf_counter()
    st = ctx.build_stats
    sizes = {name: int(ctx.device_array(which, np.uint8).nbytes) if which != 4 else 0 for which, name in ((0, "nodes"), (1, "nodes4"), (2, "tris"), (3, "shade"), (7, "nodes8"))}
    print(json.dumps({"config": 5, "triangles": nx * ny * 32136, "instance_s": round(t1 - t0, 3), "triangles_to_context_s": round(t2 - t1, 3),
                      "gpu_build": {k: getattr(st, k) for k, _ in st._fields_}, "scene_bytes": sizes,
                      "fast_only_bytes_wide8": sizes["nodes8"] + sizes["tris"] + sizes["shade"]}), flush=True)
    ctx.render_frame(rt.default_params(width=W, height=H, mode=rt.RT_MODE_STRICT, aov_mask=2 | 4))
    strict = {k: v.copy() for k, v in ctx.load_from_gpu(tri_id=True, depth=True).items()}
    tmw = ctx.render_frame(rt.default_params(width=W, height=H, mode=rt.RT_MODE_STRICT, aov_mask=rt.RT_AOV_WORK))
    alg = 64 * tmw.inner_visits + 40 * tmw.tri_tests
    for trav, ctas in ((2, 8), (3, 6), (4, 6), (4, 7), (4, 8)):
        p = rt.default_params(width=W, height=H, traversal=trav, ctas_per_sm=ctas)
        for _ in range(3):
            ctx.render_frame(p)
        ms = [ctx.render_frame(p).kernel_ms[0] for _ in range(frames)]
        tm = ctx.render_frame(rt.default_params(width=W, height=H, traversal=trav, ctas_per_sm=ctas, aov_mask=2 | 4))
        got = ctx.load_from_gpu(tri_id=True, depth=True)
        m = O.compare_aovs(got, strict)
        tw = ctx.render_frame(rt.default_params(width=W, height=H, traversal=trav, ctas_per_sm=ctas, aov_mask=rt.RT_AOV_WORK))
        med = statistics.median(ms)
        rays = tm.rays_closest + tm.rays_shadow
        print(json.dumps({"config": 5, "traversal": trav, "ctas": ctas, "kernel_ms": round(med, 3), "min_ms": round(min(ms), 3), "rays": rays,
                          "mrays_s": round(rays / med / 1e3, 1), "id_match": m["id_match"], "rgb8_within1": m["rgb8_within1"], "depth_ok": m["depth_within_1e-4"],
                          "node_visits": tw.inner_visits, "tri_tests": tw.tri_tests,
                          "algorithmic_bytes_reference_order": alg, "algorithmic_gbs": round(alg / med / 1e6, 1)}), flush=True)
    ctx.close(); big.close(); base.close()


if __name__ == "__main__":
    main()
