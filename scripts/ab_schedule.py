#!/usr/bin/env python
"""Same-box A/B of heaviest-tiles-first scheduling (schedule = 0) against spatial tile order (schedule = -1): median kernel ms
of N frames after warm-up (the timed window includes the two ordering kernels), per workload and traversal."""
import json, os, statistics, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import parallel_ray_tracer_b200 as rt

CASES = [("car_only", 1920, 1080, 1), ("car_boxed", 1920, 1080, 1), ("car_only", 1920, 1080, 4), ("car_boxed", 3840, 2160, 8), ("car_only", 1280, 720, 1)]
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for scene, w, h, parts in CASES:
    sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6)
    ctx = rt.Context(sc, [0])
    for rep in range(2):
        for trav in (3,):
            for sched in (-1, 0):
                p = rt.default_params(width=w, height=h, traversal=trav, schedule=sched, part_index=0, part_count=parts)
                t_end = time.perf_counter() + 0.15
                while time.perf_counter() < t_end:
                    ctx.render_frame(p)
                ms = [ctx.render_frame(p).kernel_ms[0] for _ in range(frames)]
                print(json.dumps({"scene": scene, "w": w, "h": h, "part_count": parts, "traversal": trav, "schedule": sched, "rep": rep,
                                  "ms": round(statistics.median(ms), 4), "min": round(min(ms), 4)}), flush=True)
    ctx.close()
