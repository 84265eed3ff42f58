#!/usr/bin/env python
"""50.1 M triangles -> render-ready context (rt_create_gpu) several times in one process: wall seconds and the build's own stage
timers.  usage: python scripts/config5_create.py [repeats]"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import parallel_ray_tracer_b200 as rt
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
base = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / "car_only.rtsc")
warm = rt.Context.build_on_gpu(rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / "car_only.rtsc"), [0]); warm.close()   # CUDA context, module load
big = base.instance_grid(39, 40, 1, (11.5, 6.5, 3.0))
for k in range(reps):
    t0 = time.perf_counter()
    ctx = rt.Context.build_on_gpu(big, [0])
    t1 = time.perf_counter()
    st = ctx.build_stats
    print(json.dumps({"repeat": k, "triangles": 39 * 40 * 32136, "triangles_to_context_s": round(t1 - t0, 3),
                      "gpu_build_ms": {k2: round(getattr(st, k2), 1) for k2, _ in st._fields_ if k2.endswith("_ms")}}), flush=True)
    ctx.close()
