#!/usr/bin/env python
"""2-wide against 4-wide tree on the large workloads: car_boxed at 4K and 8K, 8K at 2 spp, and the 50.1 M-triangle scene at 4K
(kernel ms, median of N frames).  usage: python scripts/ab_large.py [frames]"""
import json, statistics, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import parallel_ray_tracer_b200 as rt
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 10

def run(ctx, name, **kw):
    out = {"case": name}
    for trav in (2, 3):
        p = rt.default_params(traversal=trav, **kw)
        for _ in range(3): ctx.render_frame(p)
        out[f"t{trav}"] = round(statistics.median(ctx.render_frame(p).kernel_ms[0] for _ in range(frames)), 3)
    print(json.dumps(out), flush=True)

sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / "car_boxed.rtsc").build_bvh(6)
ctx = rt.Context(sc, [0])
run(ctx, "car_boxed_4k", width=3840, height=2160)
run(ctx, "car_boxed_1440p", width=2560, height=1440)
run(ctx, "car_boxed_8k", width=7680, height=4320)
run(ctx, "car_boxed_8k_spp2", width=7680, height=4320, spp=2)
ctx.close(); sc.close()
base = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / "car_only.rtsc")
run_ctx = rt.Context.build_on_gpu(base.instance_grid(39, 40, 1, (11.5, 6.5, 3.0)), [0])
run(run_ctx, "config5_50M_4k", width=3840, height=2160)
run_ctx.close()
