#!/usr/bin/env python
"""One rank's share of a tile-split frame (part 0 of N) at several occupancies: where do 4- and 8-GPU frames lose their time?"""
import json, statistics, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import parallel_ray_tracer_b200 as rt
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for scene, w, h in (("car_only", 1920, 1080), ("car_boxed", 3840, 2160)):
    sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6)
    ctx = rt.Context(sc, [0])
    for parts in (2, 4, 8):
        for trav, ctas, refill in ((3, 6, 0), (3, 4, 0), (3, 3, 0), (3, 2, 0), (2, 8, 0), (2, 4, 0), (3, 6, 12), (3, 6, 28), (3, 3, 12)):
            p = rt.default_params(width=w, height=h, traversal=trav, ctas_per_sm=ctas, refill_threshold=refill, part_index=0, part_count=parts)
            t_end = time.perf_counter() + 0.12
            while time.perf_counter() < t_end:
                ctx.render_frame(p)
            ms = [ctx.render_frame(p).kernel_ms[0] for _ in range(frames)]
            print(json.dumps({"scene": scene, "w": w, "parts": parts, "traversal": trav, "ctas": ctas, "refill": refill, "ms": round(statistics.median(ms), 4)}), flush=True)
    ctx.close()
