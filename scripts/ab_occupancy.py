#!/usr/bin/env python
"""CTAs per SM (4-wide fast kernel) against how chain-bound a frame is: full frames and one rank's share of a tile-split frame.
Prints kernel ms (median) per occupancy, plus the frame's cost statistics: heaviest pixel, total steps, and their ratio to the
steps one lane slot would get if the frame were spread evenly over 148 SMs x 768 lanes."""
import json, statistics, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import parallel_ray_tracer_b200 as rt
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 30
CASES = [("car_only", 1920, 1080, (1, 2, 4, 8)), ("car_only", 1280, 720, (1,)), ("car_boxed", 1920, 1080, (1, 2, 4)), ("car_boxed", 3840, 2160, (4, 8)), ("soup2k", 1920, 1080, (1,))]
for scene, w, h, parts_list in CASES:
    f = ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc"
    if not f.exists(): continue
    sc = rt.Scene.load_rtsc(f).build_bvh(6)
    ctx = rt.Context(sc, [0])
    for parts in parts_list:
        out = {"scene": scene, "w": w, "parts": parts}
        for ctas in (6, 5, 4, 3, 2):
            p = rt.default_params(width=w, height=h, traversal=3, ctas_per_sm=ctas, part_index=0, part_count=parts)
            t_end = time.perf_counter() + 0.12
            while time.perf_counter() < t_end: ctx.render_frame(p)
            out[f"c{ctas}"] = round(statistics.median(ctx.render_frame(p).kernel_ms[0] for _ in range(frames)), 4)
        cost, hdr = ctx.cost_map(w, h)
        own = np.zeros(((h + 7) // 8, (w + 15) // 16), bool).reshape(-1); own[ctx.tile_order()] = True
        m = np.repeat(np.repeat(own.reshape((h + 7) // 8, (w + 15) // 16), 8, 0), 16, 1)[:h, :w]
        c = cost[m].astype(np.int64)
        out["max_cost"], out["total_cost"] = int(c.max()), int(c.sum())
        out["chain_ratio"] = round(float(c.max()) / (c.sum() / (148 * 768)), 2)
        print(json.dumps(out), flush=True)
    ctx.close()
