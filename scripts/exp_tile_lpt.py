#!/usr/bin/env python
"""Experiment: TILE order by the heaviest pixel of each 16x8 tile (from the per-pixel cost map), against the per-pixel heavy
list (schedule 0) and plain order (schedule -1).  Kernel ms, median of N frames."""
import sys, json, statistics, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import parallel_ray_tracer_b200 as rt
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40

def med(ctx, p):
    t_end = time.perf_counter() + 0.15
    while time.perf_counter() < t_end: ctx.render_frame(p)
    return round(statistics.median(ctx.render_frame(p).kernel_ms[0] for _ in range(frames)), 4)

for scene, w, h in (("car_only", 1920, 1080), ("car_boxed", 1920, 1080), ("car_only", 1280, 720)):
    sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6)
    ctx = rt.Context(sc, [0])
    out = {"scene": scene, "w": w}
    out["pixel_list"] = med(ctx, rt.default_params(width=w, height=h, schedule=0))
    cost, _ = ctx.cost_map(w, h)
    pn = rt.default_params(width=w, height=h, schedule=-1)
    out["plain"] = med(ctx, pn)
    txn, tyn = (w + 15) // 16, (h + 7) // 8
    tl = np.array([ty * txn + tx for by in range(0, tyn, 2) for bx in range(0, txn, 2) for ty in range(by, min(by + 2, tyn)) for tx in range(bx, min(bx + 2, txn))], np.uint32)
    pad = np.zeros((tyn * 8, txn * 16), np.int64); pad[:h, :w] = cost
    tiles = pad.reshape(tyn, 8, txn, 16)
    tmax = tiles.max(axis=(1, 3)).reshape(-1)[tl]
    tsum = tiles.sum(axis=(1, 3)).reshape(-1)[tl]
    cls = np.where(tmax >= 32, np.floor(2 * np.log2(np.maximum(tmax, 32) / 32.0)).astype(int) + 1, 0)
    for name, key in (("tile_max_desc", -tmax), ("tile_maxclass_desc", -cls), ("tile_sum_desc", -tsum)):
        ctx.set_tile_order(tl[np.argsort(key, kind="stable")])
        out[name] = med(ctx, pn)
    print(json.dumps(out), flush=True)
    ctx.close()
