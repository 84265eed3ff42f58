"""Experiment: heavy tiles first.  Cost proxy = first-hit pixels per 16x8 tile (depth AOV)."""
import sys; sys.path.insert(0, '.')
import json, time, statistics
import numpy as np, parallel_ray_tracer_b200 as rt
def tile_list(w, h):
    txn, tyn = (w + 15) // 16, (h + 7) // 8
    out = []
    for by in range(0, tyn, 2):
        for bx in range(0, txn, 2):
            for ty in range(by, min(by + 2, tyn)):
                for tx in range(bx, min(bx + 2, txn)):
                    out.append(ty * txn + tx)
    return np.array(out, np.uint32), txn, tyn
def med(ctx, p, n=40):
    t_end = time.perf_counter() + 0.15
    while time.perf_counter() < t_end: ctx.render_frame(p)
    return round(statistics.median(ctx.render_frame(p).kernel_ms[0] for _ in range(n)), 4)
for scene, w, h in (('car_only', 1920, 1080), ('car_boxed', 1920, 1080), ('car_boxed', 3840, 2160)):
    sc = rt.Scene.load_rtsc(f'tests/golden/scenes/{scene}.rtsc').build_bvh(6); ctx = rt.Context(sc, [0])
    ctx.render_frame(rt.default_params(width=w, height=h, aov_mask=2 | 4))
    a = ctx.load_from_gpu(tri_id=True, depth=True)
    base_frame = a["bgra"].copy()
    hit = (a["id"] >= 0)
    tl, txn, tyn = tile_list(w, h)
    pad = np.zeros((tyn * 8, txn * 16), bool); pad[:h, :w] = hit
    cost = pad.reshape(tyn, 8, txn, 16).sum(axis=(1, 3)).reshape(-1)  # per tile id
    p = rt.default_params(width=w, height=h)
    res = {"scene": scene, "w": w, "default": med(ctx, p)}
    c = cost[tl]
    for name, key in (("heavy_first_exact", -c.astype(int)), ("heavy_first_4buckets", -(c.astype(int) * 4 // 129)), ("heavy_first_2buckets", -(c > 0).astype(int)),
                      ("light_first", c.astype(int))):
        order = tl[np.argsort(key, kind="stable")]
        ctx.set_tile_order(order)
        res[name] = med(ctx, p)
        assert np.array_equal(ctx.load_from_gpu()["bgra"], base_frame)
    ctx.set_tile_order(tl)
    res["default_again"] = med(ctx, p)
    print(json.dumps(res), flush=True)
    ctx.close()
