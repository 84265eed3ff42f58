"""Per-warp timeline of one frame (rt_debug_warp_trace): where does the tail of a frame go?
usage: python scripts/trace_tail.py [scene w h [traversal [ctas]]]"""
import sys; sys.path.insert(0, '.')
import json
import numpy as np, parallel_ray_tracer_b200 as rt
scene, w, h = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ('car_only', 1920, 1080)
trav = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ctas = int(sys.argv[5]) if len(sys.argv) > 5 else 0
order_mode = sys.argv[6] if len(sys.argv) > 6 else "raster"   # raster | heavy_first | light_first
sched = int(sys.argv[7]) if len(sys.argv) > 7 else 0           # rt_render_params.schedule (0 = heaviest tiles first, -1 = spatial order)
parts = int(sys.argv[8]) if len(sys.argv) > 8 else 1           # one rank's share of a tile-split frame: part_index of part_count
part = int(sys.argv[9]) if len(sys.argv) > 9 else 0
sc = rt.Scene.load_rtsc(f'tests/golden/scenes/{scene}.rtsc').build_bvh(6); ctx = rt.Context(sc, [0])
if order_mode != "raster":
    # cost proxy per 16x8 tile: first-hit pixels (scripts/exp_tile_order.py)
    ctx.render_frame(rt.default_params(width=w, height=h, aov_mask=2, traversal=trav, ctas_per_sm=ctas))
    hit = ctx.load_from_gpu(tri_id=True)["id"] >= 0
    txn, tyn = (w + 15) // 16, (h + 7) // 8
    tl = np.array([ty * txn + tx for by in range(0, tyn, 2) for bx in range(0, txn, 2) for ty in range(by, min(by + 2, tyn)) for tx in range(bx, min(bx + 2, txn))], np.uint32)
    pad = np.zeros((tyn * 8, txn * 16), bool); pad[:h, :w] = hit
    cost = pad.reshape(tyn, 8, txn, 16).sum(axis=(1, 3)).reshape(-1)[tl].astype(int)
    ctx.set_tile_order(tl[np.argsort(-cost if order_mode == "heavy_first" else cost, kind="stable")])
ctx.warp_trace(True)
p = rt.default_params(width=w, height=h, aov_mask=rt.RT_AOV_WORK, traversal=trav, ctas_per_sm=ctas, schedule=sched, part_count=parts, part_index=part)
pn = rt.default_params(width=w, height=h, traversal=trav, ctas_per_sm=ctas, schedule=sched, part_count=parts, part_index=part)
for _ in range(30): ctx.render_frame(pn)
plain = np.median([ctx.render_frame(pn).kernel_ms[0] for _ in range(20)])
for _ in range(4): tm = ctx.render_frame(p)
t = ctx.warp_trace(True).astype(np.float64)
t0 = t[:, 0].min(); start = t[:, 0] - t0; empty = t[:, 1] - t0; exit_ = t[:, 2] - t0
dur = exit_.max()
out = {"order": order_mode, "schedule": sched, "parts": parts, "part": part, "scene": scene, "w": w, "h": h, "trav": trav, "kernel_ms_plain": float(plain), "kernel_ms_traced": float(tm.kernel_ms[0]), "warps": len(t), "dur_ns": dur,
       "queue_empty_ns": {"min": empty[empty > 0].min(), "median": float(np.median(empty[empty > 0]))},
       "exit_ns": {"mean": exit_.mean(), "p50": float(np.median(exit_)), "p90": float(np.percentile(exit_, 90)), "p99": float(np.percentile(exit_, 99)), "max": dur},
       "iters": {"mean": t[:, 4].mean(), "p90": float(np.percentile(t[:, 4], 90)), "max": t[:, 4].max(), "sum": t[:, 4].sum()},
       "lanes_per_iter": (t[:, 5].sum() + t[:, 6].sum()) / max(1, t[:, 4].sum()),
       "exit_hist_20": np.histogram(exit_ / dur, bins=20, range=(0, 1))[0].tolist()}
late = np.argsort(-exit_)[:24]
out["latest"] = [{"exit": exit_[i] / dur, "empty": empty[i] / dur, "iters": t[i, 4], "chunks": t[i, 3], "sm": int(t[i, 7]),
                  "ns_per_iter_overall": exit_[i] / max(t[i, 4], 1), "lanes_per_iter": (t[i, 5] + t[i, 6]) / max(t[i, 4], 1)} for i in late]
# how many warps are still alive on the SMs of the latest warps during the tail
sm = t[:, 7].astype(int)
alive = []
for frac in (0.4, 0.5, 0.6, 0.7, 0.8, 0.9):
    a = exit_ > frac * dur
    per_sm = np.bincount(sm[a], minlength=sm.max() + 1)
    alive.append({"at": frac, "warps_alive": int(a.sum()), "sms_with_work": int((per_sm > 0).sum()), "max_per_sm": int(per_sm.max())})
out["alive"] = alive
print(json.dumps(out, default=float))
