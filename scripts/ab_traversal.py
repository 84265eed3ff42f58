#!/usr/bin/env python
"""Same-box A/B of the fast build's traversal variants (2-wide speculative, 4-wide, compressed 8-wide) over the three
workloads and a few occupancies: median kernel ms of N frames after >= 150 ms warm-up each, L2 warm, plus the parity of
every variant against the strict build on that frame.  usage: python scripts/ab_traversal.py [frames] [variant list]
variant = traversal:ctas_per_sm[:cull[:drain_k]], e.g. 2:8 3:6 4:5 4:6:-1:-1 4:6:0:8"""
import json, statistics, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import parallel_ray_tracer_b200 as rt
import oracle as O

CASES = [("car_only", 1920, 1080), ("car_boxed", 1920, 1080), ("car_boxed", 3840, 2160)]


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    variants = [tuple(int(x) for x in v.split(":")) for v in sys.argv[2:]] or [(2, 8), (3, 6), (4, 5), (4, 6), (4, 7), (4, 6, 1, 0), (4, 6, 1, 2), (4, 6, 1, 8)]
    variants = [tuple(v) + (0, 0)[len(v) - 2:] for v in variants]
    for scene, w, h in CASES:
        sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6)
        ctx = rt.Context(sc, [0])
        ctx.render_frame(rt.default_params(width=w, height=h, mode=rt.RT_MODE_STRICT, aov_mask=7))
        strict = ctx.load_from_gpu(rgb=True, tri_id=True, depth=True)
        for rep in range(2):
            for trav, ctas, cull, drain in variants:
                kw = dict(width=w, height=h, traversal=trav, ctas_per_sm=ctas, cull=cull, drain_k=drain)
                p = rt.default_params(**kw)
                t_end = time.perf_counter() + 0.15
                while time.perf_counter() < t_end:
                    ctx.render_frame(p)
                ms = [ctx.render_frame(p).kernel_ms[0] for _ in range(frames)]
                rec = {"scene": scene, "w": w, "traversal": trav, "ctas": ctas, "cull": cull, "drain_k": drain, "rep": rep, "ms": round(statistics.median(ms), 4), "min": round(min(ms), 4)}
                if rep == 0:
                    tm = ctx.render_frame(rt.default_params(**kw, aov_mask=7))
                    got = ctx.load_from_gpu(rgb=True, tri_id=True, depth=True)
                    m = O.compare_aovs(got, strict)
                    rec.update({"id_match": m["id_match"], "rgb8_within1": m["rgb8_within1"], "depth_ok": m["depth_within_1e-4"],
                                "rays": tm.rays_closest + tm.rays_shadow})
                    tw = ctx.render_frame(rt.default_params(**kw, aov_mask=8))
                    rec.update({"node_visits": tw.inner_visits, "tri_tests": tw.tri_tests})
                print(json.dumps(rec), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
