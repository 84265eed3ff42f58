"""Distribution of the per-pixel traversal cost (steps of the 4-wide kernel) — what the heaviest-tiles-first scheduler sees."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import parallel_ray_tracer_b200 as rt
for scene, w, h in (("car_only", 1920, 1080), ("car_boxed", 1920, 1080)):
    sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6)
    ctx = rt.Context(sc, [0])
    for _ in range(3):
        tm = ctx.render_frame(rt.default_params(width=w, height=h, traversal=3))
    cost, hdr = ctx.cost_map(w, h)
    c = cost.reshape(-1).astype(np.int64)
    qs = [50, 90, 99, 99.9, 99.99, 100]
    out = {"scene": scene, "w": w, "kernel_ms": tm.kernel_ms[0], "hdr": hdr.tolist(), "total_steps": int(c.sum()),
           "percentiles": {str(q): int(np.percentile(c, q)) for q in qs}}
    mx = int(c.max())
    for frac in (0.9, 0.7, 0.5, 0.4, 0.3, 0.22, 0.15, 0.1, 0.07, 0.05, 0.03):
        m = c >= frac * mx
        out[f">={frac}max"] = {"pixels": int(m.sum()), "share_of_steps": round(float(c[m].sum() / c.sum()), 4)}
    print(json.dumps(out), flush=True)
    ctx.close()
