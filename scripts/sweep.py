#!/usr/bin/env python
"""Launch-parameter sweep of the render kernel on one GPU (the analogue of the reference's .bat block
sweeps, gpu/*.bat): prints median kernel ms (L2 warm, CUDA events) per configuration as JSON lines."""
import itertools
import json
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import parallel_ray_tracer_b200 as rt  # noqa: E402


def clocks():
    import subprocess
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"],
                              capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return str(e)


def main():
    print(json.dumps({"clocks_idle": clocks()}), flush=True)
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    cases = [("car_only", 1920, 1080), ("car_boxed", 1920, 1080), ("car_boxed", 3840, 2160)]
    for scene, w, h in cases:
        sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6)
        ctx = rt.Context(sc, [0])
        for mode in (rt.RT_MODE_STRICT, rt.RT_MODE_FAST):
            grid = [(b, c, r, t, 0) for rep in (0, 1) for (b, c) in ((128, 6), (128, 7), (64, 12)) for t in (2, 3) for r in (12, 16, 20, 24, 28)]
            if mode == rt.RT_MODE_STRICT:
                grid = [(128, 5, 24, 1, 2), (128, 5, 24, 1, 2)]
            for block, ctas, refill, trav, fb in grid:
                p = rt.default_params(width=w, height=h, mode=mode, block_threads=block, ctas_per_sm=ctas, refill_threshold=refill, traversal=trav)
                ms = []
                import time
                t_end = time.perf_counter() + 0.15   # >= 150 ms of warm-up per configuration (clock ramp)
                while time.perf_counter() < t_end:
                    ctx.render_frame(p)
                for i in range(frames):
                    tm = ctx.render_frame(p)
                    ms.append(tm.kernel_ms[0])
                rays = tm.rays_closest + tm.rays_shadow
                med = statistics.median(ms)
                print(json.dumps({"clocks": clocks(), "scene": scene, "w": w, "h": h, "mode": "strict" if mode else "fast", "block": block, "ctas_per_sm": ctas,
                                  "refill": refill, "traversal": trav, "stack": fb, "kernel_ms_median": round(med, 4), "kernel_ms_min": round(min(ms), 4),
                                  "mrays_s": round(rays / med / 1e3, 1), "rays": rays}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
