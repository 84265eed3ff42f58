#!/usr/bin/env python
"""Same-box A/B of environment settings of ONE library build (kernel ms, median of N frames after >= 150 ms warm-up):
python scripts/ab_env.py frames NAME=v1,v2,... [k=v render params]   — e.g. RT_ADAPTIVE_CTAS=0,1 or RT_W4_LEAF_MAX=2,3,4"""
import json, os, statistics, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
CASES = [("car_only", 1920, 1080, 1), ("car_boxed", 1920, 1080, 1), ("car_only", 1280, 720, 1), ("car_boxed", 3840, 2160, 8), ("car_only", 1920, 1080, 4)]

def child(frames, extra):
    sys.path.insert(0, str(ROOT))
    import parallel_ray_tracer_b200 as rt
    out = {}
    for scene, w, h, parts in CASES:
        sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6)
        ctx = rt.Context(sc, [0])
        p = rt.default_params(width=w, height=h, **extra)
        if parts > 1: p.part_count = parts; p.part_index = 1
        t_end = time.perf_counter() + 0.15
        while time.perf_counter() < t_end: ctx.render_frame(p)
        ms = [ctx.render_frame(p).kernel_ms[0] for _ in range(frames)]
        out[f"{scene}_{w}" + (f"_p{parts}" if parts > 1 else "")] = round(statistics.median(ms), 4)
        ctx.close()
    print(json.dumps(out))

if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]), json.loads(sys.argv[3]))
    else:
        frames = int(sys.argv[1])
        name, vals = sys.argv[2].split("=")
        extra = {k: int(v) for k, v in (a.split("=") for a in sys.argv[3:])}
        for rep in range(2):
            for v in vals.split(","):
                env = dict(os.environ); env[name] = v
                r = subprocess.run([sys.executable, __file__, "--child", str(frames), json.dumps(extra)], capture_output=True, text=True, env=env)
                print(f"{name}={v}", r.stdout.strip() or r.stderr[-500:], flush=True)
