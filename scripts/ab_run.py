#!/usr/bin/env python
"""Same-box A/B of two library builds: alternates them over the three workloads (kernel ms, L2 warm, median of N
frames after >= 150 ms warm-up each).  usage: python scripts/ab_run.py [frames] [extra k=v render params]"""
import json, os, statistics, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
CASES = [("car_only", 1920, 1080), ("car_boxed", 1920, 1080), ("car_boxed", 3840, 2160)]

def child(lib, frames, extra):
    os.environ["RT_B200_LIB"] = lib
    sys.path.insert(0, str(ROOT))
    import parallel_ray_tracer_b200 as rt
    out = {}
    for scene, w, h in CASES:
        sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6)
        ctx = rt.Context(sc, [0])
        p = rt.default_params(width=w, height=h, **extra)
        t_end = time.perf_counter() + 0.15
        while time.perf_counter() < t_end: ctx.render_frame(p)
        ms = [ctx.render_frame(p).kernel_ms[0] for _ in range(frames)]
        out[f"{scene}_{w}"] = round(statistics.median(ms), 4)
        ctx.close()
    print(json.dumps(out))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), json.loads(sys.argv[4]))
    else:
        frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
        extra = {k: int(v) for k, v in (a.split("=") for a in sys.argv[2:])}
        libs = {p.stem.replace("librt_", ""): str(p) for p in sorted((ROOT / "ab").glob("librt_*.so"))}
        for rep in range(2):
            for name, lib in libs.items():
                r = subprocess.run([sys.executable, __file__, "--child", lib, str(frames), json.dumps(extra)], capture_output=True, text=True)
                print(name, r.stdout.strip() or r.stderr[-500:], flush=True)
