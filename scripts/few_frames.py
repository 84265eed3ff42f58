#!/usr/bin/env python
"""A few frames of one workload with default parameters — the target of the ncu captures in scripts/r02_profile.sh.
usage: python scripts/few_frames.py [scene w h [frames]]"""
import sys; sys.path.insert(0, '.')
import parallel_ray_tracer_b200 as rt
scene, w, h = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("car_only", 1920, 1080)
n = int(sys.argv[4]) if len(sys.argv) > 4 else 6
sc = rt.Scene.load_rtsc(f'tests/golden/scenes/{scene}.rtsc').build_bvh(6); ctx = rt.Context(sc, [0])
p = rt.default_params(width=w, height=h)
for _ in range(n): tm = ctx.render_frame(p)
print("kernel_ms", tm.kernel_ms[0], "rays", tm.rays_closest + tm.rays_shadow)
