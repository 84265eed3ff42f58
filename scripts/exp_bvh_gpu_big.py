import sys, time, json; sys.path.insert(0, '.')
import numpy as np, parallel_ray_tracer_b200 as rt
nx, ny, nz = (int(v) for v in sys.argv[1:4])
base = rt.Scene.load_rtsc("tests/golden/scenes/car_only.rtsc"); g = base.instance_grid(nx, ny, nz, (11.5, 6.5, 3.0)); base.close()
small = rt.Scene.soup(1000, 1); small.build_bvh_gpu(6); small.close()  # context + module load
for rep in range(2):
    t = time.perf_counter(); st = g.build_bvh_gpu(6); w = time.perf_counter() - t
    print(json.dumps({"tris": g.view().n_tris, "wall_s": w, **{k: getattr(st, k) for k, _ in st._fields_}}), flush=True)
