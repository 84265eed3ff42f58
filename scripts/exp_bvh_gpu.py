import sys, time, json; sys.path.insert(0, '.')
import numpy as np, parallel_ray_tracer_b200 as rt
def stats(st): return {k: getattr(st, k) for k, _ in st._fields_}
def run(name, make):
    a = make(); t = time.perf_counter(); a.build_bvh(6); th = (time.perf_counter() - t) * 1e3
    ha = a.arrays(); a.close()
    b = make(); b.build_bvh_gpu(6); b.close()   # warm-up (context creation, module load)
    b = make(); t = time.perf_counter(); st = b.build_bvh_gpu(6); tg = (time.perf_counter() - t) * 1e3
    ga = b.arrays(); b.close()
    same = bool(np.array_equal(ha["bvh_nodes"], ga["bvh_nodes"]) and np.array_equal(ha["tri_idx"], ga["tri_idx"]))
    print(json.dumps({"scene": name, "tris": len(ha["tri"]), "host_ms": th, "gpu_ms": tg, "same": same, **stats(st)}), flush=True)
for sc in ("car_only", "car_boxed"):
    run(sc, lambda: rt.Scene.load_rtsc(f"tests/golden/scenes/{sc}.rtsc"))
run("soup300k", lambda: rt.Scene.soup(300000, 1))
def grid(nx, ny, nz):
    base = rt.Scene.load_rtsc("tests/golden/scenes/car_only.rtsc"); g = base.instance_grid(nx, ny, nz, (6.0, 12.0, 4.0)); base.close(); return g
run("car_only x32 (1.03M)", lambda: grid(4, 4, 2))
if len(sys.argv) > 1: run("car_only x320 (10.3M)", lambda: grid(8, 8, 5))
