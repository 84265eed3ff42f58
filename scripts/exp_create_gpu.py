"""Scene -> render-ready context: host build + host flatten, GPU build + host flatten, all on the device (rt_create_gpu)."""
import sys, time, json; sys.path.insert(0, '.')
import numpy as np, parallel_ray_tracer_b200 as rt
def grid(nx, ny, nz):
    base = rt.Scene.load_rtsc("tests/golden/scenes/car_only.rtsc"); g = base.instance_grid(nx, ny, nz, (11.5, 6.5, 3.0)); base.close(); return g
cases = [("car_only", lambda: rt.Scene.load_rtsc("tests/golden/scenes/car_only.rtsc")), ("car_boxed", lambda: rt.Scene.load_rtsc("tests/golden/scenes/car_boxed.rtsc")),
         ("car_only x32", lambda: grid(4, 4, 2))]
if len(sys.argv) > 1: cases.append(("car_only x1560 (50.1M)", lambda: grid(39, 40, 1)))
w = rt.Scene.soup(2000, 1); c = rt.Context.build_on_gpu(w, [0]); c.close(); w.close()  # context + module load
for name, make in cases:
    r = {"scene": name}
    big = "50.1M" in name
    sc = make(); r["tris"] = sc.view().n_tris
    if not big or len(sys.argv) > 2:
        t = time.perf_counter(); sc.build_bvh(6); t1 = time.perf_counter(); ctx = rt.Context(sc, [0]); t2 = time.perf_counter()
        r["host_build_s"] = t1 - t; r["host_flatten_upload_s"] = t2 - t1
        ctx.render_frame(width=1280, height=720); want = ctx.load_from_gpu()["bgra"].copy(); ctx.close()
    else:
        want = None
    sc.close()
    for rep in range(2):
        sc = make(); t = time.perf_counter(); ctx = rt.Context.build_on_gpu(sc, [0]); t1 = time.perf_counter()
        st = ctx.build_stats
        r[f"device_pipeline_s_{rep}"] = t1 - t; r[f"build_ms_{rep}"] = st.total_ms
        ctx.render_frame(width=1280, height=720); got = ctx.load_from_gpu()["bgra"].copy()
        if want is not None: r["same_image"] = bool(np.array_equal(want, got))
        ctx.close(); sc.close()
    print(json.dumps(r), flush=True)
