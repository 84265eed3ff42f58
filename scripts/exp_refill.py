import sys, json, time, statistics; sys.path.insert(0, '.')
import parallel_ray_tracer_b200 as rt
for scene, w, h in (("car_only", 1920, 1080), ("car_boxed", 1920, 1080), ("car_boxed", 3840, 2160)):
    sc = rt.Scene.load_rtsc(f"tests/golden/scenes/{scene}.rtsc").build_bvh(6); ctx = rt.Context(sc, [0])
    row = {}
    for rep in range(2):
        for refill in (12, 16, 20, 24, 28, 32):
            p = rt.default_params(width=w, height=h, refill_threshold=refill)
            t_end = time.perf_counter() + 0.12
            while time.perf_counter() < t_end: ctx.render_frame(p)
            row.setdefault(refill, []).append(round(statistics.median(ctx.render_frame(p).kernel_ms[0] for _ in range(25)), 3))
    print(scene, w, json.dumps(row), flush=True)
    ctx.close()
