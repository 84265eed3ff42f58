"""BASELINE config 5 (instanced ~50 M triangles, 4K) on 1/2/4/8 GPUs of one box, one process: scene built and laid out on
device 0 (rt_create_gpu), fanned out over NVLink, tiles interleaved, fused peer stores.  One JSON line per device count."""
import sys, json, time, statistics; sys.path.insert(0, '.')
import numpy as np, parallel_ray_tracer_b200 as rt
nmax = min(rt.device_count(), 8)
base = rt.Scene.load_rtsc("tests/golden/scenes/car_only.rtsc")
want = None
for nd in [n for n in (1, 2, 4, 8) if n <= nmax]:
    big = base.instance_grid(39, 40, 1, (11.5, 6.5, 3.0))
    t0 = time.perf_counter()
    ctx = rt.Context.build_on_gpu(big, list(range(nd)))
    t1 = time.perf_counter()
    p = rt.default_params(width=3840, height=2160)
    for _ in range(3): ctx.render_frame(p)
    tms = [ctx.render_frame(p) for _ in range(7)]
    ms = statistics.median(t.total_ms for t in tms)
    rays = tms[-1].rays_closest + tms[-1].rays_shadow
    frame = ctx.load_from_gpu()["bgra"]
    if want is None: want = frame.copy()
    print(json.dumps({"config": 5, "gpus": nd, "triangles": big.view().n_tris, "triangles_to_context_s": t1 - t0, "frame_ms": ms, "rays": rays,
                      "mrays_s": rays / ms / 1e3, "frame_equals_1gpu": bool(np.array_equal(frame, want)),
                      "kernel_ms_per_device": [round(tms[-1].kernel_ms[d], 3) for d in range(nd)]}), flush=True)
    ctx.close(); big.close()
