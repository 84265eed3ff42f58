"""Where does the pipelined e2e loop lose time?  kernel ms inside the loop vs wall per frame, with/without the copy."""
import sys; sys.path.insert(0, '.')
import json, time, statistics
import numpy as np, parallel_ray_tracer_b200 as rt
w, h, n = 1920, 1080, 60
sc = rt.Scene.load_rtsc('tests/golden/scenes/car_only.rtsc').build_bvh(6); ctx = rt.Context(sc, [0])
p = [rt.default_params(width=w, height=h, frame_slot=s) for s in range(2)]
bufs = [rt.PinnedBuffer(w * h * 4) for _ in range(2)]
for _ in range(50): ctx.render_frame(p[0])
def run(copy, wait_first=True):
    km = []
    t0 = time.perf_counter()
    for k in range(n):
        s = k & 1
        if k >= 2: km.append(ctx.frame_wait(s).kernel_ms[0])
        ctx.render_frame_async(p[s])
        if copy: ctx.download_async(s, bufs[s].ptr)
    for s in range(2): km.append(ctx.frame_wait(s).kernel_ms[0])
    return {"copy": copy, "wall_ms_per_frame": (time.perf_counter() - t0) * 1e3 / n, "kernel_ms_median": statistics.median(km), "kernel_ms_max": max(km)}
for copy in (False, True, False, True):
    print(json.dumps(run(copy)), flush=True)
# host-side cost of one enqueue
t0 = time.perf_counter()
for k in range(2):
    ctx.render_frame_async(p[k])
t1 = time.perf_counter()
for k in range(2): ctx.frame_wait(k)
print(json.dumps({"enqueue_ms_each": (t1 - t0) * 1e3 / 2}))
sync = []
t0 = time.perf_counter()
for k in range(n): sync.append(ctx.render_frame(p[0]).kernel_ms[0])
print(json.dumps({"sync_wall_ms_per_frame": (time.perf_counter() - t0) * 1e3 / n, "kernel_ms_median": statistics.median(sync)}))
