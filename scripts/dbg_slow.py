"""Debug: per-frame host enqueue time, wait time and kernel time for the library named by RT_B200_LIB."""
import os, sys, time, statistics, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import parallel_ray_tracer_b200 as rt
scene, w, h = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
extra = {k: int(v) for k, v in (a.split("=") for a in sys.argv[4:])}
sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6)
ctx = rt.Context(sc, [0])
p = rt.default_params(width=w, height=h, **extra)
for _ in range(200): ctx.render_frame(p)
enq, wait, k = [], [], []
for _ in range(30):
    t0 = time.perf_counter(); ctx.render_frame_async(p); t1 = time.perf_counter(); tm = ctx.frame_wait(0); t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e3); wait.append((t2 - t1) * 1e3); k.append(tm.kernel_ms[0])
print(json.dumps({"lib": os.environ.get("RT_B200_LIB", "default"), "scene": scene, "w": w, **extra, "enqueue_ms": round(statistics.median(enq), 4),
                  "wait_ms": round(statistics.median(wait), 4), "kernel_ms": round(statistics.median(k), 4), "launches": tm.launches}), flush=True)
