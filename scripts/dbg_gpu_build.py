"""Debug: device-side build of the 8-wide tree with per-kernel synchronisation (RT_SYNC_DEBUG=1, RT_W8_DEVICE_BUILD=1)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["RT_SYNC_DEBUG"] = "1"
os.environ["RT_W8_DEVICE_BUILD"] = "1"
import numpy as np
import parallel_ray_tracer_b200 as rt
for scene in ("soup2k", "car_only", "car_boxed"):
    sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc")
    try:
        ctx = rt.Context.build_on_gpu(sc, [0])
        host = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6).flatten_host()["nodes8"]
        dev = ctx.device_array(7)
        print(scene, "ok; equal to host:", np.array_equal(dev, host), len(dev), len(host), flush=True)
        if not np.array_equal(dev, host):
            n = min(len(dev), len(host))
            d = np.nonzero(dev[:n] != host[:n])[0]
            print("first diffs at words", d[:10], "node", d[0] // 24 if len(d) else None, "word", d[0] % 24 if len(d) else None, flush=True)
    except Exception as e:
        print(scene, "FAILED:", e, flush=True)
        break
