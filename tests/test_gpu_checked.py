"""The checked build (librt_b200_checked.so, -DRT_DEBUG_BOUNDS=1: every traversal-stack, node, triangle, tile and queue index
is tested on the device, render_kernel.cuh: RT_BCHECK) on the workloads that stress those bounds: the depth-capped soup
(deepest tree, largest leaves), every fast traversal variant incl. culling and the drain kernel, the strict build, and a
partitioned frame.  The substitute for compute-sanitizer, which the GPU pool does not offer.  Runs in a subprocess because
the library is chosen at import time (RT_B200_LIB)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "parallel_ray_tracer_b200" / "librt_b200_checked.so"

CODE = r'''
import sys, json
sys.path.insert(0, %r)
import numpy as np
import parallel_ray_tracer_b200 as rt
out = {}
def frames(ctx, w, h, tag, **kw):
    for mode, trav, extra in [(1, 0, {}), (0, 1, {}), (0, 2, {}), (0, 3, {}), (0, 4, {}), (0, 4, dict(cull=1, drain_k=8)), (0, 4, dict(cull=1, drain_k=32, ctas_per_sm=2))]:
        tm = ctx.render_frame(rt.default_params(width=w, height=h, mode=mode, traversal=trav, aov_mask=7, **extra, **kw))
        out[f"{tag}/{mode}/{trav}/{sorted(extra.items())}"] = tm.rays_closest + tm.rays_shadow
G = %r
for scene, (w, h) in (("car_boxed", (640, 360)), ("soup2k", (320, 180))):
    sc = rt.Scene.load_rtsc(G + f"/{scene}.rtsc").build_bvh(6)
    ctx = rt.Context(sc, [0])
    frames(ctx, w, h, scene)
    frames(ctx, w, h, scene + "/part", part_index=1, part_count=3)
    frames(ctx, 97, 61, scene + "/ragged", spp=3)
    ctx.close(); sc.close()
# deepest tree the reference can build: coincident triangles -> 32 levels, one leaf of 40 (count escape)
rng = np.random.default_rng(5)
one = rng.uniform(-1, 1, (1, 9)).astype(np.float32); one[:, 1::3] += 2.0
heap = np.concatenate([np.repeat(one, 40, axis=0), rng.uniform(-1, 1, (300, 9)).astype(np.float32) + np.float32([0, 2, 0] * 3)])
sc = rt.Scene.from_arrays(heap, np.zeros(len(heap), np.uint32), np.array([[0.2, 0.2, 0.2, 0.7, 0.6, 0.5, 0.4, 0.4, 0.4]], np.float32),
                          np.array([[0, -8, 3, 50, 50, 50]], np.float32)).build_bvh(6)
ctx = rt.Context(sc, [0])
frames(ctx, 200, 120, "heap")
ctx.close()
soup = rt.Scene.soup(20000, 1).build_bvh(6)
ctx = rt.Context.build_on_gpu(rt.Scene.soup(20000, 1), [0])
frames(ctx, 320, 180, "soup20k")
print(json.dumps(out))
'''


def test_checked_build_reports_no_out_of_bounds_index(rt):
    if rt.device_count() < 1:
        pytest.skip("no CUDA device")
    if not LIB.exists():
        pytest.skip("librt_b200_checked.so not built (make -C parallel_ray_tracer_b200/csrc checked)")
    r = subprocess.run([sys.executable, "-c", CODE % (str(ROOT), str(ROOT / "tests" / "golden" / "scenes"))], capture_output=True, text=True,
                       env={**os.environ, "RT_B200_LIB": str(LIB)}, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert len(out) == 8 * 7 and all(v > 0 for v in out.values())
