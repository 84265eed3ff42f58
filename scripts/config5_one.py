"""Render a few 4K frames of the 50.1 M-triangle scene (BASELINE.json configs[4]) with the default fast traversal — the
program scripts/r02_profile.sh puts under ncu for the HBM-side counters (dram bytes, L2 / L1 hit rates)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import parallel_ray_tracer_b200 as rt
base = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / "car_only.rtsc")
big = base.instance_grid(39, 40, 1, (11.5, 6.5, 3.0))
ctx = rt.Context.build_on_gpu(big, [0])
trav = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(6):
    tm = ctx.render_frame(rt.default_params(width=3840, height=2160, traversal=trav))
print("kernel_ms", tm.kernel_ms[0], "rays", tm.rays_closest + tm.rays_shadow)
