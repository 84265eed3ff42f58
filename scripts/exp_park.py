"""A/B: shading state in registers (base) vs parked in shared memory (park), over CTAs/SM and traversal layout."""
import json, os, statistics, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
CASES = [("car_only", 1920, 1080), ("car_boxed", 1920, 1080), ("car_boxed", 3840, 2160)]
def child(lib):
    os.environ["RT_B200_LIB"] = lib
    sys.path.insert(0, str(ROOT))
    import parallel_ray_tracer_b200 as rt
    for scene, w, h in CASES:
        sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6)
        ctx = rt.Context(sc, [0])
        for trav in (2, 3):
            row = {}
            for ctas in (6, 7, 8, 9, 10):
                p = rt.default_params(width=w, height=h, traversal=trav, ctas_per_sm=ctas)
                t_end = time.perf_counter() + 0.12
                while time.perf_counter() < t_end: ctx.render_frame(p)
                row[ctas] = round(statistics.median(ctx.render_frame(p).kernel_ms[0] for _ in range(25)), 3)
            print(Path(lib).stem, scene, w, "trav", trav, json.dumps(row), flush=True)
        ctx.close()
if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        for rep in range(2):
            for lib in sorted((ROOT / "ab").glob("librt_*.so")):
                r = subprocess.run([sys.executable, __file__, "--child", str(lib)], capture_output=True, text=True)
                print(r.stdout.strip() or r.stderr[-800:], flush=True)
