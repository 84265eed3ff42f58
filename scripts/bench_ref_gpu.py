#!/usr/bin/env python
"""Run the reference GPU program (baseline/_ref/gpu, built by scripts/stage_ref_gpu.py) on one B200 over a
block-shape sweep in the spirit of its .bat files (gpu/naive.bat: Vx1; gpu/fuse.bat: VxV) and print one JSON
line per (config, shape) plus the best shape per config.  Timing is the reference's own: CUDA events around
the launch, 50 warm-up + 100 timed frames, median (gpu/src/main.cu:111-127)."""
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G = ROOT / "baseline" / "_ref" / "gpu"
RAYS = {("car_only", 1920, 1080): 2978532, ("car_boxed", 1920, 1080): 13247876, ("car_boxed", 3840, 2160): 52998266}
SHAPES = [(32, 1), (64, 1), (128, 1), (256, 1), (4, 4), (8, 4), (8, 8), (16, 4), (16, 8), (16, 16), (32, 4), (32, 8)]


def main():
    out = []
    for exe in sorted(G.glob("raytracer_*")):
        m = re.match(r"raytracer_(\w+)_(\d+)x(\d+)", exe.name)
        scene, w, h = m.group(1), int(m.group(2)), int(m.group(3))
        best = None
        for tx, ty in SHAPES:
            r = subprocess.run([str(exe), str(tx), str(ty)], cwd=G, capture_output=True, text=True, timeout=600)
            mm = re.search(r"Frame time \(median\): ([0-9.]+) ms", r.stdout)
            if r.returncode != 0 or not mm:
                print(json.dumps({"scene": scene, "w": w, "h": h, "tx": tx, "ty": ty, "error": (r.stderr or r.stdout)[-300:]}), flush=True)
                continue
            ms = float(mm.group(1))
            rec = {"impl": "reference gpu/ (fuse)", "scene": scene, "w": w, "h": h, "tx": tx, "ty": ty, "frame_ms_median": ms,
                   "mrays_s": RAYS[(scene, w, h)] / ms / 1e3}
            print(json.dumps(rec), flush=True)
            if best is None or ms < best["frame_ms_median"]:
                best = rec
        if best:
            out.append(best)
            (G / "render.bmp").exists() and (G / "render.bmp").rename(G / f"render_{scene}_{w}x{h}.bmp")
    print(json.dumps({"best": out}), flush=True)


if __name__ == "__main__":
    main()
