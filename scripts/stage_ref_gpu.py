#!/usr/bin/env python
"""Stage and build the reference's GPU program ("fuse", the only stage in the tree — SURVEY.md §0.6) as the
one-B200 baseline the north star asks for.  Run in the build container (needs /root/reference):

    python scripts/stage_ref_gpu.py

* copies /root/reference/gpu to baseline/_ref/gpu (git-ignored; nothing of it enters this repository) and the two
  usable scenes to baseline/_ref/assets;
* applies the two-spot patch without which gcc-hosted nvcc cannot compile it: `hvec_t` (anonymous
  structs with __half members) becomes `struct { __half2 xy; __half2 zw; }`, and the unused hvec_* helpers go;
* builds one binary per (scene, resolution) — the reference's configuration is compile-time (options.cuh) — with the
  reference's flags (gpu/makefile:9) and -arch=sm_100.
The kernel, its launch shape protocol (<exe> <tx> <ty>) and its timing (CUDA events around the launch, 50 warm-up + 100
timed frames) are the reference's own.  scripts/bench_ref_gpu.py runs the block-shape sweep of the reference's .bat files.
"""
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"

CONFIGS = [("car_only", 1920, 1080), ("car_boxed", 1920, 1080), ("car_boxed", 3840, 2160)]


def main():
    if not (REF / "gpu").exists():
        print("no /root/reference here: keeping prebuilt baseline/_ref as is")
        return
    g = DST / "gpu"
    if g.exists():
        shutil.rmtree(g)
    shutil.copytree(REF / "gpu", g)
    for s in ("car_only", "car_boxed"):
        (DST / "assets" / s).mkdir(parents=True, exist_ok=True)
        for f in ("triangles.obj", "triangles.mtl", "lights.obj"):
            shutil.copy(REF / "assets" / s / f, DST / "assets" / s / f)
    # patch 1: hvec_t
    vh = g / "include" / "vec.cuh"
    t = vh.read_text()
    t, n1 = re.subn(r"struct hvec_t \{.*?\n\};\n", "struct hvec_t {\n    __half2 xy;\n    __half2 zw;\n};\n", t, count=1, flags=re.S)
    t, n2 = re.subn(r"^__device__ [^\n]*hvec_[a-z]+\([^\n]*\n", "", t, flags=re.M)
    vh.write_text(t)
    # patch 2: unused half helpers
    vc = g / "src" / "vec.cu"
    t = vc.read_text()
    i = t.index("/// __half VEC")
    vc.write_text(t[:i])
    print(f"patched: hvec_t ({n1}), {n2} prototypes, vec.cu truncated at {i}")
    opt = (g / "include" / "options.cuh").read_text()
    srcs = sorted(str(p) for p in (g / "src").glob("*.cu"))
    for scene, w, h in CONFIGS:
        o = re.sub(r'#define SCENE "[^"]*"', f'#define SCENE "{scene}"', opt)
        o = re.sub(r"#define WIDTH \(\d+\)", f"#define WIDTH ({w})", o)
        o = re.sub(r"#define HEIGHT \(\d+\)", f"#define HEIGHT ({h})", o)
        # options.cuh has no include guard and is included by quote from both src/ and include/: rewrite the copy
        (g / "include" / "options.cuh").write_text(o)
        exe = g / f"raytracer_{scene}_{w}x{h}"
        cmd = ["nvcc", "-O3", "-use_fast_math", f"-I{g / 'include'}", "-split-compile=0", "-lineinfo", "-rdc=true", "-dlto",
               *srcs, "-o", str(exe), "-arch=sm_100"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        print(exe.name, "ok" if r.returncode == 0 else "FAILED\n" + r.stderr[-3000:])
        if r.returncode != 0:
            sys.exit(1)


if __name__ == "__main__":
    main()
