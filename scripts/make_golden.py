#!/usr/bin/env python
"""Generate tests/golden/ from the REFERENCE ITSELF (oracle/_ref, i.e. the unmodified reference CPU
renderer compiled by oracle/Makefile from /root/reference).  Run in the build container only:

    python scripts/make_golden.py

What it writes (all small, committed):
  scenes/<scene>.rtsc           the scene exactly as the reference's own loader parsed it
                                (triangles_load / lights_load; dumped by ref_harness --dump-scene)
  scenes/soup2k.rtsc            the reference's synthetic generator (cpu/src/main.c:115-131), 2 000 tris
  ref_<scene>_<cam>_<WxH>.npz   per-pixel first-hit ID, t, float RGB and 8-bit BGRA computed by the
                                reference (bvh_traverse / raytrace), heuristic-6 tree
  ref_car_only_64x36.bmp        a BMP written by the reference's bmp_write_file
  manifest.json                 sha256 of the reference-built BVH arrays (heuristic 6, the binary's
                                own tree), node counts, the cameras used, md5 of 1080p BMPs
"""
from __future__ import annotations

import hashlib
import json
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402

GOLD = ROOT / "tests" / "golden"

CAMS = {
    "default": None,  # cpu/src/main.c:105-106
    # the reference's commented-out yaw (main.c:107) plus a closer, lower viewpoint
    "yaw": ((1.5, -7.0, 2.0), (float(np.float32(-np.pi / 14)), 0.0, float(np.float32(np.pi / 10))), float(np.float32(np.pi / 3.2))),
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    O.build_ref()
    ref = O.RefCpu(6)
    assert ref.available, "oracle/_ref is not built"
    (GOLD / "scenes").mkdir(parents=True, exist_ok=True)
    manifest = {"generator": "scripts/make_golden.py", "reference_binary": ref.exe.name, "scenes": {}, "cams": {}}
    for k, v in CAMS.items():
        manifest["cams"][k] = None if v is None else {"pos": v[0], "rot": v[1], "fov": v[2]}

    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        jobs = [("car_only", dict(scene_dir=ref.scene_dir("car_only"))),
                ("car_boxed", dict(scene_dir=ref.scene_dir("car_boxed"))),
                ("soup2k", dict(soup=2000))]
        for name, src in jobs:
            rtsc = GOLD / "scenes" / f"{name}.rtsc"
            bvh = td / f"{name}.bvh"
            info = ref.run(**src, width=8, height=8, frames=0, aov=False, dump_scene=rtsc, dump_bvh=bvh)
            nodes, tri_idx = O.load_bvh_dump(bvh)
            entry = {"triangles": info["triangles"], "lights": info["lights"], "bvh_nodes": info["bvh_nodes"],
                     "bvh_nodes_sha256": sha(nodes), "tri_idx_sha256": sha(tri_idx), "aov": {}}
            for cam_name, cam in CAMS.items():
                for (w, h) in ((160, 90),):
                    r = ref.run(rtsc=rtsc, width=w, height=h, cam=cam)
                    f = GOLD / f"ref_{name}_{cam_name}_{w}x{h}.npz"
                    np.savez_compressed(f, id=r["id"], depth=r["depth"], rgb=r["rgb"], bgra=r["bgra"])
                    entry["aov"][f.name] = {"hit_pixels": int((r["id"] >= 0).sum())}
            if name != "soup2k":
                # full-resolution digest of the reference frame (not compared bit-for-bit: -ffast-math)
                r = ref.run(rtsc=rtsc, width=1920, height=1080)
                entry["ref_1080p"] = {"bgra_sha256": sha(r["bgra"]), "hit_pixels": int((r["id"] >= 0).sum()),
                                      "mean_rgb": [float(x) for x in r["rgb"].reshape(-1, 3).mean(0)]}
            manifest["scenes"][name] = entry

        # a BMP written by the reference's own writer
        exe = str(ref.exe)
        subprocess.run([exe, "--rtsc", str(GOLD / "scenes" / "car_only.rtsc"), "--width", "64", "--height", "36",
                        "--out", str(td / "b"), "--bmp"], check=True, capture_output=True)
        shutil.copy(td / "b.bmp", GOLD / "ref_car_only_64x36.bmp")

        # spp > 1 convention (new surface, include/rt_sampling.h) through the reference's raytrace()
        r = ref.run(rtsc=GOLD / "scenes" / "car_only.rtsc", width=96, height=54, spp=4, seed=7)
        np.savez_compressed(GOLD / "ref_car_only_default_96x54_spp4.npz", rgb=r["rgb"], bgra=r["bgra"])

    (GOLD / "manifest.json").write_text(json.dumps(manifest, indent=1))
    print(json.dumps(manifest, indent=1)[:1500])


if __name__ == "__main__":
    main()
