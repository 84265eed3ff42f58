"""TEST INFRASTRUCTURE — Python access to the two CPU checkers.

* ``Oracle``  — ctypes binding of oracle/librt_oracle.so (the committed strict-IEEE C
  restatement, oracle/rt_oracle.c).
* ``RefCpu``  — runs the reference's own CPU renderer, compiled unmodified into
  oracle/_ref/ref_cpu_h{6,3}.{native,v3} by oracle/Makefile (see oracle/ref_harness.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this package.
Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
ORACLE_SO = HERE / "librt_oracle.so"

DEFAULT_CAM_POS = (0.0, -9.0, 3.0)                 # cpu/src/main.c:105
DEFAULT_CAM_ROT = (float(np.float32(-np.pi / 12)), 0.0, 0.0)  # cpu/src/main.c:106
DEFAULT_FOV = float(np.float32(np.pi / 3.2))       # cpu/src/main.c:105


def build_oracle(force: bool = False) -> Path:
    """Compile oracle/rt_oracle.c (gcc, seconds).  Building the checker is not using it."""
    src = HERE / "rt_oracle.c"
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "oracle"], check=True, capture_output=True)
    return ORACLE_SO


def build_ref() -> bool:
    """Compile the reference into oracle/_ref when /root/reference exists (this container)."""
    r = subprocess.run(["make", "-C", str(HERE), "ref"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref build failed:\n" + r.stdout + r.stderr)
    return any(REF_DIR.glob("ref_cpu_h6.*"))


# ---------------------------------------------------------------------------------------
# scene pack (.rtsc) I/O in numpy — layout documented in include/rt_b200.h
# ---------------------------------------------------------------------------------------
def load_rtsc(path) -> dict:
    raw = Path(path).read_bytes()
    if raw[:8] != b"RTSC0001":
        raise ValueError(f"{path}: not a scene pack")
    nt, nm, nl, _ = np.frombuffer(raw, np.uint32, 4, 8)
    amb = np.frombuffer(raw, np.float32, 3, 24).copy()
    off = 40
    tri = np.frombuffer(raw, np.float32, 9 * nt, off).reshape(nt, 9).copy(); off += 36 * nt
    mat_idx = np.frombuffer(raw, np.uint32, nt, off).copy(); off += 4 * nt
    mats = np.frombuffer(raw, np.float32, 9 * nm, off).reshape(nm, 9).copy(); off += 36 * nm
    lights = np.frombuffer(raw, np.float32, 6 * nl, off).reshape(nl, 6).copy()
    return {"tri": tri, "mat_idx": mat_idx, "mats": mats, "lights": lights, "ambient": amb}


def save_rtsc(path, sc: dict) -> None:
    tri = np.ascontiguousarray(sc["tri"], np.float32)
    with open(path, "wb") as f:
        f.write(b"RTSC0001")
        f.write(np.array([len(tri), len(sc["mats"]), len(sc["lights"]), 0], np.uint32).tobytes())
        f.write(np.array([*sc["ambient"], 0], np.float32).tobytes())
        f.write(tri.tobytes())
        f.write(np.ascontiguousarray(sc["mat_idx"], np.uint32).tobytes())
        f.write(np.ascontiguousarray(sc["mats"], np.float32).tobytes())
        f.write(np.ascontiguousarray(sc["lights"], np.float32).tobytes())


def load_bvh_dump(path):
    """File written by ref_cpu --dump-bvh: i32 bvh_len, i32 n_tris, 32-byte nodes, tri_idx."""
    raw = Path(path).read_bytes()
    n, nt = np.frombuffer(raw, np.int32, 2)
    nodes = np.frombuffer(raw, np.uint8, 32 * n, 8).copy()
    tri_idx = np.frombuffer(raw, np.int32, nt, 8 + 32 * n).copy()
    return nodes, tri_idx


def _fp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---------------------------------------------------------------------------------------
class Oracle:
    """ctypes wrapper of oracle/librt_oracle.so."""

    def __init__(self):
        build_oracle()
        L = C.CDLL(str(ORACLE_SO))
        L.ro_scene_create.restype = C.c_void_p
        L.ro_scene_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                      C.c_uint32, C.c_void_p]
        L.ro_scene_free.argtypes = [C.c_void_p]
        L.ro_build_bvh.argtypes = [C.c_void_p, C.c_int]
        L.ro_set_bvh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.ro_bvh_len.argtypes = [C.c_void_p]; L.ro_bvh_len.restype = C.c_int32
        L.ro_bvh_nodes.argtypes = [C.c_void_p]; L.ro_bvh_nodes.restype = C.c_void_p
        L.ro_tri_idx.argtypes = [C.c_void_p]; L.ro_tri_idx.restype = C.c_void_p
        L.ro_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ro_camera_basis.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.ro_hit_triangle.restype = C.c_float
        L.ro_hit_triangle.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.ro_aabb_intersect.restype = C.c_float
        L.ro_aabb_intersect.argtypes = [C.c_void_p] * 4
        L.ro_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        self.L = L

    def scene(self, sc: dict) -> "OracleScene":
        return OracleScene(self, sc)

    def camera_basis(self, pos, rot, fov, w, h):
        out = np.zeros(12, np.float32)
        p = np.asarray(pos, np.float32); r = np.asarray(rot, np.float32)
        self.L.ro_camera_basis(_fp(p), _fp(r), C.c_float(fov), w, h, _fp(out))
        return out.reshape(4, 3)

    def hit_triangle(self, o, d, tri9):
        o = np.asarray(o, np.float32); d = np.asarray(d, np.float32); t9 = np.asarray(tri9, np.float32)
        nd = C.c_int(0)
        t = self.L.ro_hit_triangle(_fp(o), _fp(d), _fp(t9), C.byref(nd))
        return np.float32(t), nd.value

    def aabb_intersect(self, bmin, bmax, o, d):
        a = [np.asarray(v, np.float32) for v in (bmin, bmax, o, d)]
        return np.float32(self.L.ro_aabb_intersect(*[_fp(v) for v in a]))


class OracleScene:
    def __init__(self, orc: Oracle, sc: dict):
        self.L = orc.L
        self.sc = sc
        tri = np.ascontiguousarray(sc["tri"], np.float32)
        mi = np.ascontiguousarray(sc["mat_idx"], np.uint32)
        mats = np.ascontiguousarray(sc["mats"], np.float32)
        li = np.ascontiguousarray(sc["lights"], np.float32).reshape(-1, 6)
        amb = np.ascontiguousarray(sc["ambient"], np.float32)
        self.n_tris = len(tri)
        self.h = self.L.ro_scene_create(_fp(tri), _fp(mi), len(tri), _fp(mats), len(mats), _fp(li), len(li), _fp(amb))

    def __del__(self):
        try:
            self.L.ro_scene_free(self.h)
        except Exception:
            pass

    def build_bvh(self, heuristic=6):
        if self.L.ro_build_bvh(self.h, heuristic) != 0:
            raise RuntimeError("ro_build_bvh failed")
        return self.bvh()

    def set_bvh(self, nodes_u8, tri_idx):
        nodes_u8 = np.ascontiguousarray(nodes_u8, np.uint8); tri_idx = np.ascontiguousarray(tri_idx, np.int32)
        self.L.ro_set_bvh(self.h, _fp(nodes_u8), _fp(tri_idx), len(nodes_u8) // 32)

    def bvh(self):
        n = self.L.ro_bvh_len(self.h)
        nodes = np.ctypeslib.as_array(C.cast(self.L.ro_bvh_nodes(self.h), C.POINTER(C.c_uint8)), (n * 32,)).copy()
        ti = np.ctypeslib.as_array(C.cast(self.L.ro_tri_idx(self.h), C.POINTER(C.c_int32)), (self.n_tris,)).copy()
        return nodes, ti

    def render(self, width, height, pos=DEFAULT_CAM_POS, rot=DEFAULT_CAM_ROT, fov=DEFAULT_FOV, spp=1, seed=1,
               bounces=4, threads=None):
        threads = threads or os.cpu_count() or 1
        n = width * height
        rgb = np.zeros((height, width, 3), np.float32); bgra = np.zeros((height, width, 4), np.uint8)
        tid = np.zeros((height, width), np.int32); dep = np.zeros((height, width), np.float32)
        cnt = np.zeros(5, np.uint64)
        p = np.asarray(pos, np.float32); r = np.asarray(rot, np.float32)
        rc = self.L.ro_render(self.h, _fp(p), _fp(r), C.c_float(fov), width, height, spp, seed, bounces, threads,
                              _fp(rgb), _fp(bgra), _fp(tid), _fp(dep), _fp(cnt))
        if rc != 0:
            raise RuntimeError("ro_render failed (BVH built?)")
        return {"rgb": rgb, "bgra": bgra, "id": tid, "depth": dep,
                "rays_closest": int(cnt[0]), "rays_shadow": int(cnt[1]), "inner_visits": int(cnt[2]),
                "tri_tests": int(cnt[3]), "box_tests": int(cnt[4])}

    def trace_closest(self, o, d):
        o = np.asarray(o, np.float32); d = np.asarray(d, np.float32)
        t = C.c_float(0); nd = C.c_int(0)
        i = self.L.ro_trace_closest(self.h, _fp(o), _fp(d), C.byref(t), C.byref(nd))
        return i, np.float32(t.value), nd.value


# ---------------------------------------------------------------------------------------
class RefCpu:
    """The reference's own CPU renderer (oracle/_ref), run as a subprocess."""

    def __init__(self, heuristic=6):
        self.exe = None
        for isa in ("native", "v3"):
            exe = REF_DIR / f"ref_cpu_h{heuristic}.{isa}"
            if not exe.exists():
                continue
            try:  # -march=native of the build container may SIGILL on this host: self-test
                r = subprocess.run([str(exe), "--soup", "64", "--width", "8", "--height", "8", "--frames", "1"],
                                   capture_output=True, timeout=60)
                if r.returncode == 0:
                    self.exe = exe
                    self.isa = isa
                    break
            except Exception:
                continue
        self.heuristic = heuristic

    @property
    def available(self) -> bool:
        return self.exe is not None

    @staticmethod
    def scene_dir(name: str) -> Path:
        return REF_DIR / "assets" / name

    def run(self, *, scene_dir=None, rtsc=None, soup=0, width=1920, height=1080, spp=1, seed=1, threads=None,
            frames=1, warmup=0, cam=None, aov=True, dump_bvh=None, dump_scene=None, timeout=3600):
        if not self.available:
            raise RuntimeError("oracle/_ref reference binary not available")
        threads = threads or os.cpu_count() or 1
        cmd = [str(self.exe), "--width", str(width), "--height", str(height), "--spp", str(spp), "--seed", str(seed),
               "--threads", str(threads), "--frames", str(frames), "--warmup", str(warmup)]
        if scene_dir is not None:
            cmd += ["--scene-dir", str(scene_dir)]
        elif rtsc is not None:
            cmd += ["--rtsc", str(rtsc)]
        else:
            cmd += ["--soup", str(soup)]
        if cam is not None:
            pos, rot, fov = cam
            cmd += ["--cam", *[repr(float(v)) for v in (*pos, *rot, fov)]]
        if dump_bvh:
            cmd += ["--dump-bvh", str(dump_bvh)]
        if dump_scene:
            cmd += ["--dump-scene", str(dump_scene)]
        out = {}
        with tempfile.TemporaryDirectory() as td:
            if aov:
                cmd += ["--out", os.path.join(td, "f")]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
            if r.returncode != 0:
                raise RuntimeError(f"reference run failed: {r.stderr[-2000:]}")
            out.update(json.loads(r.stdout.strip().splitlines()[-1]))
            if aov and frames + warmup > 0:
                p = os.path.join(td, "f")
                out["id"] = np.fromfile(p + ".id.i32", np.int32).reshape(height, width)
                out["depth"] = np.fromfile(p + ".t.f32", np.float32).reshape(height, width)
                out["rgb"] = np.fromfile(p + ".rgb.f32", np.float32).reshape(height, width, 3)
                out["bgra"] = np.fromfile(p + ".bgra.u8", np.uint8).reshape(height, width, 4)
        return out


# ---------------------------------------------------------------------------------------
def compare_aovs(test: dict, ref: dict) -> dict:
    """The north-star parity metrics (BASELINE.json): fraction of pixels with matching first-hit
    ID, fraction with 8-bit RGB within 1 LSB, fraction with depth within 1e-4 relative."""
    res = {}
    n = ref["id"].size
    if "id" in test:
        res["id_match"] = float((test["id"] == ref["id"]).sum()) / n
    if "bgra" in test:
        d = np.abs(test["bgra"].astype(np.int16) - ref["bgra"].astype(np.int16)).max(axis=-1)
        res["rgb8_within1"] = float((d <= 1).sum()) / n
        res["rgb8_exact"] = float((d == 0).sum()) / n
        res["rgb8_maxdiff"] = int(d.max())
    if "depth" in test:
        a = test["depth"].astype(np.float64); b = ref["depth"].astype(np.float64)
        ok = np.abs(a - b) <= 1e-4 * np.abs(b)
        res["depth_within_1e-4"] = float(ok.sum()) / n
        res["depth_exact"] = float((test["depth"] == ref["depth"]).sum()) / n
    return res
