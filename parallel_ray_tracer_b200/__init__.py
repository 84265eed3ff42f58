"""parallel_ray_tracer_b200 — ctypes binding of librt_b200.so (C-ABI in include/rt_b200.h).

Python here is plumbing for tests and bench.py only: every scene, BVH, kernel and frame
operation is native (C++ / CUDA in csrc/).  Names mirror the reference's free functions:

    reference (gpu/include/gpu.cuh:23-26)        here
    -----------------------------------------    -------------------------------------
    triangles_load / lights_load / bvh_build     Scene.load_dir / Scene.build_bvh
    load_to_gpu()                                Context(scene, devices)      (rt_create)
    render_frame(is_metrics, tx, ty) -> ms       Context.render_frame(...)    (rt_render)
    load_from_gpu()  (fills pixels[])            Context.load_from_gpu(...)   (rt_download)
    bmp_write_file(pixels, w, h, name)           write_bmp(path, bgra)

There is no CPU fallback: importing works without a GPU (so the CPU test tier can check the
library loads and exports its symbols), but Context() raises RtError(RT_ERR_NO_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("RT_B200_LIB", PKG_DIR / "librt_b200.so"))  # override: A/B runs of two builds

RT_OK, RT_ERR_INVALID, RT_ERR_IO, RT_ERR_NO_DEVICE, RT_ERR_CUDA, RT_ERR_NOMEM, RT_ERR_STATE = 0, -1, -2, -3, -4, -5, -6
RT_MODE_FAST, RT_MODE_STRICT = 0, 1
RT_AOV_RGB_F32, RT_AOV_TRI_ID, RT_AOV_DEPTH, RT_AOV_WORK = 1, 2, 4, 8
RT_GATHER_PEER_STORE, RT_GATHER_PEER_COPY = 0, 1
RT_BVH_REFBIN = 0x100
RT_TRAVERSAL_DEFAULT, RT_TRAVERSAL_PLAIN, RT_TRAVERSAL_SPECULATIVE, RT_TRAVERSAL_WIDE, RT_TRAVERSAL_WIDE8 = 0, 1, 2, 3, 4
RT_TILE_W, RT_TILE_H = 16, 8
RT_MAX_DEVICES = 16
RT_FRAME_SLOTS = 2
RT_FRAME_BOTTOM_UP = 1

# every symbol include/rt_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "rt_scene_load_obj", "rt_scene_load_dir", "rt_scene_load_rtsc", "rt_scene_save_rtsc", "rt_scene_from_arrays",
    "rt_scene_soup", "rt_scene_instance_grid", "rt_scene_build_bvh", "rt_scene_view", "rt_scene_free",
    "rt_render_params_default", "rt_create", "rt_render", "rt_download", "rt_destroy", "rt_last_error",
    "rt_part_tile_count", "rt_packed_tiles", "rt_unpack_tiles", "rt_frame_ipc_export", "rt_frame_ipc_import",
    "rt_frame_device_ptr", "rt_write_bmp", "rt_abi_version", "rt_device_count", "rt_debug_warp_trace",
    "rt_render_async", "rt_download_async", "rt_frame_wait", "rt_host_alloc", "rt_host_free",
    "rt_frame_ipc_export_slot", "rt_frame_ipc_import_slot", "rt_write_bmp_bottom_up", "rt_debug_set_tile_order", "rt_scene_build_bvh_gpu", "rt_debug_gather_bandwidth", "rt_create_gpu", "rt_debug_flatten_host", "rt_debug_device_array", "rt_debug_copy_bandwidth", "rt_debug_cost_map", "rt_debug_tile_order",
]


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rt_b200 error {code}: {msg}")
        self.code = code


class rt_bvh_node(C.Structure):
    _fields_ = [("min", C.c_float * 3), ("max", C.c_float * 3), ("tr_len", C.c_int32), ("idx", C.c_int32)]


class rt_scene_desc(C.Structure):
    _fields_ = [("tri_coords", C.c_void_p), ("tri_mat", C.c_void_p), ("n_tris", C.c_uint32),
                ("materials", C.c_void_p), ("n_mats", C.c_uint32), ("lights", C.c_void_p), ("n_lights", C.c_uint32),
                ("ambient", C.c_float * 3), ("bvh", C.c_void_p), ("tri_idx", C.c_void_p), ("bvh_len", C.c_uint32)]


class rt_camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("rot", C.c_float * 3), ("fov", C.c_float)]


class rt_render_params(C.Structure):
    _fields_ = [("cam", rt_camera), ("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32),
                ("seed", C.c_uint32), ("bounces", C.c_int32), ("mode", C.c_int32), ("aov_mask", C.c_int32),
                ("gather", C.c_int32), ("part_index", C.c_int32), ("part_count", C.c_int32),
                ("block_threads", C.c_int32), ("ctas_per_sm", C.c_int32), ("refill_threshold", C.c_int32),
                ("traversal", C.c_int32), ("frame_flags", C.c_int32), ("frame_slot", C.c_int32),
                ("drain_k", C.c_int32), ("cull", C.c_int32), ("schedule", C.c_int32), ("reserved", C.c_int32)]


class rt_timing(C.Structure):
    _fields_ = [("kernel_ms", C.c_float * RT_MAX_DEVICES), ("gather_ms", C.c_float), ("total_ms", C.c_float),
                ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("inner_visits", C.c_uint64),
                ("tri_tests", C.c_uint64), ("launches", C.c_uint32), ("n_devices", C.c_uint32)]


class rt_bvh_gpu_stats(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("upload_ms", C.c_float), ("top_ms", C.c_float), ("subtree_ms", C.c_float),
                ("assemble_ms", C.c_float), ("download_ms", C.c_float), ("levels", C.c_int32), ("top_nodes", C.c_int32),
                ("subtrees", C.c_int32), ("fell_back", C.c_int32), ("nodes", C.c_uint32)]


def build(verbose: bool = False) -> Path:
    """Compile csrc/ for sm_100a into librt_b200.so (in-tree; nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(PKG_DIR / "csrc"), "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building librt_b200.so failed")
    # the checked twin (RT_DEBUG_BOUNDS, tests/test_gpu_checked.py)
    r = subprocess.run(["make", "-C", str(PKG_DIR / "csrc"), "-j", str(os.cpu_count() or 4), "checked"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building librt_b200_checked.so failed")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RtError(RT_ERR_STATE, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(there is no Python/CPU fallback for the render path)")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint32
    L.rt_scene_load_obj.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(vp)]
    L.rt_scene_load_dir.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.rt_scene_load_rtsc.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.rt_scene_save_rtsc.argtypes = [vp, C.c_char_p]
    L.rt_scene_from_arrays.argtypes = [C.POINTER(rt_scene_desc), C.POINTER(vp)]
    L.rt_scene_soup.argtypes = [u32, u32, C.POINTER(vp)]
    L.rt_scene_instance_grid.argtypes = [vp, u32, u32, u32, C.POINTER(C.c_float), u32, C.POINTER(vp)]
    L.rt_scene_build_bvh.argtypes = [vp, i32]
    L.rt_scene_build_bvh_gpu.argtypes = [vp, i32, i32, C.POINTER(rt_bvh_gpu_stats)]
    L.rt_scene_view.argtypes = [vp, C.POINTER(rt_scene_desc)]
    L.rt_scene_free.argtypes = [vp]; L.rt_scene_free.restype = None
    L.rt_render_params_default.argtypes = [C.POINTER(rt_render_params)]; L.rt_render_params_default.restype = None
    L.rt_create.argtypes = [C.POINTER(rt_scene_desc), C.POINTER(i32), i32, C.POINTER(vp)]
    L.rt_debug_flatten_host.argtypes = [C.POINTER(rt_scene_desc), i32, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(i32)]
    L.rt_create_gpu.argtypes = [vp, i32, C.POINTER(i32), i32, i32, C.POINTER(vp), C.POINTER(rt_bvh_gpu_stats)]
    if hasattr(L, "rt_debug_cost_map"):
        L.rt_debug_cost_map.argtypes = [vp, vp, C.c_size_t, vp]
    if hasattr(L, "rt_debug_tile_order"):
        L.rt_debug_tile_order.argtypes = [vp, C.c_int, vp, C.c_int, vp]
    if hasattr(L, "rt_debug_copy_bandwidth"):
        L.rt_debug_copy_bandwidth.argtypes = [i32, C.c_size_t, i32, C.POINTER(C.c_float)]
    if hasattr(L, "rt_debug_device_array"):  # (absent from older builds loaded through RT_B200_LIB for A/B runs)
        L.rt_debug_device_array.argtypes = [vp, i32, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.rt_render.argtypes = [vp, C.POINTER(rt_render_params), C.POINTER(rt_timing)]
    L.rt_download.argtypes = [vp, vp, vp, vp, vp]
    L.rt_destroy.argtypes = [vp]; L.rt_destroy.restype = None
    L.rt_last_error.argtypes = [vp]; L.rt_last_error.restype = C.c_char_p
    L.rt_part_tile_count.argtypes = [i32, i32, i32, i32]
    L.rt_packed_tiles.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.rt_unpack_tiles.argtypes = [vp, vp, C.c_size_t, i32]
    L.rt_frame_ipc_export.argtypes = [vp, i32, i32, vp]
    L.rt_frame_ipc_import.argtypes = [vp, vp, i32, i32]
    L.rt_frame_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.rt_write_bmp.argtypes = [C.c_char_p, vp, i32, i32]
    L.rt_debug_warp_trace.argtypes = [vp, i32, vp, i32]
    L.rt_debug_set_tile_order.argtypes = [vp, vp, i32]
    L.rt_debug_gather_bandwidth.argtypes = [i32, C.c_size_t, C.POINTER(C.c_float)]
    L.rt_render_async.argtypes = [vp, C.POINTER(rt_render_params)]
    L.rt_download_async.argtypes = [vp, i32, vp]
    L.rt_frame_wait.argtypes = [vp, i32, C.POINTER(rt_timing)]
    L.rt_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.rt_host_free.argtypes = [vp]; L.rt_host_free.restype = None
    L.rt_frame_ipc_export_slot.argtypes = [vp, i32, i32, i32, vp]
    L.rt_frame_ipc_import_slot.argtypes = [vp, i32, vp, i32, i32]
    L.rt_write_bmp_bottom_up.argtypes = [C.c_char_p, vp, i32, i32]
    _lib = L
    return L


def _check(rc, ctx=None):
    if rc != 0:
        msg = lib().rt_last_error(ctx)
        raise RtError(rc, msg.decode(errors="replace") if msg else "")


def device_count() -> int:
    return lib().rt_device_count()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ------------------------------------------------------------------------------------------
class Scene:
    """Library-owned host scene (rt_scene)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    # -- constructors (triangles_load / lights_load, cpu/src/triangle.c:74-126, cpu/src/light.c:6-29)
    @classmethod
    def load_dir(cls, path):
        h = C.c_void_p()
        _check(lib().rt_scene_load_dir(str(path).encode(), C.byref(h)))
        return cls(h.value)

    @classmethod
    def load_obj(cls, obj, mtl, lights=None):
        h = C.c_void_p()
        _check(lib().rt_scene_load_obj(str(obj).encode(), str(mtl).encode(), str(lights).encode() if lights else None, C.byref(h)))
        return cls(h.value)

    @classmethod
    def load_rtsc(cls, path):
        h = C.c_void_p()
        _check(lib().rt_scene_load_rtsc(str(path).encode(), C.byref(h)))
        return cls(h.value)

    @classmethod
    def soup(cls, n_tris, seed=1):
        h = C.c_void_p()
        _check(lib().rt_scene_soup(n_tris, seed, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_arrays(cls, tri, mat_idx, mats, lights, ambient=(0.5, 0.5, 0.5), bvh_nodes=None, tri_idx=None):
        tri = np.ascontiguousarray(tri, np.float32).reshape(-1, 9)
        mat_idx = np.ascontiguousarray(mat_idx, np.uint32)
        mats = np.ascontiguousarray(mats, np.float32).reshape(-1, 9)
        lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 6)
        d = rt_scene_desc()
        d.tri_coords, d.tri_mat, d.n_tris = _ptr(tri), _ptr(mat_idx), len(tri)
        d.materials, d.n_mats = _ptr(mats), len(mats)
        d.lights, d.n_lights = (_ptr(lights) if len(lights) else None), len(lights)
        d.ambient[:] = [float(a) for a in ambient]
        if bvh_nodes is not None:
            bvh_nodes = np.ascontiguousarray(bvh_nodes, np.uint8); tri_idx = np.ascontiguousarray(tri_idx, np.int32)
            d.bvh, d.tri_idx, d.bvh_len = _ptr(bvh_nodes), _ptr(tri_idx), bvh_nodes.size // 32
        h = C.c_void_p()
        _check(lib().rt_scene_from_arrays(C.byref(d), C.byref(h)))
        return cls(h.value)

    def instance_grid(self, nx, ny, nz, pitch, light_every=0):
        p = (C.c_float * 3)(*[float(v) for v in pitch])
        h = C.c_void_p()
        _check(lib().rt_scene_instance_grid(self._h, nx, ny, nz, p, light_every, C.byref(h)))
        return Scene(h.value)

    def save_rtsc(self, path):
        _check(lib().rt_scene_save_rtsc(self._h, str(path).encode()))

    def build_bvh(self, heuristic=6):
        """bvh_build (cpu/src/bvh.c:360-388); OR RT_BVH_REFBIN for the reference CPU binary's tree."""
        _check(lib().rt_scene_build_bvh(self._h, heuristic))
        return self

    def build_bvh_gpu(self, heuristic=6, device=0) -> rt_bvh_gpu_stats:
        """The same tree built on the GPU (csrc/bvh_build_gpu.cu); returns the stage timings."""
        st = rt_bvh_gpu_stats()
        _check(lib().rt_scene_build_bvh_gpu(self._h, heuristic, device, C.byref(st)))
        return st

    def flatten_host(self) -> dict:
        """The staging arrays rt_create would upload (host only, rt_debug_flatten_host)."""
        d = self.view()
        out, meta = {}, (C.c_int * 4)()
        for which, (name, dt) in enumerate([("nodes", np.float32), ("nodes4", np.float32), ("tris", np.float32), ("shade", np.float32),
                                            ("leaf_cnt", np.int32), ("mats", np.float32), ("lights", np.float32), ("nodes8", np.uint32)]):
            n = C.c_size_t()
            _check(lib().rt_debug_flatten_host(C.byref(d), which, None, 0, C.byref(n), meta))
            a = np.empty(n.value // 4, dt)
            _check(lib().rt_debug_flatten_host(C.byref(d), which, _ptr(a), a.nbytes, C.byref(n), meta))
            out[name] = a
        out["max_depth"], out["stack_need4"], out["n_lights"], out["depth8"] = meta[0], meta[1], meta[2], meta[3]
        return out

    def view(self) -> rt_scene_desc:
        d = rt_scene_desc()
        _check(lib().rt_scene_view(self._h, C.byref(d)))
        return d

    def arrays(self) -> dict:
        d = self.view()

        def arr(p, ctype, n, shape):
            if not p or n == 0:
                return np.zeros(shape, np.dtype(ctype))
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), (n,)).reshape(shape).copy()
        out = {"tri": arr(d.tri_coords, C.c_float, 9 * d.n_tris, (d.n_tris, 9)),
               "mat_idx": arr(d.tri_mat, C.c_uint32, d.n_tris, (d.n_tris,)),
               "mats": arr(d.materials, C.c_float, 9 * d.n_mats, (d.n_mats, 9)),
               "lights": arr(d.lights, C.c_float, 6 * d.n_lights, (d.n_lights, 6)),
               "ambient": np.array(list(d.ambient), np.float32)}
        if d.bvh:
            out["bvh_nodes"] = arr(d.bvh, C.c_uint8, 32 * d.bvh_len, (32 * d.bvh_len,))
            out["tri_idx"] = arr(d.tri_idx, C.c_int32, d.n_tris, (d.n_tris,))
        return out

    def close(self):
        if self._h:
            lib().rt_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------
def default_params(**kw) -> rt_render_params:
    p = rt_render_params()
    lib().rt_render_params_default(C.byref(p))
    cam = kw.pop("cam", None)
    if cam is not None:
        pos, rot, fov = cam
        p.cam.pos[:] = [float(v) for v in pos]
        p.cam.rot[:] = [float(v) for v in rot]
        p.cam.fov = float(fov)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown render parameter {k}")
        setattr(p, k, v)
    return p


class Context:
    """Device context (rt_ctx): the scene resident in HBM on one or more GPUs."""

    def __init__(self, scene: Scene, devices=None):
        d = scene.view()
        h = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            rc = lib().rt_create(C.byref(d), arr, len(devices), C.byref(h))
        else:
            rc = lib().rt_create(C.byref(d), None, 0, C.byref(h))
        _check(rc)
        self._h = h
        self.last = None

    @classmethod
    def build_on_gpu(cls, scene: "Scene", devices=None, heuristic=6, download_tree=False):
        """rt_create_gpu: BVH build + flatten on the device, no host round trip of the tree.  Returns (context, stats)."""
        h = C.c_void_p()
        st = rt_bvh_gpu_stats()
        arr = (C.c_int * len(devices))(*devices) if devices else None
        _check(lib().rt_create_gpu(scene._h, heuristic, arr, len(devices) if devices else 0, int(download_tree), C.byref(h), C.byref(st)))
        self = cls.__new__(cls)
        self._h = h
        self.last = None
        self.build_stats = st
        return self

    def render_frame(self, params: rt_render_params = None, **kw) -> rt_timing:
        """render_frame (gpu/src/gpu.cu:98-127): blocking; returns CUDA-event timings + ray counts."""
        p = params if params is not None else default_params(**kw)
        t = rt_timing()
        _check(lib().rt_render(self._h, C.byref(p), C.byref(t)), self._h)
        self.last = (p.width, p.height, p.aov_mask)
        return t

    # -- frame sequences (the reference's ITERATIONS loop): queue, copy on a second stream, wait
    def render_frame_async(self, params: rt_render_params = None, **kw) -> int:
        """rt_render_async: queue a frame on params.frame_slot; returns the slot."""
        p = params if params is not None else default_params(**kw)
        _check(lib().rt_render_async(self._h, C.byref(p)), self._h)
        self.last = (p.width, p.height, p.aov_mask)
        return p.frame_slot

    def download_async(self, slot: int, host_ptr: int):
        """rt_download_async: queue the copy of the slot's BGRA frame into (pinned) host memory."""
        _check(lib().rt_download_async(self._h, slot, C.c_void_p(host_ptr)), self._h)

    def frame_wait(self, slot: int) -> rt_timing:
        t = rt_timing()
        _check(lib().rt_frame_wait(self._h, slot, C.byref(t)), self._h)
        return t

    def load_from_gpu(self, rgb=False, tri_id=False, depth=False, out_bgra=None) -> dict:
        """load_from_gpu (gpu/src/gpu.cu:203-228): frame (and optional AOVs) to host arrays."""
        w, h, _ = self.last
        bgra = out_bgra if out_bgra is not None else np.empty((h, w, 4), np.uint8)
        out = {"bgra": bgra}
        a_rgb = np.empty((h, w, 3), np.float32) if rgb else None
        a_id = np.empty((h, w), np.int32) if tri_id else None
        a_dep = np.empty((h, w), np.float32) if depth else None
        _check(lib().rt_download(self._h, _ptr(bgra), _ptr(a_rgb), _ptr(a_id), _ptr(a_dep)), self._h)
        if rgb: out["rgb"] = a_rgb
        if tri_id: out["id"] = a_id
        if depth: out["depth"] = a_dep
        return out

    def download_into(self, host_ptr: int):
        """rt_download of the BGRA frame into caller memory (e.g. a pinned torch tensor's data_ptr)."""
        _check(lib().rt_download(self._h, C.c_void_p(host_ptr), None, None, None), self._h)

    def packed_tiles(self):
        p = C.c_void_p(); n = C.c_size_t()
        _check(lib().rt_packed_tiles(self._h, C.byref(p), C.byref(n)), self._h)
        return p.value, n.value

    def unpack_tiles(self, dev_ptr: int, stride_bytes: int, part_count: int):
        _check(lib().rt_unpack_tiles(self._h, C.c_void_p(dev_ptr), stride_bytes, part_count), self._h)

    def frame_device_ptr(self):
        p = C.c_void_p(); n = C.c_size_t()
        _check(lib().rt_frame_device_ptr(self._h, C.byref(p), C.byref(n)), self._h)
        return p.value, n.value

    def frame_ipc_export(self, width, height, slot=0) -> bytes:
        buf = C.create_string_buffer(64)
        _check(lib().rt_frame_ipc_export_slot(self._h, slot, width, height, buf), self._h)
        return buf.raw

    def frame_ipc_import(self, handle: bytes, width, height, slot=0):
        buf = C.create_string_buffer(handle, 64)
        _check(lib().rt_frame_ipc_import_slot(self._h, slot, buf, width, height), self._h)

    def device_array(self, which: int, dtype=np.uint32) -> np.ndarray:
        """Diagnostics: copy of a scene array as it lies on the context's first device (7 = nodes8)."""
        n = C.c_size_t()
        _check(lib().rt_debug_device_array(self._h, which, None, 0, C.byref(n)), self._h)
        a = np.empty(n.value // np.dtype(dtype).itemsize, dtype)
        _check(lib().rt_debug_device_array(self._h, which, _ptr(a), a.nbytes, C.byref(n)), self._h)
        return a

    def cost_map(self, width, height):
        """Diagnostics: (per-pixel traversal steps of the last fast frame, {tiles ordered ahead of the cheapest class, tiles})."""
        a = np.empty(width * height, np.uint16)
        hdr = np.zeros(2, np.uint32)
        _check(lib().rt_debug_cost_map(self._h, _ptr(a), a.size, _ptr(hdr)), self._h)
        return a.reshape(height, width), hdr

    def tile_order(self, sorted=False, cap=1 << 22):
        """Diagnostics: the first device's tile list in base order, or in the cost order the next frame will render in."""
        out = np.empty(cap, np.uint32)
        n = C.c_int(0)
        _check(lib().rt_debug_tile_order(self._h, int(bool(sorted)), _ptr(out), cap, C.byref(n)), self._h)
        return out[:min(n.value, cap)].copy()

    def warp_trace(self, enable=True, max_warps=8192):
        """Diagnostics: arm / read the per-warp timeline of RT_AOV_WORK renders (see rt_debug_warp_trace)."""
        buf = np.zeros((max_warps, 8), np.uint64)
        n = lib().rt_debug_warp_trace(self._h, int(enable), _ptr(buf), max_warps)
        return buf[:max(n, 0)]

    def set_tile_order(self, tiles):
        t = np.ascontiguousarray(tiles, np.uint32)
        _check(lib().rt_debug_set_tile_order(self._h, _ptr(t), len(t)), self._h)

    def close(self):
        if self._h:
            lib().rt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def copy_bandwidth(nbytes: int, mode: int, device: int = 0) -> float:
    """GB/s of a host->device copy of pageable memory: 0 plain cudaMemcpy, 1 staged through the pinned ring, 2 from pinned memory."""
    g = C.c_float()
    _check(lib().rt_debug_copy_bandwidth(device, nbytes, mode, C.byref(g)))
    return g.value


def write_bmp(path, bgra: np.ndarray):
    """bmp_write_file (cpu/src/bmp_writer.c:177-211) for a top-down BGRA frame."""
    bgra = np.ascontiguousarray(bgra, np.uint8)
    h, w = bgra.shape[:2]
    _check(lib().rt_write_bmp(str(path).encode(), _ptr(bgra), w, h))


def write_bmp_bottom_up(path, bgra: np.ndarray):
    """Same file from a frame rendered with RT_FRAME_BOTTOM_UP (rows already in BMP order): no flip."""
    bgra = np.ascontiguousarray(bgra, np.uint8)
    h, w = bgra.shape[:2]
    _check(lib().rt_write_bmp_bottom_up(str(path).encode(), _ptr(bgra), w, h))


class PinnedBuffer:
    """Page-locked host memory from rt_host_alloc, viewed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        _check(lib().rt_host_alloc(nbytes, C.byref(p)))
        self.ptr = p.value
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (nbytes,))

    def close(self):
        if self.ptr:
            self.array = None
            lib().rt_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_bandwidth(ws_bytes: int, device: int = 0) -> float:
    """GB/s of random 64-byte-record gathers from a working set of ws_bytes (rt_debug_gather_bandwidth)."""
    g = C.c_float()
    _check(lib().rt_debug_gather_bandwidth(device, ws_bytes, C.byref(g)))
    return g.value


def part_tile_count(width, height, part_index, part_count) -> int:
    return lib().rt_part_tile_count(width, height, part_index, part_count)
