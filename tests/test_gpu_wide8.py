"""The fast build on the compressed 8-wide tree (csrc/wide8.h): per-lane traversal, chunk culling and the cooperative
drain kernel.  The three are interchangeable by construction — culling only removes traversals that find nothing, the
drain kernel traces the same rays with the same arithmetic and an order-independent closest hit — so every combination
must give the SAME bytes (frame, first-hit ID, depth, float colour) and the same ray counts; and the image itself is
tied to the oracle / strict build with the north-star tolerances."""
import numpy as np
import pytest

import oracle as O
from conftest import GOLD, SCENES

pytestmark = pytest.mark.gpu
AOV = 1 | 2 | 4


@pytest.fixture(autouse=True)
def _need_gpu(rt):
    if rt.device_count() < 1:
        pytest.skip("no CUDA device")


def render(rt, ctx, w, h, **kw):
    tm = ctx.render_frame(rt.default_params(width=w, height=h, aov_mask=AOV, **kw))
    out = ctx.load_from_gpu(rgb=True, tri_id=True, depth=True)
    out = {k: v.copy() for k, v in out.items()}
    out["rays"] = (tm.rays_closest, tm.rays_shadow)
    out["launches"] = tm.launches
    return out


def same(a, b):
    return (np.array_equal(a["bgra"], b["bgra"]) and np.array_equal(a["id"], b["id"]) and
            np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32)) and
            np.array_equal(a["rgb"].view(np.uint32), b["rgb"].view(np.uint32)) and a["rays"] == b["rays"])


VARIANTS = [dict(cull=0, drain_k=0), dict(cull=1, drain_k=0), dict(cull=0, drain_k=8), dict(cull=1, drain_k=8),
            dict(cull=1, drain_k=32), dict(cull=1, drain_k=3, ctas_per_sm=2)]


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("wh", [(480, 270), (1920, 1080)])
def test_culling_and_drain_do_not_change_a_byte(rt, gpu_scenes, scene, wh):
    w, h = wh
    ctx = gpu_scenes[scene][1]
    base = render(rt, ctx, w, h, traversal=rt.RT_TRAVERSAL_WIDE8, schedule=-1, **VARIANTS[0])
    assert base["launches"] == 1
    for v in VARIANTS[1:]:
        got = render(rt, ctx, w, h, traversal=rt.RT_TRAVERSAL_WIDE8, schedule=-1, **v)
        assert same(got, base), (scene, wh, v)
        assert got["launches"] == (2 if v["drain_k"] > 0 else 1)   # render kernel (+ drain kernel)
        assert same(render(rt, ctx, w, h, traversal=rt.RT_TRAVERSAL_WIDE8, **v), base), (scene, wh, v, "heavy first")
    # against the bit-exact build: the north-star tolerances
    ctx.render_frame(rt.default_params(width=w, height=h, aov_mask=AOV, mode=rt.RT_MODE_STRICT))
    strict = ctx.load_from_gpu(rgb=True, tri_id=True, depth=True)
    m = O.compare_aovs(base, strict)
    assert m["id_match"] >= 0.9999 and m["rgb8_within1"] >= 0.999 and m["depth_within_1e-4"] >= 0.9999, m


@pytest.mark.parametrize("scene", ["car_only", "car_boxed"])
def test_wide8_vs_oracle_with_jitter_and_moved_camera(rt, gpu_scenes, oracle_scenes, manifest, scene):
    from conftest import cam_of
    pos, rot, fov = cam_of(manifest, "yaw")
    w, h, spp = 400, 225, 4
    ref = oracle_scenes[scene].render(w, h, pos=pos, rot=rot, fov=fov, spp=spp, seed=3)
    ctx = gpu_scenes[scene][1]
    got = render(rt, ctx, w, h, traversal=rt.RT_TRAVERSAL_WIDE8, cam=(pos, rot, fov), spp=spp, seed=3, cull=1, drain_k=8)
    m = O.compare_aovs(got, ref)
    assert m["id_match"] >= 0.9999 and m["rgb8_within1"] >= 0.999 and m["depth_within_1e-4"] >= 0.9999, m
    n_ref = ref["rays_closest"] + ref["rays_shadow"]
    assert abs(sum(got["rays"]) - n_ref) <= 1e-3 * n_ref
    plain = render(rt, ctx, w, h, traversal=rt.RT_TRAVERSAL_WIDE8, cam=(pos, rot, fov), spp=spp, seed=3)
    assert same(got, plain)


def test_camera_inside_and_behind_the_scene(rt, gpu_scenes):
    """Culling must stay conservative when the eye is inside the scene's box or the scene is behind / beside the camera."""
    ctx = gpu_scenes["car_boxed"][1]
    w, h = 320, 180
    for cam in [((0.0, 0.0, 1.0), (0.0, 0.0, 0.0), O.DEFAULT_FOV), ((0.0, -9.0, 3.0), (0.0, 0.0, 3.1), O.DEFAULT_FOV),
                ((0.0, -40.0, 3.0), (0.3, 0.0, 0.0), 0.4), ((30.0, -9.0, 3.0), (-0.26, 0.0, 1.2), O.DEFAULT_FOV)]:
        a = render(rt, ctx, w, h, traversal=rt.RT_TRAVERSAL_WIDE8, cam=cam)
        b = render(rt, ctx, w, h, traversal=rt.RT_TRAVERSAL_WIDE8, cam=cam, cull=1, drain_k=8)
        assert same(a, b), cam


@pytest.mark.parametrize("scene", ["car_only", "soup2k"])
def test_device_built_tree_equals_host_built_tree(rt, gpu_scenes, scene, monkeypatch):
    """rt_create_gpu builds the 8-wide tree on the device with the same level-synchronous passes: same bytes."""
    sc, ctx = gpu_scenes[scene]
    host = sc.flatten_host()["nodes8"]
    assert np.array_equal(ctx.device_array(7), host)
    sc2 = rt.Scene.load_rtsc(GOLD / "scenes" / f"{scene}.rtsc")
    ctx2 = rt.Context.build_on_gpu(sc2, [0])
    assert np.array_equal(ctx2.device_array(7), host)
    monkeypatch.setenv("RT_W8_DEVICE_BUILD", "0")       # the host-side twin of the same passes (wide8.cpp) on the device-built tree
    sc3 = rt.Scene.load_rtsc(GOLD / "scenes" / f"{scene}.rtsc")
    ctx3 = rt.Context.build_on_gpu(sc3, [0])
    assert np.array_equal(ctx3.device_array(7), host)
    ctx3.close(); sc3.close()
    a = render(rt, ctx, 480, 270, traversal=rt.RT_TRAVERSAL_WIDE8)
    b = render(rt, ctx2, 480, 270, traversal=rt.RT_TRAVERSAL_WIDE8)
    assert same(a, b)
    ctx2.close(); sc2.close()


def test_degenerate_scenes_on_the_wide_tree(rt, orc):
    """1, 2, 3 and 5 triangles; a heap of coincident triangles (depth-capped leaf with the count escape)."""
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 5):
        tri = rng.uniform(-1, 1, (n, 9)).astype(np.float32)
        tri[:, 1::3] += 2.0
        sc = rt.Scene.from_arrays(tri, np.zeros(n, np.uint32), np.array([[0.2, 0.2, 0.2, 0.7, 0.6, 0.5, 0.3, 0.3, 0.3]], np.float32),
                                  np.array([[0, -8, 3, 50, 50, 50]], np.float32)).build_bvh(6)
        ctx = rt.Context(sc, [0])
        a = render(rt, ctx, 200, 120, traversal=rt.RT_TRAVERSAL_WIDE8, cull=1, drain_k=8)
        ctx.render_frame(rt.default_params(width=200, height=120, aov_mask=AOV, mode=rt.RT_MODE_STRICT))
        s = ctx.load_from_gpu(rgb=True, tri_id=True, depth=True)
        m = O.compare_aovs(a, s)
        assert m["id_match"] >= 0.999 and m["rgb8_within1"] >= 0.999, (n, m)
        ctx.close(); sc.close()
    one = rng.uniform(-1, 1, (1, 9)).astype(np.float32); one[:, 1::3] += 2.0
    heap = np.repeat(one, 40, axis=0)
    sc = rt.Scene.from_arrays(heap, np.zeros(40, np.uint32), np.array([[0.2, 0.2, 0.2, 0.7, 0.6, 0.5, 0, 0, 0]], np.float32),
                              np.array([[0, -8, 3, 50, 50, 50]], np.float32)).build_bvh(6)
    ctx = rt.Context(sc, [0])
    a = render(rt, ctx, 200, 120, traversal=rt.RT_TRAVERSAL_WIDE8, cull=1, drain_k=8)
    b = render(rt, ctx, 200, 120, traversal=rt.RT_TRAVERSAL_WIDE8)
    ctx.render_frame(rt.default_params(width=200, height=120, aov_mask=AOV, mode=rt.RT_MODE_STRICT))
    s = ctx.load_from_gpu(rgb=True, tri_id=True, depth=True)
    assert same(a, b)
    m = O.compare_aovs(a, s)
    assert m["depth_within_1e-4"] >= 0.999 and m["rgb8_within1"] >= 0.999 and np.mean((a["id"] >= 0) == (s["id"] >= 0)) >= 0.999, m
    ctx.close(); sc.close()
