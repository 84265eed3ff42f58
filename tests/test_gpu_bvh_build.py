"""GPU BVH build (csrc/bvh_build_gpu.cu) against the host builder, which tests/test_scene_host.py
ties node for node to the reference's own bvh[] / tri_idx[] arrays.  The bar is equality of both arrays."""
import numpy as np
import pytest

from conftest import GOLD

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu(rt):
    if rt.device_count() < 1:
        pytest.skip("no CUDA device")


def nodes_of(a):
    return np.frombuffer(a["bvh_nodes"].tobytes(), dtype=[("mn", "<f4", 3), ("mx", "<f4", 3), ("tr_len", "<i4"), ("idx", "<i4")])


def assert_same_tree(host, gpu):
    h, g = nodes_of(host), nodes_of(gpu)
    assert len(h) == len(g)
    assert np.array_equal(h["tr_len"], g["tr_len"])
    assert np.array_equal(h["idx"], g["idx"])
    # boxes: equal as numbers (the GPU path canonicalises -0.0 to +0.0)
    assert np.array_equal(h["mn"], g["mn"]) and np.array_equal(h["mx"], g["mx"])
    assert np.array_equal(host["tri_idx"], gpu["tri_idx"])


def both(rt, make, flags=6):
    a = make(); a.build_bvh(flags); ha = a.arrays(); a.close()
    b = make(); st = b.build_bvh_gpu(flags); ga = b.arrays(); b.close()
    return ha, ga, st


@pytest.mark.parametrize("scene", ["car_only", "car_boxed", "soup2k"])
@pytest.mark.parametrize("flags", [6, 6 | 0x100])
def test_gpu_build_equals_host_build_on_the_shipped_scenes(rt, scene, flags):
    ha, ga, st = both(rt, lambda: rt.Scene.load_rtsc(GOLD / "scenes" / f"{scene}.rtsc"), flags)
    assert not st.fell_back
    assert st.nodes == len(nodes_of(ha))
    assert_same_tree(ha, ga)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 1023, 1024, 1025, 5000, 300000])
def test_gpu_build_equals_host_build_on_soups(rt, n):
    ha, ga, st = both(rt, lambda: rt.Scene.soup(n, 1))
    assert_same_tree(ha, ga)


def test_gpu_build_of_an_instanced_scene(rt):
    """car_only x 4 x 4 x 2 instances = 1.03 M triangles: many top levels, thousands of one-thread subtrees."""
    def make():
        base = rt.Scene.load_rtsc(GOLD / "scenes" / "car_only.rtsc")
        g = base.instance_grid(4, 4, 2, (6.0, 12.0, 4.0))
        base.close()
        return g
    ha, ga, st = both(rt, make)
    assert not st.fell_back and st.levels >= 8 and st.subtrees > 500
    assert_same_tree(ha, ga)


def test_degenerate_input_falls_back_to_the_host_builder(rt):
    """5000 copies of one triangle: no plane separates them, every split has an empty side down to depth 32."""
    tri = np.tile(np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32), (5000, 1))
    mk = lambda: rt.Scene.from_arrays(tri, np.zeros(5000, np.uint32), np.ones((1, 9), np.float32), np.zeros((0, 6), np.float32))
    ha, ga, st = both(rt, mk)
    assert_same_tree(ha, ga)


def test_rendering_with_the_gpu_built_tree(rt):
    sc = rt.Scene.load_rtsc(GOLD / "scenes" / "car_boxed.rtsc").build_bvh(6)
    ctx = rt.Context(sc, [0]); ctx.render_frame(width=320, height=180); a = ctx.load_from_gpu()["bgra"].copy(); ctx.close(); sc.close()
    sc = rt.Scene.load_rtsc(GOLD / "scenes" / "car_boxed.rtsc"); sc.build_bvh_gpu(6)
    ctx = rt.Context(sc, [0]); ctx.render_frame(width=320, height=180); b = ctx.load_from_gpu()["bgra"]; ctx.close(); sc.close()
    assert np.array_equal(a, b)


# ---- rt_create_gpu: build + flatten on the device, the tree never visits the host (csrc/flatten_gpu.cu) ----
def render_all(rt, ctx, w, h):
    out = {}
    tm = ctx.render_frame(rt.default_params(width=w, height=h, mode=rt.RT_MODE_STRICT, aov_mask=1 | 2 | 4 | 8))
    a = ctx.load_from_gpu(rgb=True, tri_id=True, depth=True)
    out["strict"] = (a["bgra"].copy(), a["rgb"].copy(), a["id"].copy(), a["depth"].copy(),
                     (tm.rays_closest, tm.rays_shadow, tm.inner_visits, tm.tri_tests))
    for trav in (2, 3):  # 2-wide speculative, 4-wide: both layouts of the device-side flatten
        tm = ctx.render_frame(rt.default_params(width=w, height=h, traversal=trav, aov_mask=2 | 4))
        a = ctx.load_from_gpu(tri_id=True, depth=True)
        out[trav] = (a["bgra"].copy(), a["id"].copy(), a["depth"].copy(), (tm.rays_closest, tm.rays_shadow))
    return out


def assert_same_renders(x, y):
    for k in x:
        for a, b in zip(x[k], y[k]):
            if isinstance(a, tuple):
                assert a == b, k
            else:
                assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), k


@pytest.mark.parametrize("scene", ["car_only", "car_boxed", "soup2k"])
def test_device_side_flatten_renders_exactly_like_the_host_path(rt, scene):
    """Strict frames (bit-exact vs the oracle in test_gpu_parity.py) including the visit / test counters, and both fast
    traversal layouts, must not change when build and flatten run on the device."""
    path = GOLD / "scenes" / f"{scene}.rtsc"
    sc = rt.Scene.load_rtsc(path).build_bvh(6)
    ctx = rt.Context(sc, [0]); want = render_all(rt, ctx, 480, 270); ctx.close(); sc.close()
    sc = rt.Scene.load_rtsc(path)
    ctx = rt.Context.build_on_gpu(sc, [0])
    assert not ctx.build_stats.fell_back and ctx.build_stats.nodes > 0
    assert sc.view().bvh_len == 0  # the tree stayed on the device
    got = render_all(rt, ctx, 480, 270); ctx.close(); sc.close()
    assert_same_renders(want, got)


def test_device_side_flatten_on_an_instanced_scene_and_with_download(rt):
    def make():
        base = rt.Scene.load_rtsc(GOLD / "scenes" / "car_only.rtsc")
        g = base.instance_grid(3, 3, 1, (11.5, 6.5, 3.0))
        base.close()
        return g
    sc = make().build_bvh(6)
    ctx = rt.Context(sc, [0]); want = render_all(rt, ctx, 640, 360); ctx.close(); sc.close()
    sc = make(); ctx = rt.Context.build_on_gpu(sc, [0]); got = render_all(rt, ctx, 640, 360); ctx.close(); sc.close()
    assert_same_renders(want, got)
    sc = make(); ctx = rt.Context.build_on_gpu(sc, [0], download_tree=True)
    assert sc.view().bvh_len == ctx.build_stats.nodes
    got = render_all(rt, ctx, 640, 360); ctx.close(); sc.close()
    assert_same_renders(want, got)


@pytest.mark.parametrize("n", [1, 2, 3, 40])
def test_device_side_path_on_tiny_and_degenerate_scenes(rt, n):
    sc = rt.Scene.soup(n, 1).build_bvh(6)
    ctx = rt.Context(sc, [0]); want = render_all(rt, ctx, 160, 90); ctx.close(); sc.close()
    sc = rt.Scene.soup(n, 1); ctx = rt.Context.build_on_gpu(sc, [0]); got = render_all(rt, ctx, 160, 90); ctx.close(); sc.close()
    assert_same_renders(want, got)
    tri = np.tile(np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32), (600, 1))  # falls back to the host builder
    mk = lambda: rt.Scene.from_arrays(tri, np.zeros(600, np.uint32), np.ones((1, 9), np.float32), np.zeros((0, 6), np.float32))
    a = mk().build_bvh(6); ctx = rt.Context(a, [0]); want = render_all(rt, ctx, 96, 54); ctx.close()
    b = mk(); ctx = rt.Context.build_on_gpu(b, [0]); got = render_all(rt, ctx, 96, 54); ctx.close()
    assert_same_renders(want, got)


def test_device_side_path_on_two_devices(rt):
    if rt.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    path = GOLD / "scenes" / "car_boxed.rtsc"
    sc = rt.Scene.load_rtsc(path).build_bvh(6)
    ctx = rt.Context(sc, [0]); ctx.render_frame(width=800, height=450); want = ctx.load_from_gpu()["bgra"].copy(); ctx.close(); sc.close()
    sc = rt.Scene.load_rtsc(path); ctx = rt.Context.build_on_gpu(sc, [0, 1])
    tm = ctx.render_frame(width=800, height=450)
    assert tm.n_devices == 2 and np.array_equal(ctx.load_from_gpu()["bgra"], want)
    ctx.close(); sc.close()
