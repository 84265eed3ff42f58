"""Heaviest-tiles-first scheduling (rt_render_params.schedule = 0, the default of the fast build on the wide trees): every
frame records its per-pixel traversal cost and the next frame of the same shape renders its tiles in the order of their most
expensive pixel.  It is scheduling only — the bytes of a pixel must not depend on when or by which warp it is rendered, whether
the cost map is fresh, stale (the camera moved) or absent (first frame, new resolution, new partition)."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
AOV = 1 | 2 | 4


@pytest.fixture(autouse=True)
def _need_gpu(rt):
    if rt.device_count() < 1:
        pytest.skip("no CUDA device")


def render(rt, ctx, w, h, **kw):
    tm = ctx.render_frame(rt.default_params(width=w, height=h, aov_mask=AOV, **kw))
    out = {k: v.copy() for k, v in ctx.load_from_gpu(rgb=True, tri_id=True, depth=True).items()}
    out["rays"] = (tm.rays_closest, tm.rays_shadow)
    out["launches"] = tm.launches
    return out


def same(a, b):
    return (np.array_equal(a["bgra"], b["bgra"]) and np.array_equal(a["id"], b["id"]) and
            np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32)) and
            np.array_equal(a["rgb"].view(np.uint32), b["rgb"].view(np.uint32)) and a["rays"] == b["rays"])


@pytest.mark.parametrize("scene", ["car_only", "car_boxed", "soup2k"])
@pytest.mark.parametrize("traversal", [2, 3, 4])
def test_heavy_first_frames_equal_chunk_order_frames(rt, gpu_scenes, scene, traversal):
    sc, _ = gpu_scenes[scene]
    ctx = rt.Context(sc, [0])          # a context of its own: the cost history starts empty
    w, h = 960, 540
    plain = render(rt, ctx, w, h, traversal=traversal, schedule=-1)
    assert plain["launches"] == 1
    with pytest.raises(rt.RtError) as no_history:   # nothing was recorded yet: there is no cost order to show
        ctx.tile_order(sorted=True)
    assert no_history.value.code == rt.RT_ERR_STATE
    for k in range(3):                 # frame 0 has no history, frames 1-2 render in cost order
        got = render(rt, ctx, w, h, traversal=traversal)
        assert got["launches"] == 3    # render kernel + the two kernels that order the next frame's tiles
        assert same(got, plain), (scene, traversal, k)
    # the order the next frame will use: a permutation of the tile list, heaviest class first, list order within a class
    base, order = ctx.tile_order(), ctx.tile_order(sorted=True)
    cost, hdr = ctx.cost_map(w, h)
    assert len(order) == len(base) == ((w + 15) // 16) * ((h + 7) // 8) == hdr[1]
    assert np.array_equal(np.sort(order), np.sort(base))
    tx = (w + 15) // 16
    pad = np.zeros((((h + 7) // 8) * 8, tx * 16), np.int64); pad[:h, :w] = cost
    tmax = pad.reshape(-1, 8, tx, 16).max(axis=(1, 3)).reshape(-1)
    v = np.maximum(tmax, 32)
    e = np.floor(np.log2(v)).astype(np.int64)                         # csrc/rt_api.cu: cost_class — half octaves from 32 steps
    cls = np.where(tmax >= 32, np.minimum(1 + 2 * (e - 5) + ((v >> (e - 1)) & 1), 20), 0)
    assert (np.diff(cls[order]) <= 0).all() and hdr[0] == (cls > 0).sum()
    pos = {int(t): i for i, t in enumerate(base)}
    for c in np.unique(cls):
        seq = [pos[int(t)] for t in order[cls[order] == c]]
        assert seq == sorted(seq), c
    # stale map: the camera moved between the frame that produced the map and the frame that uses it
    cam = (O.DEFAULT_CAM_POS, (O.DEFAULT_CAM_ROT[0], 0.0, 0.35), O.DEFAULT_FOV)
    moved = render(rt, ctx, w, h, traversal=traversal, cam=cam)
    assert same(moved, render(rt, ctx, w, h, traversal=traversal, cam=cam, schedule=-1))
    # new shape, new partition, supersampling: the history is dropped / rebuilt
    for kw in (dict(width=500, height=281), dict(width=960, height=540, part_index=1, part_count=3), dict(width=320, height=180, spp=3)):
        wh = (kw.pop("width"), kw.pop("height"))
        ref = render(rt, ctx, *wh, traversal=traversal, schedule=-1, **kw)
        for _ in range(2):
            assert same(render(rt, ctx, *wh, traversal=traversal, **kw), ref), kw
    ctx.close()


def test_pipelined_sequence_with_history(rt, gpu_scenes):
    """Two frame slots in flight: frame k + 1 is queued while frame k's cost map is still being produced (stream order)."""
    sc, _ = gpu_scenes["car_only"]
    ctx = rt.Context(sc, [0])
    w, h = 640, 360
    ctx.render_frame(rt.default_params(width=w, height=h, schedule=-1))
    want = ctx.load_from_gpu()["bgra"].copy()
    bufs = [rt.PinnedBuffer(w * h * 4) for _ in range(2)]
    for k in range(6):
        s = k % 2
        if k >= 2:
            ctx.frame_wait(s)
            assert np.array_equal(bufs[s].array.reshape(h, w, 4), want)
        ctx.render_frame_async(rt.default_params(width=w, height=h, frame_slot=s))
        ctx.download_async(s, bufs[s].ptr)
    for s in (0, 1):
        ctx.frame_wait(s)
        assert np.array_equal(bufs[s].array.reshape(h, w, 4), want)
    ctx.close()
