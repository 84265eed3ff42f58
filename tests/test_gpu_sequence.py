"""Frame sequences through the C-ABI: rt_render_async / rt_download_async / rt_frame_wait (the reference's
ITERATIONS loop, cpu/src/main.c:169-185, with the frame copy overlapped) and RT_FRAME_BOTTOM_UP (the BMP row
order of cpu/src/bmp_writer.c:122-146 produced by the kernel's own writeback).  The bar is byte equality with
the blocking path, which tests/test_gpu_parity.py ties to the oracle."""
import numpy as np
import pytest

from conftest import GOLD

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu(rt):
    if rt.device_count() < 1:
        pytest.skip("no CUDA device")


def cams(n):
    """Moving camera: the reference's commented-out cam.rot.z sweep (cpu/src/main.c:107)."""
    import oracle as O
    out = []
    for k in range(n):
        rot = (O.DEFAULT_CAM_ROT[0], O.DEFAULT_CAM_ROT[1], 0.05 * k)
        out.append((O.DEFAULT_CAM_POS, rot, O.DEFAULT_FOV))
    return out


@pytest.mark.parametrize("scene", ["car_only", "car_boxed"])
def test_pipelined_sequence_equals_blocking_frames(rt, gpu_scenes, scene):
    _, ctx = gpu_scenes[scene]
    w, h, n = 640, 360, 7
    want = []
    for cam in cams(n):
        ctx.render_frame(rt.default_params(width=w, height=h, cam=cam))
        want.append(ctx.load_from_gpu()["bgra"].copy())
    assert not np.array_equal(want[0], want[1])  # the camera really moves
    bufs = [rt.PinnedBuffer(w * h * 4) for _ in range(rt.RT_FRAME_SLOTS)]
    got = []
    for k, cam in enumerate(cams(n)):
        s = k % rt.RT_FRAME_SLOTS
        if k >= rt.RT_FRAME_SLOTS:
            tm = ctx.frame_wait(s)
            assert tm.rays_closest >= w * h and tm.kernel_ms[0] > 0
            got.append(bufs[s].array.reshape(h, w, 4).copy())
        assert ctx.render_frame_async(rt.default_params(width=w, height=h, cam=cam, frame_slot=s)) == s
        ctx.download_async(s, bufs[s].ptr)
    for k in range(n - rt.RT_FRAME_SLOTS, n):
        s = k % rt.RT_FRAME_SLOTS
        ctx.frame_wait(s)
        got.append(bufs[s].array.reshape(h, w, 4).copy())
    assert len(got) == n
    for k in range(n):
        assert np.array_equal(got[k], want[k]), f"frame {k}"
    for b in bufs:
        b.close()


def test_slot_must_be_waited_before_reuse(rt, gpu_scenes):
    _, ctx = gpu_scenes["soup2k"]
    p = rt.default_params(width=320, height=180, frame_slot=1)
    ctx.render_frame_async(p)
    with pytest.raises(rt.RtError) as e:
        ctx.render_frame_async(p)
    assert e.value.code == rt.RT_ERR_STATE
    ctx.frame_wait(1)
    ctx.render_frame_async(p)
    ctx.frame_wait(1)
    with pytest.raises(rt.RtError):
        ctx.render_frame_async(rt.default_params(width=320, height=180, frame_slot=rt.RT_FRAME_SLOTS))


@pytest.mark.parametrize("mode", ["fast", "strict"])
def test_bottom_up_frame_is_the_flipped_frame(rt, gpu_scenes, mode, tmp_path):
    _, ctx = gpu_scenes["car_only"]
    m = rt.RT_MODE_FAST if mode == "fast" else rt.RT_MODE_STRICT
    w, h = 333, 187  # ragged: partial tiles on both edges
    ctx.render_frame(rt.default_params(width=w, height=h, mode=m, aov_mask=2))
    top = ctx.load_from_gpu(tri_id=True)
    ctx.render_frame(rt.default_params(width=w, height=h, mode=m, aov_mask=2, frame_flags=rt.RT_FRAME_BOTTOM_UP))
    bot = ctx.load_from_gpu(tri_id=True)
    assert np.array_equal(bot["bgra"], top["bgra"][::-1])
    assert np.array_equal(bot["id"], top["id"])  # AOVs stay top-down
    # the downloaded bottom-up buffer IS the BMP pixel array
    rt.write_bmp(tmp_path / "a.bmp", top["bgra"])
    rt.write_bmp_bottom_up(tmp_path / "b.bmp", bot["bgra"])
    assert (tmp_path / "a.bmp").read_bytes() == (tmp_path / "b.bmp").read_bytes()


def test_bottom_up_bmp_matches_the_reference_bmp(rt, gpu_scenes, manifest, tmp_path):
    """tests/golden/ref_car_only_64x36.bmp was written by the reference's own bmp_write_file."""
    gold = (GOLD / "ref_car_only_64x36.bmp").read_bytes()
    _, ctx = gpu_scenes["car_only"]
    ctx.render_frame(rt.default_params(width=64, height=36, mode=rt.RT_MODE_STRICT, frame_flags=rt.RT_FRAME_BOTTOM_UP))
    rt.write_bmp_bottom_up(tmp_path / "c.bmp", ctx.load_from_gpu()["bgra"])
    mine = (tmp_path / "c.bmp").read_bytes()
    assert mine[:54] == gold[:54]
    a = np.frombuffer(mine[54:], np.uint8).astype(int)
    b = np.frombuffer(gold[54:], np.uint8).astype(int)
    # the reference binary is built with -ffast-math: 1 LSB on a handful of pixels (tests/test_oracle_vs_reference.py)
    assert (np.abs(a - b) <= 1).mean() >= 0.999


def test_multi_device_sequence(rt, gpu_scenes):
    n = rt.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    sc, ctx1 = gpu_scenes["car_boxed"]
    w, h = 800, 450
    ctxn = rt.Context(sc, [0, 1])
    bufs = [rt.PinnedBuffer(w * h * 4) for _ in range(2)]
    cs = cams(4)
    for k, cam in enumerate(cs):
        s = k % 2
        if k >= 2:
            ctxn.frame_wait(s)
        ctxn.render_frame_async(rt.default_params(width=w, height=h, cam=cam, frame_slot=s))
        ctxn.download_async(s, bufs[s].ptr)
    for s in (0, 1):
        ctxn.frame_wait(s)
    for k in (2, 3):
        ctx1.render_frame(rt.default_params(width=w, height=h, cam=cs[k]))
        assert np.array_equal(ctx1.load_from_gpu()["bgra"].reshape(-1), bufs[k % 2].array)
    ctxn.close()


def test_sequence_error_paths(rt, gpu_scenes):
    _, ctx = gpu_scenes["soup2k"]
    fresh = rt.Context(gpu_scenes["soup2k"][0], [0])
    buf = rt.PinnedBuffer(64 * 36 * 4)
    with pytest.raises(rt.RtError) as e:
        fresh.download_async(0, buf.ptr)          # nothing rendered on the slot
    assert e.value.code == rt.RT_ERR_STATE
    with pytest.raises(rt.RtError) as e:
        fresh.frame_wait(1)
    assert e.value.code == rt.RT_ERR_STATE
    with pytest.raises(rt.RtError) as e:
        fresh.frame_wait(5)
    assert e.value.code == rt.RT_ERR_INVALID
    fresh.render_frame_async(rt.default_params(width=64, height=36, frame_slot=0))
    fresh.download_async(0, buf.ptr)
    t1 = fresh.frame_wait(0)
    t2 = fresh.frame_wait(0)                      # waiting twice returns the same timing
    assert (t1.rays_closest, t1.kernel_ms[0]) == (t2.rays_closest, t2.kernel_ms[0]) and t1.rays_closest == 64 * 36
    ctx.render_frame(rt.default_params(width=64, height=36))
    assert np.array_equal(buf.array.reshape(36, 64, 4), ctx.load_from_gpu()["bgra"])
    # bottom-up rows are refused where tiles are packed
    fresh.render_frame(rt.default_params(width=64, height=36, frame_flags=rt.RT_FRAME_BOTTOM_UP, part_index=0, part_count=2))
    with pytest.raises(rt.RtError) as e:
        fresh.packed_tiles()
    assert e.value.code == rt.RT_ERR_STATE
    fresh.close(); buf.close()


def test_aovs_belong_to_their_frame_slot(rt, gpu_scenes):
    """AOV buffers are per frame slot — a render queued on slot 1 must not overwrite what rt_download returns for
    slot 0 (they used to be shared by the context)."""
    _, ctx = gpu_scenes["car_boxed"]
    w, h = 320, 180
    aov = rt.RT_AOV_RGB_F32 | rt.RT_AOV_TRI_ID | rt.RT_AOV_DEPTH
    cs = cams(2)
    want = []
    for cam in cs:
        ctx.render_frame(rt.default_params(width=w, height=h, cam=cam, aov_mask=aov))
        want.append({k: v.copy() for k, v in ctx.load_from_gpu(rgb=True, tri_id=True, depth=True).items()})
    assert not np.array_equal(want[0]["id"], want[1]["id"])
    for s in (0, 1):
        ctx.render_frame_async(rt.default_params(width=w, height=h, cam=cs[s], aov_mask=aov, frame_slot=s))
    ctx.frame_wait(1)
    ctx.frame_wait(0)   # slot 0 is now "the last finished render": rt_download must return ITS AOVs
    got = ctx.load_from_gpu(rgb=True, tri_id=True, depth=True)
    for k in ("bgra", "id", "depth", "rgb"):
        assert np.array_equal(got[k], want[0][k]), k
