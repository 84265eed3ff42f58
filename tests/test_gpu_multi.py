"""Several GPUs in one context (NVLink peer stores vs staged peer copies).  Needs >= 2 devices."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu(rt):
    if rt.device_count() < 1:
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("gather", ["peer_store", "peer_copy"])
def test_n_gpu_frame_is_byte_identical_to_one_gpu(rt, gpu_scenes, gather):
    n = rt.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    sc, ctx1 = gpu_scenes["car_boxed"]
    w, h = 1280, 720
    ctx1.render_frame(rt.default_params(width=w, height=h))
    one = ctx1.load_from_gpu()["bgra"]
    for nd in sorted({2, n}):
        ctxn = rt.Context(sc, list(range(nd)))
        g = rt.RT_GATHER_PEER_STORE if gather == "peer_store" else rt.RT_GATHER_PEER_COPY
        tm = ctxn.render_frame(rt.default_params(width=w, height=h, gather=g))
        assert tm.n_devices == nd
        assert np.array_equal(ctxn.load_from_gpu()["bgra"], one)
        ctxn.close()
