"""The one-process-per-GPU frame assembly of bench.py (fused peer stores into rank 0's frame through CUDA IPC), exercised on
ONE device: two processes on cuda:0, rank 1 imports rank 0's frame slots and stores its tiles into them.  CUDA IPC works
across processes on one device, so a 1-GPU box runs the exact code path the 2/4/8-GPU scaling runs time
(the N-part frame must be byte-identical to the 1-GPU frame)."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(autouse=True)
def _need_gpu(rt):
    if rt.device_count() < 1:
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("scene,wh,traversal", [("car_boxed", (1280, 720), 0), ("car_only", (1920, 1080), 0), ("car_only", (500, 281), 2)])
def test_two_processes_assemble_the_single_gpu_frame_through_cuda_ipc(rt, gpu_scenes, scene, wh, traversal):
    w, h = wh
    sc, ctx1 = gpu_scenes[scene]
    ctx1.render_frame(rt.default_params(width=w, height=h, traversal=traversal))
    one = ctx1.load_from_gpu()["bgra"].copy()
    rays_one = ctx1.render_frame(rt.default_params(width=w, height=h, traversal=traversal))
    rays_one = rays_one.rays_closest + rays_one.rays_shadow

    ctx0 = rt.Context(sc, [0])  # rank 0: owns the frame slots
    n_slots = rt.RT_FRAME_SLOTS
    p = subprocess.Popen([sys.executable, str(ROOT / "tests" / "ipc_worker.py"), scene, str(w), str(h), "2", "1", str(n_slots), str(traversal)],
                         stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
    try:
        for slot in range(n_slots):
            p.stdin.write(ctx0.frame_ipc_export(w, h, slot).hex() + "\n")
        p.stdin.flush()
        assert p.stdout.readline().strip() == "ready"
        for slot in range(n_slots):
            # poison the slot, then both ranks render their tiles into it
            tm0 = ctx0.render_frame(rt.default_params(width=w, height=h, part_index=0, part_count=2, frame_slot=slot, traversal=traversal))
            p.stdin.write(f"render {slot}\n"); p.stdin.flush()
            line = p.stdout.readline().split()
            assert line and line[0] == "done", line
            got = ctx0.load_from_gpu()["bgra"]
            assert np.array_equal(got, one), f"slot {slot}: assembled frame differs from the single-part frame"
            assert tm0.rays_closest + tm0.rays_shadow + int(line[1]) == rays_one  # the two parts trace exactly the frame's rays
        p.stdin.write("quit\n"); p.stdin.flush()
        assert p.wait(timeout=60) == 0
    finally:
        if p.poll() is None:
            p.kill()
        ctx0.close()
