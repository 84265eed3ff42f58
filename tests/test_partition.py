"""Multi-GPU host logic without GPUs: tile ownership, packed-tile layout, and the world_size-2
gather path over the gloo backend (the N > 1 plumbing bench.py uses with NCCL on the GPU box)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT
from parallel_ray_tracer_b200 import partition as P


@pytest.mark.parametrize("wh", [(1920, 1080), (3840, 2160), (250, 131), (16, 8), (17, 9)])
@pytest.mark.parametrize("parts", [1, 2, 3, 4, 8])
def test_every_tile_has_exactly_one_owner_and_counts_agree_with_the_library(rt, wh, parts):
    w, h = wh
    nx, ny = P.tiles_xy(w, h)
    seen = np.zeros(nx * ny, int)
    for p in range(parts):
        ids = P.part_tiles(w, h, p, parts)
        seen[ids] += 1
        assert rt.part_tile_count(w, h, p, parts) == len(ids)
    assert np.all(seen == 1)
    assert rt.part_tile_count(w, h, parts, parts) < 0  # invalid part index


@pytest.mark.parametrize("parts", [2, 4, 8])
def test_interleave_balances_work(parts, oracle_scenes):
    """static interleaved tiles keep the per-GPU load within a few percent.
    Cost proxy here: hit pixels per part on car_only (80 % background)."""
    w, h = 960, 544
    hit = (oracle_scenes["car_only"].render(w, h)["id"] >= 0).astype(np.uint8)
    frame = np.repeat(hit[:, :, None], 4, axis=2)
    loads = [int(P.pack(frame, p, parts)[:, :, 0].sum()) for p in range(parts)]
    assert max(loads) / (sum(loads) / parts) < 1.10, loads


@pytest.mark.parametrize("wh", [(250, 131), (64, 64)])
@pytest.mark.parametrize("parts", [1, 2, 3, 8])
def test_pack_unpack_roundtrip(wh, parts):
    w, h = wh
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    assert np.array_equal(P.unpack([P.pack(frame, p, parts) for p in range(parts)], w, h), frame)


def test_two_rank_gather_over_gloo(tmp_path):
    """world_size 2, gloo: each rank 'renders' (here: slices an oracle frame into) its own tiles, the
    packed buffers are all-gathered with padding to a common length, rank 0 unpacks; the assembled
    frame must be byte-identical to the single-process frame."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {str(ROOT)!r})
        import numpy as np, torch, torch.distributed as dist
        from parallel_ray_tracer_b200 import partition as P
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        w, h = 250, 131
        full = np.random.default_rng(7).integers(0, 256, (h, w, 4), dtype=np.uint8)   # same on every rank
        mine = P.pack(full, rank, world)
        n_max = max(len(P.part_tiles(w, h, p, world)) for p in range(world))
        buf = torch.zeros(n_max * P.TILE_PIXELS * 4, dtype=torch.uint8)
        buf[:mine.size] = torch.from_numpy(mine.reshape(-1))
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
        if rank == 0:
            parts = [o.numpy()[:len(P.part_tiles(w, h, p, world)) * P.TILE_PIXELS * 4] for p, o in enumerate(out)]
            frame = P.unpack(parts, w, h)
            assert np.array_equal(frame, full)
            print("GATHER_OK")
        dist.barrier()
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "GATHER_OK" in r.stdout
