"""Host-side image partition arithmetic (mirror of rt_tile_owner / the pack-unpack kernels in
csrc/rt_api.cu) used by bench.py's one-process-per-GPU frame assembly and by the CPU tests.

The image is cut into RT_TILE_W x RT_TILE_H tiles; tile (tx, ty) belongs to part (tx + ty) % N
(diagonal interleave: every row and every column of tiles is spread over all parts, so the
80 %-background / 20 %-car imbalance of car_only averages out — SURVEY.md Appendix D).  A part's
"packed" buffer holds its tiles in row-major tile order, 128 BGRA pixels per tile, row-major inside.
"""
from __future__ import annotations

import numpy as np

TILE_W, TILE_H = 16, 8
TILE_PIXELS = TILE_W * TILE_H


def tiles_xy(width: int, height: int):
    return (width + TILE_W - 1) // TILE_W, (height + TILE_H - 1) // TILE_H


def tile_owner(tx, ty, parts: int):
    return (tx + ty) % parts if parts > 1 else 0 * (tx + ty)


def part_tiles(width: int, height: int, part: int, parts: int) -> np.ndarray:
    """Tile ids (ty * tiles_x + tx) owned by `part`, in the order the kernel's tile list has them: 2x2 blocks of
    tiles in raster order (make_tile_list in csrc/rt_api.cu)."""
    nx, ny = tiles_xy(width, height)
    ty, tx = np.divmod(np.arange(nx * ny), nx)
    ids = np.flatnonzero(tile_owner(tx, ty, parts) == part)
    ty, tx = ty[ids], tx[ids]
    key = ((ty // 2) * ((nx + 1) // 2) + tx // 2) * 4 + (ty % 2) * 2 + (tx % 2)
    return ids[np.argsort(key, kind="stable")].astype(np.uint32)


def pack(frame: np.ndarray, part: int, parts: int) -> np.ndarray:
    """frame (H, W, 4) u8 -> packed (n_tiles, TILE_PIXELS, 4); pixels outside the image are 0."""
    h, w = frame.shape[:2]
    nx, _ = tiles_xy(w, h)
    ids = part_tiles(w, h, part, parts)
    out = np.zeros((len(ids), TILE_H, TILE_W, 4), np.uint8)
    for i, t in enumerate(ids):
        ty, tx = divmod(int(t), nx)
        blk = frame[ty * TILE_H:(ty + 1) * TILE_H, tx * TILE_W:(tx + 1) * TILE_W]
        out[i, :blk.shape[0], :blk.shape[1]] = blk
    return out.reshape(len(ids), TILE_PIXELS, 4)


def unpack(gathered: list, width: int, height: int) -> np.ndarray:
    """Inverse of pack over all parts: gathered[p] is part p's packed buffer."""
    parts = len(gathered)
    nx, _ = tiles_xy(width, height)
    frame = np.zeros((height, width, 4), np.uint8)
    for p in range(parts):
        ids = part_tiles(width, height, p, parts)
        buf = np.asarray(gathered[p]).reshape(-1, TILE_H, TILE_W, 4)
        for i, t in enumerate(ids):
            ty, tx = divmod(int(t), nx)
            y0, x0 = ty * TILE_H, tx * TILE_W
            hh, ww = min(TILE_H, height - y0), min(TILE_W, width - x0)
            frame[y0:y0 + hh, x0:x0 + ww] = buf[i, :hh, :ww]
    return frame
