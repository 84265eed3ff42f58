"""CPU-tier checks of the compressed 8-wide tree (csrc/wide8.h, built by csrc/wide8.cpp, read back through the host-only
rt_debug_flatten_host): the layout rules the kernels rely on, without a GPU.

  * every triangle slot is referenced by exactly one leaf, every node by exactly one parent (breadth-first numbering);
  * every decoded child box encloses, with the promised 1/16-step margin, all triangles below it (outward rounding —
    the reference GPU program's FP16 boxes, gpu/src/gpu.cu:176-185, fail exactly this);
  * a plain numpy restatement of the kernels' decode-and-test arithmetic (float32, same operation order) walking the
    tree in octant order finds the oracle's first hit for every probe ray;
  * the bytes do not depend on the number of host threads.
"""
import numpy as np
import pytest

import oracle as O
from conftest import GOLD

NONE = -2 ** 31


def decode_node(w):
    """-> p[3], s[3] (grid step per axis, float64), qlo[3][8], qhi[3][8], refs[8], n_children"""
    p = w[0:3].view(np.float32).astype(np.float64)
    e = np.array([(int(w[3]) >> (8 * a)) & 0xff for a in range(3)]) - 127 + 7
    s = np.ldexp(1.0, e)
    b = w[4:16].view(np.uint8).reshape(6, 8)
    return p, s, b[0:3].astype(np.int64), b[3:6].astype(np.int64), w[16:24].view(np.int32), int(w[3]) >> 24


def leaf_range(ref, leaf_cnt):
    v = ~int(ref)
    first, cnt = v >> 4, v & 15
    if cnt == 15:
        cnt = int(leaf_cnt[first])
    return first, cnt


@pytest.fixture(scope="module")
def flats(rt):
    out = {}
    for name in ("soup2k", "car_only", "car_boxed"):
        sc = rt.Scene.load_rtsc(GOLD / "scenes" / f"{name}.rtsc").build_bvh(6)
        out[name] = (sc, sc.flatten_host(), sc.arrays())
    yield out
    for sc, _, _ in out.values():
        sc.close()


@pytest.mark.parametrize("scene", ["soup2k", "car_only", "car_boxed"])
def test_structure_and_enclosure(flats, scene):
    sc, f, a = flats[scene]
    n8 = f["nodes8"].reshape(-1, 24)
    n_tris = a["tri"].shape[0]
    tri = a["tri"].reshape(n_tris, 3, 3)[a["tri_idx"]].astype(np.float64)   # leaf-order slots
    slot_mn, slot_mx = tri.min(axis=1), tri.max(axis=1)
    covered = np.zeros(n_tris, np.int32)
    parents = np.zeros(len(n8), np.int32)
    sub_mn = np.full((len(n8), 3), np.inf)
    sub_mx = np.full((len(n8), 3), -np.inf)
    worst = 1.0
    for k in range(len(n8) - 1, -1, -1):                                     # children have larger indices (breadth-first)
        p, s, qlo, qhi, refs, nch = decode_node(n8[k])
        assert (refs != NONE).sum() == nch
        for slot in range(8):
            r = int(refs[slot])
            if r == NONE:
                assert np.all(qlo[:, slot] == 255) and np.all(qhi[:, slot] == 0)  # inverted: can never be hit
                continue
            if r >= 0:
                assert r > k
                parents[r] += 1
                mn, mx = sub_mn[r], sub_mx[r]
            else:
                first, cnt = leaf_range(r, f["leaf_cnt"])
                assert cnt >= 1 and first + cnt <= n_tris
                covered[first:first + cnt] += 1
                mn, mx = slot_mn[first:first + cnt].min(0), slot_mx[first:first + cnt].max(0)
            lo = p + s * qlo[:, slot]
            hi = p + s * qhi[:, slot]
            # outward rounding with margin: at least 1/16 step outside the true box on every side
            m = np.minimum((mn - lo) / s, (hi - mx) / s).min()
            worst = min(worst, m)
            sub_mn[k] = np.minimum(sub_mn[k], mn)
            sub_mx[k] = np.maximum(sub_mx[k], mx)
    assert worst >= 0.0625 - 1e-9, worst
    assert np.all(covered == 1)
    assert parents[0] == 0 and np.all(parents[1:] == 1)
    assert f["depth8"] + 2 <= 40


def kernel_box_hits(w, o, d, tmax):
    """The kernels' test of the eight children of one node (render_kernel.cuh: wide8_visit), float32 op for op:
    returns the hit mask in SLOT order."""
    f32 = np.float32
    d = np.where(np.abs(d) < f32(1e-30), np.copysign(f32(1e-30), d), d).astype(f32)   # ray_begin: zero components for the box test
    idv = (f32(1.0) / d).astype(f32)
    ob = (-o * idv).astype(f32)
    p = w[0:3].view(np.float32)
    b = w[4:16].view(np.uint8).reshape(6, 8)
    hits = 0
    a = np.zeros(3, f32); bb = np.zeros(3, f32)
    for ax in range(3):
        E = (int(w[3]) >> (8 * ax)) & 0xff
        S = np.array([E << 23], np.uint32).view(np.float32)[0]             # s / 128
        P0 = f32(np.float64(f32(-8388608.0)) * np.float64(S) + np.float64(p[ax]))  # fma: one rounding
        a[ax] = f32(S * idv[ax])
        bb[ax] = f32(np.float64(P0) * np.float64(idv[ax]) + np.float64(ob[ax]))
    for slot in range(8):
        tn, tf = f32(0.0), f32(tmax)
        for ax in range(3):
            neg = d[ax] < 0
            qn = int(b[3 + ax, slot] if neg else b[ax, slot])
            qf = int(b[ax, slot] if neg else b[3 + ax, slot])
            vn = np.array([0x4B000000 + 128 * qn], np.uint32).view(np.float32)[0]
            vf = np.array([0x4B000000 + 128 * qf], np.uint32).view(np.float32)[0]
            t_n = f32(np.float64(vn) * np.float64(a[ax]) + np.float64(bb[ax]))
            t_f = f32(np.float64(vf) * np.float64(a[ax]) + np.float64(bb[ax]))
            if not np.isnan(t_n): tn = max(tn, t_n)
            if not np.isnan(t_f): tf = min(tf, t_f)
        if tn <= tf:
            hits |= 1 << slot
    return hits


def walk(n8, leaf_cnt, tris, o, d, orc):
    """Closest hit through the 8-wide tree in octant order; triangles tested with the oracle's hit_triangle."""
    ds = np.where(np.abs(d) < np.float32(1e-30), np.copysign(np.float32(1e-30), d), d)
    octant = (1 if ds[0] < 0 else 0) | (2 if ds[1] < 0 else 0) | (4 if ds[2] < 0 else 0)
    best_t, best = np.float32(3.4028234663852886e38), -1
    stack = [0]
    visited = 0
    while stack:
        k = stack.pop()
        visited += 1
        with np.errstate(all="ignore"):   # axis-parallel rays: inf - inf = NaN, dropped by the min/max as on the device
            m = kernel_box_hits(n8[k], o, d, best_t)
        refs = n8[k][16:24].view(np.int32)
        order = [key ^ octant for key in range(8) if (m >> (key ^ octant)) & 1]
        for slot in reversed(order):                                         # far first onto the stack
            r = int(refs[slot])
            assert r != NONE
            if r >= 0:
                stack.append(r)
            else:
                first, cnt = leaf_range(r, leaf_cnt)
                for j in range(first, first + cnt):
                    t, _ = orc.hit_triangle(o, d, tris[j])
                    if t < best_t:
                        best_t, best = t, j
    return best, best_t, visited


@pytest.mark.parametrize("scene", ["soup2k", "car_only"])
def test_kernel_arithmetic_finds_the_oracle_first_hit(flats, scene, orc):
    sc, f, a = flats[scene]
    n8 = f["nodes8"].reshape(-1, 24)
    n_tris = a["tri"].shape[0]
    tris_slot = a["tri"].reshape(n_tris, 9)[a["tri_idx"]]
    osc = orc.scene(O.load_rtsc(GOLD / "scenes" / f"{scene}.rtsc"))
    osc.set_bvh(a["bvh_nodes"], a["tri_idx"])
    rng = np.random.default_rng(7)
    n_hit = 0
    for i in range(160):
        if i % 2 == 0:   # camera-like rays
            o = np.array(O.DEFAULT_CAM_POS, np.float32)
            tgt = rng.uniform(-2.5, 2.5, 3).astype(np.float32)
            d = (tgt - o).astype(np.float32)
        else:            # rays from inside the scene, incl. axis-parallel ones
            o = rng.uniform(-3, 3, 3).astype(np.float32)
            d = rng.normal(size=3).astype(np.float32)
            if i % 16 == 1:
                d[rng.integers(3)] = 0.0
        if not np.any(d):
            continue
        ref_id, ref_t, _ = osc.trace_closest(o, d)
        slot, t, _ = walk(n8, f["leaf_cnt"], tris_slot, o, d, orc)
        got_id = int(a["tri_idx"][slot]) if slot >= 0 else -1
        assert t == ref_t, (i, o, d, got_id, ref_id)
        if ref_id >= 0:
            n_hit += 1
            # equal t can come from two triangles sharing an edge; otherwise the triangle is the oracle's
            assert got_id == ref_id or t == ref_t
    assert n_hit > 30


def test_bytes_do_not_depend_on_the_thread_count(rt, flats, monkeypatch):
    import subprocess, sys, hashlib
    from pathlib import Path
    code = ("import sys, hashlib; sys.path.insert(0, %r); import parallel_ray_tracer_b200 as rt;"
            "sc = rt.Scene.load_rtsc(%r).build_bvh(6); print(hashlib.sha256(sc.flatten_host()['nodes8'].tobytes()).hexdigest())"
            % (str(Path(__file__).resolve().parent.parent), str(GOLD / "scenes" / "car_boxed.rtsc")))
    import os
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env={**os.environ, "RT_FLATTEN_THREADS": "1"})
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == hashlib.sha256(flats["car_boxed"][1]["nodes8"].tobytes()).hexdigest()
