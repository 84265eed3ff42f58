"""The HBM layout (csrc/device_layout.h) as produced by the host flatten (csrc/flatten.cpp), checked on the CPU through
rt_debug_flatten_host: structural rules, the 4-wide collapse against the 2-wide records, independence from the thread count,
and — by walking the flattened records with numpy — the same first hits as the oracle's bvh_traverse."""
import numpy as np
import pytest

from conftest import GOLD, SCENES

REF_NONE = -(1 << 31)


def flat_of(rt, scene, heuristic=6):
    sc = rt.Scene.load_rtsc(GOLD / "scenes" / f"{scene}.rtsc").build_bvh(heuristic)
    f, a = sc.flatten_host(), sc.arrays()
    sc.close()
    return f, a


def refs2(f):
    return f["nodes"].reshape(-1, 16)[:, 12:14].view(np.int32)


def leaf_range(f, ref):
    v = ~int(ref)
    first, cnt = v >> 4, v & 15
    if cnt == 15:
        cnt = int(f["leaf_cnt"][first])
    return first, cnt


@pytest.mark.parametrize("scene", SCENES)
def test_layout_rules(rt, scene):
    f, a = flat_of(rt, scene)
    n = len(a["tri"])
    nodes = f["nodes"].reshape(-1, 16)
    r2 = refs2(f)
    n_inner = len(nodes)
    # every triangle slot is covered by exactly one leaf reference, every inner record is referenced exactly once (record 0 = root)
    seen_tri = np.zeros(n, int)
    seen_node = np.zeros(n_inner, int)
    for ref in r2.reshape(-1):
        if ref == REF_NONE:
            continue
        if ref >= 0:
            seen_node[ref] += 1
        else:
            first, cnt = leaf_range(f, ref)
            seen_tri[first:first + cnt] += 1
    assert (seen_tri == 1).all() and (seen_node[1:] == 1).all() and seen_node[0] == 0
    # triangle records: v0, e1, e2, n = e1 x e2 of the triangle tri_idx[j], original index in the fourth quad
    t = f["tris"].reshape(n, 16)
    orig = t[:, 12].view(np.int32)
    assert np.array_equal(orig, a["tri_idx"])
    c = a["tri"][orig]
    assert np.array_equal(t[:, 0:3], c[:, 0:3])
    assert np.array_equal(t[:, 3:6], c[:, 3:6] - c[:, 0:3]) and np.array_equal(t[:, 6:9], c[:, 6:9] - c[:, 0:3])
    e1, e2 = t[:, 3:6].astype(np.float32), t[:, 6:9].astype(np.float32)
    nn = np.stack([e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1], e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2], e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], 1)
    assert np.array_equal(t[:, 9:12], nn.astype(np.float32))
    assert (t[:, 13:16] == 0).all()
    # child boxes contain the triangles below them (leaves) and the boxes below them (inner children)
    for k in range(0, n_inner, max(1, n_inner // 300)):
        for w in (0, 1):
            ref = r2[k, w]
            mn = nodes[k, [0, 1, 2]] if w == 0 else nodes[k, [6, 7, 8]]
            mx = nodes[k, [3, 4, 5]] if w == 0 else nodes[k, [9, 10, 11]]
            if ref == REF_NONE:
                assert np.isinf(mn).all() and np.isinf(mx).all()
            elif ref >= 0:
                sub = nodes[ref]
                lo = np.minimum(sub[[0, 1, 2]], sub[[6, 7, 8]]); hi = np.maximum(sub[[3, 4, 5]], sub[[9, 10, 11]])
                fin = np.isfinite(lo) & np.isfinite(hi)
                assert (lo[fin] >= mn[fin]).all() and (hi[fin] <= mx[fin]).all()
            else:
                first, cnt = leaf_range(f, ref)
                v = a["tri"][a["tri_idx"][first:first + cnt]].reshape(-1, 3)
                assert (v.min(0) >= mn).all() and (v.max(0) <= mx).all()
    assert f["max_depth"] <= 32 and f["stack_need4"] >= 3


def expand4(bvh, root, leaf_max, width=4):
    """csrc/wide8.h: w8_expand restated on the reference node array: the frontier of reference inner node `root` after expanding
    the inner child of largest surface area until `width` children exist; a subtree of <= leaf_max triangles is one leaf.
    Returns [(reference node, is_inner, first_slot, count)] left to right."""
    def inner(b): return bvh["tr_len"][b] == 0 and bvh["idx"][b] != 0
    def empty(b): return bvh["tr_len"][b] == 0 and bvh["idx"][b] == 0
    def small(b):
        stack, total, first = [b], 0, -1
        while stack:
            x = stack.pop()
            if inner(x):
                stack += [int(bvh["idx"][x]) + 1, int(bvh["idx"][x])]
            elif bvh["tr_len"][x] > 0:
                if first < 0: first = int(bvh["idx"][x])
                total += int(bvh["tr_len"][x])
                if total > leaf_max: return None
        return (max(first, 0), total)
    def child(b):
        if inner(b):
            s = small(b)
            return (b, True, 0, 0) if s is None else (b, False, s[0], s[1])
        return (b, False, int(bvh["idx"][b]), int(bvh["tr_len"][b]))
    def area(b):
        d = bvh["max"][b].astype(np.float64) - bvh["min"][b].astype(np.float64)
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0]
    l = int(bvh["idx"][root])
    out = [child(b) for b in (l, l + 1) if not empty(b)]
    while len(out) < width:
        cand = [(area(c[0]), -i) for i, c in enumerate(out) if c[1]]
        if not cand: break
        best = -max(cand)[1]                      # largest area, first of equals
        c = int(bvh["idx"][out[best][0]])
        kids = [child(b) for b in (c, c + 1) if not empty(b)]
        out[best:best + 1] = kids
    return out


@pytest.mark.parametrize("scene", SCENES)
def test_four_wide_tree_is_the_area_ordered_expansion_of_the_reference_tree(rt, scene):
    f, a = flat_of(rt, scene)
    bvh = a["bvh_nodes"].view(np.dtype([("min", "<f4", 3), ("max", "<f4", 3), ("tr_len", "<i4"), ("idx", "<i4")]))
    n4 = f["nodes4"].reshape(-1, 32)
    r4 = n4[:, 24:28].view(np.int32)
    if not (bvh["tr_len"][0] == 0 and bvh["idx"][0] != 0):
        pytest.skip("the root is a leaf: one synthetic node")
    leaf_max = 3                                  # csrc/wide8.cpp: wide4_leaf_max default
    # breadth-first, the inner children of a level numbered in order: walk both trees level by level
    level, base, seen_tris = [0], 0, np.zeros(len(a["tri"]), int)
    while level:
        nxt = []
        next_base = base + len(level)
        for i, root in enumerate(level):
            k4 = base + i
            kids = expand4(bvh, root, leaf_max)
            assert 1 <= len(kids) <= 4
            for j, (b, is_in, first, cnt) in enumerate(kids):
                # centre / half extent (csrc/flatten.h: box_center_half): contains the reference box, at most 2 ulp wider
                c, h = n4[k4, [0 + j, 4 + j, 8 + j]].astype(np.float64), n4[k4, [12 + j, 16 + j, 20 + j]].astype(np.float64)
                lo, hi = bvh["min"][b].astype(np.float64), bvh["max"][b].astype(np.float64)
                assert (c - h <= lo).all() and (c + h >= hi).all()
                slack = 4 * np.spacing(np.maximum(np.maximum(np.abs(lo), np.abs(hi)), 1e-30).astype(np.float32)).astype(np.float64)
                assert (lo - (c - h) <= slack).all() and ((c + h) - hi <= slack).all()
                assert ((h == 0) == (lo == hi)).all()
                if is_in:
                    assert r4[k4, j] == next_base + len(nxt)
                    nxt.append(b)
                else:
                    assert cnt >= 1 and r4[k4, j] == ~((first << 4) | min(cnt, 15))
                    seen_tris[first:first + cnt] += 1
            for j in range(len(kids), 4):
                assert r4[k4, j] == REF_NONE and np.isinf(n4[k4, j]) and n4[k4, 12 + j] == 0   # empty slot: centre +inf, half 0
        base, level = next_base, nxt
    assert base == len(n4) and (seen_tris == 1).all()
    # the expansion fills the nodes: at most a quarter of the slots stay empty (the collapse of every other level left a third)
    assert (r4 != REF_NONE).mean() > 0.75


def test_flatten_does_not_depend_on_the_thread_count(rt, monkeypatch):
    sc = rt.Scene.load_rtsc(GOLD / "scenes" / "car_boxed.rtsc")
    big = sc.instance_grid(2, 2, 1, (11.5, 6.5, 3.0))   # 184 K triangles: enough for several chunks
    sc.close()
    big.build_bvh(6)
    # (the thread count is read once per process: compare a child process pinned to one thread with this one)
    import hashlib, subprocess, sys, textwrap
    def digest(f):
        h = hashlib.sha256()
        for k in ("nodes", "nodes4", "tris", "shade", "leaf_cnt", "mats", "lights"):
            h.update(f[k].tobytes())
        return h.hexdigest() + f":{f['max_depth']}:{f['stack_need4']}"
    mine = digest(big.flatten_host())
    big.close()
    code = textwrap.dedent(f"""
        import sys, hashlib; sys.path.insert(0, {str(GOLD.parent.parent)!r})
        import parallel_ray_tracer_b200 as rt
        sc = rt.Scene.load_rtsc({str(GOLD / 'scenes' / 'car_boxed.rtsc')!r}); big = sc.instance_grid(2, 2, 1, (11.5, 6.5, 3.0)); big.build_bvh(6)
        f = big.flatten_host(); h = hashlib.sha256()
        for k in ("nodes", "nodes4", "tris", "shade", "leaf_cnt", "mats", "lights"): h.update(f[k].tobytes())
        print(h.hexdigest() + f":{{f['max_depth']}}:{{f['stack_need4']}}")
    """)
    import os
    env = dict(os.environ, RT_FLATTEN_THREADS="1", RT_BVH_THREADS="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    assert r.stdout.strip() == mine


@pytest.mark.parametrize("scene", ["soup2k", "car_only"])
def test_walking_the_flattened_records_finds_the_oracle_hits(rt, oracle_scenes, scene):
    """bvh_traverse (cpu/src/bvh.c:317-358) restated over the 64-byte records with plain numpy floats: first-hit triangle and t of
    a few hundred primary rays must be the oracle's."""
    f, a = flat_of(rt, scene)
    nodes, r2 = f["nodes"].reshape(-1, 16).astype(np.float32), refs2(f)
    tris = f["tris"].reshape(-1, 16)
    w, h = 48, 27
    ref = oracle_scenes[scene].render(w, h)
    import oracle as O
    basis = O.Oracle().camera_basis(O.DEFAULT_CAM_POS, O.DEFAULT_CAM_ROT, O.DEFAULT_FOV, w, h)
    pos, ul, incx, incy = (np.asarray(basis[k], np.float32) for k in range(4))  # rows: pos, upper-left corner, inc_x, inc_y
    F = np.float32
    EPS = F(1e-3)

    def box(o, d, mn, mx):
        with np.errstate(all="ignore"):
            t1, t2 = (mn - o) / d, (mx - o) / d
        tmin = max(max(min(t1[0], t2[0]), min(t1[1], t2[1])), min(t1[2], t2[2]))
        tmax = min(min(max(t1[0], t2[0]), max(t1[1], t2[1])), max(t1[2], t2[2]))
        return tmin if (tmax >= tmin and tmax > 0) else F(np.inf)

    def tri(o, d, j):
        v0, e1, e2, n = tris[j, 0:3], tris[j, 3:6], tris[j, 6:9], tris[j, 9:12]
        det = -F(d[0] * n[0] + d[1] * n[1] + d[2] * n[2])
        if abs(det) < EPS:
            return F(np.inf)
        ao = o - v0
        dao = np.array([ao[1] * d[2] - ao[2] * d[1], ao[2] * d[0] - ao[0] * d[2], ao[0] * d[1] - ao[1] * d[0]], F)
        u = F(e2[0] * dao[0] + e2[1] * dao[1] + e2[2] * dao[2]) / det
        v = -F(e1[0] * dao[0] + e1[1] * dao[1] + e1[2] * dao[2]) / det
        t = F(ao[0] * n[0] + ao[1] * n[1] + ao[2] * n[2]) / det
        return t if (t > EPS and u >= 0 and v >= 0 and u + v <= 1) else F(np.inf)

    bad = 0
    for y in range(0, h, 3):
        for x in range(0, w, 2):
            d = ((ul - pos) + incx * F(x)) + incy * F(y)
            best_t, best = F(np.inf), -1
            stack = [0]
            while stack:
                r = stack.pop()
                if r < 0:
                    first, cnt = leaf_range(f, r)
                    for j in range(first, first + cnt):
                        t = tri(pos, d, j)
                        if t < best_t:
                            best_t, best = t, int(tris[j, 12].view(np.int32))
                    continue
                tl = box(pos, d, nodes[r, [0, 1, 2]], nodes[r, [3, 4, 5]])
                tr = box(pos, d, nodes[r, [6, 7, 8]], nodes[r, [9, 10, 11]])
                (tn, rn), (tf, rf) = ((tl, r2[r, 0]), (tr, r2[r, 1])) if not (tr < tl) else ((tr, r2[r, 1]), (tl, r2[r, 0]))
                if tf < best_t and rf != REF_NONE: stack.append(int(rf))
                if tn < best_t and rn != REF_NONE: stack.append(int(rn))
            if best != ref["id"][y, x]:
                bad += 1
            elif best >= 0:
                assert abs(best_t - ref["depth"][y, x]) <= 1e-5 * abs(ref["depth"][y, x])
    assert bad == 0


def test_center_half_extent_boxes_contain_what_they_encode():
    """csrc/flatten.h: box_center_half restated in numpy (IEEE float32, round to nearest): the interval [c - h, c + h] evaluated
    exactly contains [mn, mx] for ordinary, tiny, huge, negative and degenerate intervals, and h == 0 only for mn == mx."""
    rng = np.random.default_rng(7)
    mag = np.float32(10.0) ** rng.integers(-30, 30, 200000).astype(np.float32)
    a = (rng.standard_normal(200000).astype(np.float32) * mag).astype(np.float32)
    w = np.abs(rng.standard_normal(200000).astype(np.float32)) * np.float32(10.0) ** rng.integers(-35, 30, 200000).astype(np.float32)
    w[::7] = 0                                                          # flat boxes (axis-aligned triangles)
    mn, mx = a, (a + w.astype(np.float32)).astype(np.float32)
    ok = np.isfinite(mn) & np.isfinite(mx) & (mx >= mn)
    mn, mx = mn[ok], mx[ok]
    c = (mn * np.float32(0.5) + mx * np.float32(0.5)).astype(np.float32)
    up, dn = (mx - c).astype(np.float32), (c - mn).astype(np.float32)
    h = np.maximum(up, dn)
    bump = (h > 0) & (h < np.float32(3.0e38))
    h = np.where(bump, (h.view(np.uint32) + np.uint32(1)).view(np.float32), h)
    c64, h64 = c.astype(np.float64), h.astype(np.float64)
    assert (c64 - h64 <= mn.astype(np.float64)).all() and (c64 + h64 >= mx.astype(np.float64)).all()
    assert ((h == 0) == (mn == mx)).all()
