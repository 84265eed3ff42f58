"""Host-side scene layer of the product (C++ behind the C-ABI): loaders, scene pack, BVH builder,
BMP writer.  No GPU needed."""
import hashlib

import numpy as np
import pytest

import oracle as O
from conftest import GOLD, SCENES


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_obj_loader_rules(rt, tmp_path):
    """The reference loader's rules (cpu/src/triangle.c:54-126) on a hand-written file."""
    (tmp_path / "triangles.obj").write_text(
        "# comment\nmtllib triangles.mtl.mtl\no Thing\n"
        "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\n"
        "f 1 2 3\n"                 # before any usemtl: all-zero material
        "usemtl red\nf 1 2 4\n"
        "usemtl nosuch\nf 2 3 4\n"  # unknown name keeps the previous material
        "s 0\nusemtl late\nf 1 3 4\n")
    (tmp_path / "triangles.mtl").write_text(
        "newmtl red\nNs 1\nKa 1 1 1\nKd 0.8 0.1 0.1\nKs 0.5 0.5 0.5\nKr 0.2 0.2 0.2\nKe 0 0 0\n\n"
        "newmtl late\nNs 1\nKa 1 1 1\nKd 0.1 0.2 0.3\nKs 0.4 0.5 0.6\nKe 0 0 0\nKr 0.9 0.9 0.9\n")  # Kr on the 6th line: ignored
    (tmp_path / "lights.obj").write_text("0 -8 3 50 50 50\n\n1 2 3 4 5 6")
    sc = rt.Scene.load_dir(tmp_path)
    a = sc.arrays()
    assert a["tri"].shape == (4, 9)
    assert np.array_equal(a["tri"][1], [0, 0, 0, 1, 0, 0, 0, 0, 1])
    per_tri = a["mats"][a["mat_idx"]]
    assert np.array_equal(per_tri[0], np.zeros(9, np.float32))
    assert np.allclose(per_tri[1], [0.5, 0.5, 0.5, 0.8, 0.1, 0.1, 0.2, 0.2, 0.2])
    assert np.array_equal(per_tri[2], per_tri[1])
    assert np.allclose(per_tri[3], [0.4, 0.5, 0.6, 0.1, 0.2, 0.3, 0, 0, 0])  # Kr beyond the 5-line window stays 0
    assert np.array_equal(a["lights"], np.array([[0, -8, 3, 50, 50, 50], [1, 2, 3, 4, 5, 6]], np.float32))
    assert np.allclose(a["ambient"], 0.5)


def test_loader_errors_do_not_exit(rt, tmp_path):
    with pytest.raises(rt.RtError) as e:
        rt.Scene.load_dir(tmp_path / "missing")
    assert e.value.code == rt.RT_ERR_IO
    (tmp_path / "triangles.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1/1/1 2/2/2 3/3/3\n")
    (tmp_path / "triangles.mtl").write_text("")
    (tmp_path / "lights.obj").write_text("")
    with pytest.raises(rt.RtError) as e:
        rt.Scene.load_dir(tmp_path)   # v/vt/vn faces are undefined behaviour in the reference; an error here
    assert e.value.code == rt.RT_ERR_IO


@pytest.mark.parametrize("scene", ["car_only", "car_boxed"])
def test_obj_loader_equals_reference_loader(rt, scene, scene_arrays):
    """Our OBJ/MTL/lights parser vs the scene pack dumped from the reference's own triangles_load."""
    d = O.RefCpu.scene_dir(scene)
    if not (d / "triangles.obj").exists():
        pytest.skip("reference assets not staged (oracle/_ref/assets)")
    a = rt.Scene.load_dir(d).arrays()
    g = scene_arrays[scene]
    assert np.array_equal(a["tri"], g["tri"])
    assert np.array_equal(a["mats"][a["mat_idx"]], g["mats"][g["mat_idx"]])
    assert np.array_equal(a["lights"], g["lights"])


def test_rtsc_roundtrip(rt, tmp_path, scene_arrays):
    sc = rt.Scene.load_rtsc(GOLD / "scenes" / "soup2k.rtsc")
    sc.save_rtsc(tmp_path / "x.rtsc")
    assert (tmp_path / "x.rtsc").read_bytes() == (GOLD / "scenes" / "soup2k.rtsc").read_bytes()
    (tmp_path / "bad.rtsc").write_bytes(b"RTSC0001" + b"\x00" * 7)
    with pytest.raises(rt.RtError):
        rt.Scene.load_rtsc(tmp_path / "bad.rtsc")


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("heuristic", [6, 1, 0])
def test_bvh_builder_equals_oracle_node_for_node(rt, oracle_scenes, orc, scene_arrays, scene, heuristic):
    """O(n)-per-node binned builder vs the oracle's literal O(96 n) restatement of bvh_split."""
    sc = rt.Scene.load_rtsc(GOLD / "scenes" / f"{scene}.rtsc").build_bvh(heuristic)
    a = sc.arrays()
    s = oracle_scenes[scene] if heuristic == 6 else orc.scene(scene_arrays[scene])
    nodes, ti = s.bvh() if heuristic == 6 else s.build_bvh(heuristic)
    assert np.array_equal(a["bvh_nodes"], nodes)
    assert np.array_equal(a["tri_idx"], ti)


@pytest.mark.parametrize("scene", SCENES)
def test_bvh_builder_reproduces_the_reference_binary_tree(rt, manifest, scene):
    """RT_BVH_REFBIN: digest of bvh[] / tri_idx[] as dumped from the reference CPU binary."""
    a = rt.Scene.load_rtsc(GOLD / "scenes" / f"{scene}.rtsc").build_bvh(6 | rt.RT_BVH_REFBIN).arrays()
    m = manifest["scenes"][scene]
    assert len(a["bvh_nodes"]) // 32 == m["bvh_nodes"]
    assert sha(a["bvh_nodes"]) == m["bvh_nodes_sha256"]
    assert sha(a["tri_idx"]) == m["tri_idx_sha256"]


def test_parallel_build_equals_serial_build(rt, monkeypatch):
    """>= 200 000 triangles take the parallel path (subtrees built concurrently, reference numbering assigned
    afterwards): the arrays must equal the serial build's, node for node."""
    base = rt.Scene.load_rtsc(GOLD / "scenes" / "car_only.rtsc")
    big = base.instance_grid(3, 3, 1, (11.0, 6.5, 3.0))
    monkeypatch.setenv("RT_BVH_THREADS", "1")
    a = big.build_bvh(6).arrays()
    monkeypatch.setenv("RT_BVH_THREADS", "6")
    b = big.build_bvh(6).arrays()
    assert len(a["tri"]) == 9 * 32136 and len(a["bvh_nodes"]) // 32 > 300000
    assert np.array_equal(a["bvh_nodes"], b["bvh_nodes"]) and np.array_equal(a["tri_idx"], b["tri_idx"])
    monkeypatch.setenv("RT_BVH_THREADS", "3")
    c = big.build_bvh(6 | rt.RT_BVH_REFBIN).arrays()
    monkeypatch.setenv("RT_BVH_THREADS", "1")
    d = big.build_bvh(6 | rt.RT_BVH_REFBIN).arrays()
    assert np.array_equal(c["bvh_nodes"], d["bvh_nodes"]) and np.array_equal(c["tri_idx"], d["tri_idx"])


def test_bvh_invariants(rt):
    a = rt.Scene.load_rtsc(GOLD / "scenes" / "car_only.rtsc").build_bvh(6).arrays()
    dt = np.dtype([("min", "3f4"), ("max", "3f4"), ("len", "i4"), ("idx", "i4")])
    n = a["bvh_nodes"].view(dt)
    leaves = n[n["len"] > 0]
    assert leaves["len"].sum() == len(a["tri"])                       # every triangle in exactly one leaf
    assert np.array_equal(np.sort(a["tri_idx"]), np.arange(len(a["tri"])))
    assert leaves["len"].max() <= 2
    inner = n[(n["len"] == 0) & (n["idx"] != 0)]
    kids = np.concatenate([inner["idx"], inner["idx"] + 1])
    assert len(np.unique(kids)) == len(kids) == len(n) - 1            # a tree: every non-root node has one parent
    tri = a["tri"].reshape(-1, 3, 3)
    for node in leaves[:200]:
        t = tri[a["tri_idx"][node["idx"]:node["idx"] + node["len"]]].reshape(-1, 3)
        assert np.array_equal(t.min(0), node["min"]) and np.array_equal(t.max(0), node["max"])


def test_builder_rejects_unsupported(rt):
    sc = rt.Scene.load_rtsc(GOLD / "scenes" / "soup2k.rtsc")
    for h in (2, 3, 4, 5, 7):
        with pytest.raises(rt.RtError):
            sc.build_bvh(h)


def test_soup_generator(rt):
    a = rt.Scene.soup(500, 1).arrays()
    t = a["tri"].reshape(-1, 3, 3)
    assert t.shape[0] == 500 and np.all(t[:, 0] >= -5) and np.all(t[:, 0] <= 5)
    assert np.all(t[:, 1] - t[:, 0] >= 0) and np.all(t[:, 1] - t[:, 0] <= 1.0001)
    assert np.array_equal(a["mats"], np.array([[1, 1, 1, 0, 0, 0, 0, 0, 0]], np.float32)) and len(a["lights"]) == 0
    b = rt.Scene.soup(500, 1).arrays()
    assert np.array_equal(a["tri"], b["tri"])


def test_instance_grid(rt):
    base = rt.Scene.load_rtsc(GOLD / "scenes" / "soup2k.rtsc")
    g = base.instance_grid(3, 2, 1, (20.0, 20.0, 20.0), light_every=0).arrays()
    b = base.arrays()
    assert g["tri"].shape[0] == 6 * b["tri"].shape[0]
    assert np.allclose(g["tri"][:2000].reshape(-1, 3, 3)[:, :, 0] + 20.0, b["tri"].reshape(-1, 3, 3)[:, :, 0], atol=1e-5)
    g2 = base.instance_grid(3, 2, 1, (20.0, 20.0, 20.0)).build_bvh(6).arrays()
    assert len(g2["bvh_nodes"]) // 32 > 6 * 2000


def test_bmp_writer_is_byte_identical_to_the_reference_writer(rt, tmp_path):
    """Re-encode the pixels of a BMP written by the reference's bmp_write_file (cpu/src/bmp_writer.c)."""
    ref = (GOLD / "ref_car_only_64x36.bmp").read_bytes()
    w, h = 64, 36
    assert len(ref) == 54 + 4 * w * h
    rows_bottom_up = np.frombuffer(ref, np.uint8, 4 * w * h, 54).reshape(h, w, 4)
    top_down = rows_bottom_up[::-1].copy()
    rt.write_bmp(tmp_path / "o.bmp", top_down)
    assert (tmp_path / "o.bmp").read_bytes() == ref


def _load_both(rt, d, monkeypatch, chunks):
    monkeypatch.setenv("RT_LOADER_SERIAL", "1")
    a = rt.Scene.load_dir(d).arrays()
    monkeypatch.delenv("RT_LOADER_SERIAL")
    monkeypatch.setenv("RT_LOADER_CHUNKS", str(chunks))
    b = rt.Scene.load_dir(d).arrays()
    monkeypatch.delenv("RT_LOADER_CHUNKS")
    return a, b


@pytest.mark.parametrize("chunks", [1, 3, 64])
def test_parallel_loader_equals_the_line_by_line_loader(rt, tmp_path, monkeypatch, chunks):
    """The memory-speed loader (whole-file read, line index, chunks parsed on threads) against the sscanf-per-line one, on a
    file built to sit on the rules: faces before their vertices, usemtl changes at chunk borders, CRLF, exponents, missing
    coordinates, lines longer than the reference's 256-byte fgets buffer, unknown material names."""
    rng = np.random.default_rng(5)
    n_v = 400
    lines = ["# header\r\n", "mtllib x.mtl\n", "f 1 2 3\n"]  # a face before any vertex
    for i in range(n_v):
        x, y, z = rng.normal(size=3)
        style = i % 5
        if style == 0: lines.append(f"v {x:.6f} {y:.6f} {z:.6f}\n")
        elif style == 1: lines.append(f"v {x:.9e}  {y:.3E}\t{z:+.5f}\r\n")
        elif style == 2: lines.append(f"v {x:.4f} {y:.4f}\n")                       # z missing -> 0
        elif style == 3: lines.append(f"v   {x:.7f} {y:.7f} {z:.7f} 1.0 # w\n")
        else: lines.append("v " + " " * 260 + f"{x:.6f} {y:.6f} {z:.6f}\n")        # > 255 characters: split by fgets
        if i % 7 == 0: lines.append(f"vn {x:.3f} {y:.3f} {z:.3f}\n")
    names = ["red", "late", "nosuch", "red2"]
    for i in range(900):
        if i % 11 == 0: lines.append(f"usemtl {names[(i // 11) % 4]}\n")
        a, b, c = rng.integers(1, n_v + 1, size=3)
        lines.append(f"f {a} {b} {c}\n" if i % 3 else f"f  {a}\t{b} {c} \r\n")
        if i % 50 == 0: lines.append("s off\n")
    (tmp_path / "triangles.obj").write_text("".join(lines), newline="")
    (tmp_path / "triangles.mtl").write_text("newmtl red\nKd 0.8 0.1 0.1\nKs 0.5 0.5 0.5\nKr 0.2 0.2 0.2\n\nnewmtl late\nKd 0.1 0.2 0.3\n\nnewmtl red2\nKs 1 1 1\n")
    (tmp_path / "lights.obj").write_text("0 -8 3 50 50 50\n")
    a, b = _load_both(rt, tmp_path, monkeypatch, chunks)
    assert a["tri"].shape == (901, 9)
    for k in a:
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), k


def test_parallel_loader_reports_the_first_bad_face(rt, tmp_path, monkeypatch):
    body = "".join(f"v {i} 0 0\n" for i in range(50)) + "".join(f"f {1 + i % 48} {2 + i % 48} {3 + i % 48}\n" for i in range(300))
    body = body.replace("f 5 6 7\n", "f 5 6 99\n", 1) + "f 1/1 2/2 3/3\n"
    (tmp_path / "triangles.obj").write_text(body)
    (tmp_path / "triangles.mtl").write_text("")
    (tmp_path / "lights.obj").write_text("")
    msgs = []
    for env in ({"RT_LOADER_SERIAL": "1"}, {"RT_LOADER_CHUNKS": "16"}):
        for k, v in env.items(): monkeypatch.setenv(k, v)
        with pytest.raises(rt.RtError) as e:
            rt.Scene.load_dir(tmp_path)
        assert e.value.code == rt.RT_ERR_IO
        msgs.append(str(e.value))
        for k in env: monkeypatch.delenv(k)
    assert msgs[0] == msgs[1] and "f 5 6 99" in msgs[0]
