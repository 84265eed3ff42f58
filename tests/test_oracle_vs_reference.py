"""Pin the oracle restatement to the reference itself (oracle/_ref = the unmodified reference CPU
renderer, compiled by oracle/Makefile).  Skipped where oracle/_ref is absent."""
import numpy as np
import pytest

import oracle as O
from conftest import GOLD


@pytest.mark.parametrize("scene", ["car_only", "car_boxed"])
def test_oracle_vs_reference_frame(refcpu, oracle_scenes, scene):
    w, h = 480, 270
    ref = refcpu.run(rtsc=GOLD / "scenes" / f"{scene}.rtsc", width=w, height=h)
    got = oracle_scenes[scene].render(w, h)
    m = O.compare_aovs(got, ref)
    assert m["id_match"] >= 0.9999 and m["rgb8_within1"] >= 0.999 and m["depth_within_1e-4"] >= 0.9999, m


def test_oracle_vs_reference_soup(refcpu, orc, tmp_path):
    """The reference's own synthetic scene (argv[2], cpu/src/main.c:115-131): incoherent worst case."""
    rtsc = tmp_path / "soup.rtsc"
    ref = refcpu.run(soup=20000, width=200, height=120, dump_scene=rtsc)
    s = orc.scene(O.load_rtsc(rtsc))
    s.build_bvh(6)
    got = s.render(200, 120)
    m = O.compare_aovs(got, ref)
    assert m["id_match"] >= 0.9995 and m["depth_within_1e-4"] >= 0.9995, m


@pytest.mark.parametrize("scene", ["car_only", "car_boxed"])
def test_refbin_tree_equals_reference_binary_tree(refcpu, orc, scene_arrays, scene, tmp_path):
    f = tmp_path / "t.bvh"
    refcpu.run(rtsc=GOLD / "scenes" / f"{scene}.rtsc", width=8, height=8, frames=0, aov=False, dump_bvh=f)
    rn, rti = O.load_bvh_dump(f)
    s = orc.scene(scene_arrays[scene])
    n, ti = s.build_bvh(6 | 0x100)
    assert np.array_equal(n, rn) and np.array_equal(ti, rti)


def test_reference_image_is_thread_count_independent(refcpu):
    a = refcpu.run(rtsc=GOLD / "scenes" / "car_only.rtsc", width=160, height=90, threads=1)
    b = refcpu.run(rtsc=GOLD / "scenes" / "car_only.rtsc", width=160, height=90, threads=4)
    assert np.array_equal(a["bgra"], b["bgra"]) and np.array_equal(a["id"], b["id"])
