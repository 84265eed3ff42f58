"""Shared fixtures.  `-m "not gpu"` covers the oracle, golden vectors, host logic and the ABI;
`-m gpu` tests are the parity tests proper and call through the C-ABI."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"

import oracle as O  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def rt():
    import parallel_ray_tracer_b200 as rt
    if not rt.LIB_PATH.exists():
        rt.build()
    rt.lib()
    return rt


@pytest.fixture(scope="session")
def orc():
    return O.Oracle()


@pytest.fixture(scope="session")
def manifest():
    return json.loads((GOLD / "manifest.json").read_text())


@pytest.fixture(scope="session")
def refcpu():
    r = O.RefCpu(6)
    if not r.available:
        pytest.skip("oracle/_ref (reference CPU renderer) not built on this machine")
    return r


SCENES = ("car_only", "car_boxed", "soup2k")


@pytest.fixture(scope="session")
def scene_arrays():
    return {n: O.load_rtsc(GOLD / "scenes" / f"{n}.rtsc") for n in SCENES}


@pytest.fixture(scope="session")
def oracle_scenes(orc, scene_arrays):
    """Oracle scenes with the IEEE heuristic-6 tree (built by the oracle's literal O(96 n) restatement)."""
    out = {}
    for n, sc in scene_arrays.items():
        s = orc.scene(sc)
        s.build_bvh(6)
        out[n] = s
    return out


def cam_of(manifest, name):
    c = manifest["cams"][name]
    if c is None:
        return (O.DEFAULT_CAM_POS, O.DEFAULT_CAM_ROT, O.DEFAULT_FOV)
    return (tuple(c["pos"]), tuple(c["rot"]), c["fov"])


def load_gold(name):
    z = np.load(GOLD / name)
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def gpu_scenes(rt, scene_arrays):
    """Library scenes + device contexts (cuda:0), heuristic-6 IEEE tree built by the product builder."""
    if rt.device_count() < 1:
        pytest.skip("no CUDA device")
    out = {}
    for n in SCENES:
        sc = rt.Scene.load_rtsc(GOLD / "scenes" / f"{n}.rtsc").build_bvh(6)
        out[n] = (sc, rt.Context(sc, [0]))
    yield out
    for sc, ctx in out.values():
        ctx.close()
        sc.close()
