"""The C-ABI library loads and exports every symbol include/rt_b200.h declares (no compute calls)."""
import re
import subprocess

import pytest

from conftest import ROOT


def declared_symbols():
    text = (ROOT / "include" / "rt_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", text))
    names.discard("rt_tile_owner")  # static inline in the header
    return names


def test_header_and_binding_agree(rt):
    assert declared_symbols() == set(rt.EXPORTS)


def test_library_exports_every_declared_symbol(rt):
    out = subprocess.run(["nm", "-D", "--defined-only", str(rt.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = declared_symbols() - exported
    assert not missing, f"librt_b200.so lacks {sorted(missing)}"
    lib = rt.lib()
    for name in rt.EXPORTS:
        assert getattr(lib, name) is not None
    assert lib.rt_abi_version() == 3


def test_no_torch_or_cxx_types_cross_the_boundary():
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "rt_b200.h").read_text(), flags=re.S)  # code only
    for banned in ("std::", "torch", "at::Tensor", "cudaStream_t", "template"):
        assert banned not in text


def test_struct_sizes_match_the_header(rt, tmp_path):
    """ctypes mirrors vs. the C compiler's view of include/rt_b200.h."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "rt_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(rt_bvh_node),'
                   ' sizeof(rt_scene_desc), sizeof(rt_camera), sizeof(rt_render_params), sizeof(rt_timing));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    import ctypes as C
    assert sizes == [C.sizeof(rt.rt_bvh_node), C.sizeof(rt.rt_scene_desc), C.sizeof(rt.rt_camera),
                     C.sizeof(rt.rt_render_params), C.sizeof(rt.rt_timing)]


def test_no_cpu_fallback(rt, scene_arrays):
    """Without a CUDA device the product refuses to render (it must not route through the oracle)."""
    if rt.device_count() > 0:
        pytest.skip("a GPU is present")
    sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / "soup2k.rtsc").build_bvh(6)
    with pytest.raises(rt.RtError) as e:
        rt.Context(sc)
    assert e.value.code == rt.RT_ERR_NO_DEVICE


def test_gpu_only_entry_points_fail_loudly_without_a_device(rt):
    """The GPU BVH build, pinned allocation and the roofline microbenchmark have no host stand-in either."""
    if rt.device_count() > 0:
        pytest.skip("a GPU is present")
    sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / "soup2k.rtsc")
    with pytest.raises(rt.RtError) as e:
        sc.build_bvh_gpu(6)
    assert e.value.code == rt.RT_ERR_NO_DEVICE
    with pytest.raises(rt.RtError) as e:
        rt.PinnedBuffer(4096)
    assert e.value.code == rt.RT_ERR_NO_DEVICE
    with pytest.raises(rt.RtError) as e:
        rt.gather_bandwidth(1 << 20)
    assert e.value.code == rt.RT_ERR_NO_DEVICE
    with pytest.raises(rt.RtError) as e:
        sc.build_bvh_gpu(1)   # only heuristic 6 is built on the GPU
    assert e.value.code == rt.RT_ERR_INVALID


def test_bmp_writers_agree(rt, tmp_path):
    """rt_write_bmp (top-down input, flips) and rt_write_bmp_bottom_up (kernel-ordered rows, no flip) give the same file,
    with the reference's 54-byte header (cpu/src/bmp_writer.c:97-175)."""
    import numpy as np
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    rt.write_bmp(tmp_path / "a.bmp", img)
    rt.write_bmp_bottom_up(tmp_path / "b.bmp", img[::-1])
    a, b = (tmp_path / "a.bmp").read_bytes(), (tmp_path / "b.bmp").read_bytes()
    assert a == b and len(a) == 54 + 37 * 53 * 4 and a[:2] == b"BM"
    assert int.from_bytes(a[2:6], "little") == len(a) and int.from_bytes(a[10:14], "little") == 54
    assert int.from_bytes(a[18:22], "little") == 53 and int.from_bytes(a[22:26], "little") == 37 and a[28] == 32
    rows = np.frombuffer(a, np.uint8, 37 * 53 * 4, 54).reshape(37, 53, 4)
    assert np.array_equal(rows[::-1], img)


def test_product_does_not_reference_the_oracle():
    """Nothing under the package (sources or Python) may import, link or execute oracle/."""
    pkg = ROOT / "parallel_ray_tracer_b200"
    for p in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        if p.is_file() and p.suffix in (".py", ".cpp", ".cu", ".cuh", ".h", ".inl", "") and p.name != "Makefile":
            t = p.read_text(errors="ignore")
            assert "librt_oracle" not in t and "import oracle" not in t and "ref_cpu_h" not in t, p
    mk = "\n".join(l for l in (pkg / "csrc" / "Makefile").read_text().splitlines() if not l.lstrip().startswith("#"))
    assert "oracle" not in mk  # no rule compiles or links anything from oracle/
