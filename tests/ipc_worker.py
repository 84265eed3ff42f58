"""Second process of tests/test_gpu_ipc.py: rank 1 of a 2-rank tile split on ONE device.  It maps rank 0's frame (CUDA IPC
handles read from stdin, one hex line per frame slot), renders its interleaved tiles straight into it — exactly what
bench.py's ranks > 0 do over NVLink — and reports.  Protocol on stdout: "ready" after the import, "done <rays>" per frame."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import parallel_ray_tracer_b200 as rt  # noqa: E402


def main():
    scene, w, h, parts, part, n_slots, traversal = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
    sc = rt.Scene.load_rtsc(ROOT / "tests" / "golden" / "scenes" / f"{scene}.rtsc").build_bvh(6)
    ctx = rt.Context(sc, [0])
    for slot in range(n_slots):
        ctx.frame_ipc_import(bytes.fromhex(sys.stdin.readline().strip()), w, h, slot)
    print("ready", flush=True)
    for line in sys.stdin:
        cmd = line.split()
        if not cmd or cmd[0] == "quit":
            break
        slot = int(cmd[1])
        tm = ctx.render_frame(rt.default_params(width=w, height=h, part_index=part, part_count=parts, frame_slot=slot, traversal=traversal))
        print("done", tm.rays_closest + tm.rays_shadow, flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
