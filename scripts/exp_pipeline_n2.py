"""Per-phase wall times of the pipelined multi-process loop (torchrun, 2 ranks)."""
import os, sys, time, json; sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
import parallel_ray_tracer_b200 as rt
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W, H = 1920, 1080
sc = rt.Scene.load_rtsc('tests/golden/scenes/car_only.rtsc').build_bvh(6); ctx = rt.Context(sc, [local])
for slot in range(2):
    h = torch.zeros(64, dtype=torch.uint8, device=dev)
    if rank == 0: h.copy_(torch.frombuffer(bytearray(ctx.frame_ipc_export(W, H, slot)), dtype=torch.uint8))
    dist.broadcast(h, 0)
    if rank != 0: ctx.frame_ipc_import(bytes(h.cpu().numpy().tobytes()), W, H, slot)
p = [rt.default_params(width=W, height=H, part_index=rank, part_count=world, frame_slot=s) for s in range(2)]
pinned = [torch.empty(W * H * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]
for _ in range(20): ctx.render_frame(p[0]); dist.barrier()
def loop(n, mode):
    ph = np.zeros(5); dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        s = k & 1
        a = time.perf_counter(); ctx.render_frame_async(p[s]); b = time.perf_counter(); ctx.frame_wait(s); c = time.perf_counter()
        if rank == 0 and k >= 1 and mode != "nocopy": ctx.frame_wait(1 - s)
        d = time.perf_counter()
        if mode == "gloo_like":
            t = torch.zeros(1, device=dev); dist.all_reduce(t); t.item()
        else:
            dist.barrier()
        e = time.perf_counter()
        if rank == 0 and mode != "nocopy": ctx.download_async(s, pinned[s].data_ptr())
        f = time.perf_counter()
        ph += [b - a, c - b, d - c, e - d, f - e]
    if rank == 0 and mode != "nocopy":
        for s in range(2): ctx.frame_wait(s)
    tot = (time.perf_counter() - t0) * 1e3 / n
    print(json.dumps({"rank": rank, "mode": mode, "ms_per_frame": tot, "enqueue": ph[0] * 1e3 / n, "wait_render": ph[1] * 1e3 / n, "wait_copy": ph[2] * 1e3 / n,
                      "barrier": ph[3] * 1e3 / n, "download_enqueue": ph[4] * 1e3 / n}), flush=True)
for mode in ("nocopy", "copy", "gloo_like", "copy"):
    loop(40, mode)
dist.barrier(); dist.destroy_process_group()
