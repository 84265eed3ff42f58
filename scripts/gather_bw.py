"""L1 / L2 / HBM random 64-byte gather bandwidth of the device: one JSON line."""
import sys, json; sys.path.insert(0, '.')
import parallel_ray_tracer_b200 as rt
out = {}
for name, ws in (("64KB", 64 << 10), ("1MB", 1 << 20), ("4MB", 4 << 20), ("32MB", 32 << 20), ("96MB", 96 << 20), ("1GB", 1 << 30), ("8GB", 8 << 30)):
    out[name] = round(rt.gather_bandwidth(ws), 1)
print(json.dumps({"random_64B_gather_GBs": out}))
