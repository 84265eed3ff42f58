import sys; sys.path.insert(0, '.')
import parallel_ray_tracer_b200 as rt
sc = rt.Scene.load_rtsc('tests/golden/scenes/car_only.rtsc').build_bvh(6); ctx = rt.Context(sc, [0])
p = rt.default_params(width=1920, height=1080)
for _ in range(6): tm = ctx.render_frame(p)
print(tm.kernel_ms)
