import sys; sys.path.insert(0,'.')
import parallel_ray_tracer_b200 as rt
for scene,h in (("soup2k",0),("soup2k",6),("car_only",6)):
    sc = rt.Scene.load_rtsc(f"tests/golden/scenes/{scene}.rtsc").build_bvh(h); ctx = rt.Context(sc,[0])
    for mode in (0,1):
        for trav in (1,2,3):
            for (w,hh,spp) in ((97,53,1),(64,36,3)):
                tm = ctx.render_frame(rt.default_params(width=w,height=hh,spp=spp,mode=mode,traversal=trav,aov_mask=7))
                ctx.load_from_gpu(rgb=True,tri_id=True,depth=True)
    ctx.render_frame(rt.default_params(width=80,height=40,part_index=1,part_count=3)); ctx.packed_tiles()
    ctx.close()
print("memcheck workload done")
