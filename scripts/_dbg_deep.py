import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O, parallel_ray_tracer_b200 as rt
f='tests/golden/scenes/soup2k.rtsc'
for h in (0,1,6):
    sc = rt.Scene.load_rtsc(f).build_bvh(h); a = sc.arrays()
    dt = np.dtype([("min", "3f4"), ("max", "3f4"), ("len", "i4"), ("idx", "i4")]); n=a["bvh_nodes"].view(dt)
    ctx = rt.Context(sc,[0]); s = O.Oracle().scene(O.load_rtsc(f)); s.set_bvh(a["bvh_nodes"], a["tri_idx"])
    ref = s.render(120,68)
    for mode in (1,0):
        ctx.render_frame(rt.default_params(width=120,height=68,mode=mode,aov_mask=7)); got=ctx.load_from_gpu(rgb=True,tri_id=True,depth=True)
        bad = np.argwhere(got['id']!=ref['id'])
        print('h',h,'nodes',len(n),'maxleaf',n['len'].max(),'mode',mode,'id mismatches',len(bad),'bgra mism',(got['bgra']!=ref['bgra']).any(-1).sum(), 'depth mism', (got['depth']!=ref['depth']).sum())
        for y,x in bad[:5]: print('   px',x,y,'got',got['id'][y,x],got['depth'][y,x],'ref',ref['id'][y,x],ref['depth'][y,x])
