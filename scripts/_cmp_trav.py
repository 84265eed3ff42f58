import sys; sys.path.insert(0,'.')
import numpy as np, parallel_ray_tracer_b200 as rt
for scene,w,h in (("car_only",1920,1080),("car_boxed",1920,1080),("car_boxed",3840,2160),("soup2k",1280,720)):
    sc = rt.Scene.load_rtsc(f"tests/golden/scenes/{scene}.rtsc").build_bvh(6); ctx = rt.Context(sc,[0])
    fr = {}
    for trav in (1,2,3):
        ctx.render_frame(rt.default_params(width=w,height=h,traversal=trav,aov_mask=7)); fr[trav]=ctx.load_from_gpu(rgb=True,tri_id=True,depth=True)
    for a,b in ((1,2),(1,3),(2,3)):
        print(scene,w,'trav',a,'vs',b,'id diff',int((fr[a]['id']!=fr[b]['id']).sum()),'depth diff',int((fr[a]['depth']!=fr[b]['depth']).sum()),'bgra diff px',int((fr[a]['bgra']!=fr[b]['bgra']).any(-1).sum()),'rgb diff',int((fr[a]['rgb']!=fr[b]['rgb']).any(-1).sum()))
    # default config twice (frame 1 may use another schedule than frame 2)
    p = rt.default_params(width=w,height=h,aov_mask=7)
    ctx.render_frame(p); f1=ctx.load_from_gpu(rgb=True,tri_id=True,depth=True); ctx.render_frame(p); f2=ctx.load_from_gpu(rgb=True,tri_id=True,depth=True)
    print(scene,w,'default frame1 vs frame2: bgra diff', int((f1['bgra']!=f2['bgra']).any(-1).sum()), 'id diff', int((f1['id']!=f2['id']).sum()))
    ctx.close()
