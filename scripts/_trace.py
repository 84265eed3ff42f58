import sys; sys.path.insert(0,'.')
import numpy as np, parallel_ray_tracer_b200 as rt
for scene,w,h in (('car_only',1920,1080),):
    sc = rt.Scene.load_rtsc(f'tests/golden/scenes/{scene}.rtsc').build_bvh(6); ctx = rt.Context(sc,[0])
    ctx.warp_trace(True)
    for trav,fb in ((2,2),(2,1)):
        p = rt.default_params(width=w,height=h,aov_mask=rt.RT_AOV_WORK,traversal=trav,tile_feedback=fb,ctas_per_sm=6)
        for _ in range(4): tm = ctx.render_frame(p)
        t = ctx.warp_trace(True).astype(np.float64)
        t0 = t[:,0].min(); start=(t[:,0]-t0); empty=(t[:,1]-t0); exit_=(t[:,2]-t0)
        dur = exit_.max()
        print(scene, 'trav',trav,'fb',fb,'kernel_ms',round(tm.kernel_ms[0],3),'warps',len(t),'dur cycles',dur)
        print('  start max', start.max()/dur, ' queue-empty: median',np.median(empty)/dur,'min',empty.min()/dur,' exit: mean',exit_.mean()/dur,'median',np.median(exit_)/dur,'p10',np.percentile(exit_,10)/dur,'p90',np.percentile(exit_,90)/dur)
        print('  chunks/warp mean',t[:,3].mean(),'max',t[:,3].max(),'min',t[:,3].min(),' iters/warp mean',t[:,4].mean(),'max',t[:,4].max(), ' lanes per inner-phase iter', t[:,5].sum()/max(1,t[:,4].sum()), 'tri', t[:,6].sum()/max(1,t[:,4].sum()))
        sm = t[:,7].astype(int); ex = np.array([exit_[sm==s].max() for s in np.unique(sm)])
        print('  per-SM last exit: min',ex.min()/dur,'mean',ex.mean()/dur,' cycles/iter (mean over warps)', (exit_/np.maximum(t[:,4],1)).mean())
        hist=np.histogram(exit_/dur,bins=10,range=(0,1))[0]; print('  exit histogram',hist, ' empty histogram', np.histogram(empty/dur,bins=10,range=(0,1))[0])
        late = np.argsort(-exit_)[:5]; print('  latest warps: exit',exit_[late]/dur,'empty',empty[late]/dur,'iters',t[late,4],'chunks',t[late,3])
