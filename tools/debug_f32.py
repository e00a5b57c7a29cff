import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import __graft_entry__ as ge, oracle_api as oa
pm=ge.load_package()
W,H=2048,1536
scene=pm.build_scene(pm.SCENE_CARDIOID,W,H)
r=pm.PietRenderer(0); r.drawable_size_will_change(W,H); r.init_scene(scene); r.draw(); r.sync()
a=r.read_rgba32f(); u=r.read_rgba8()
o=oa.render(scene,W,H,f32=True)
b=o['rgba32f']
d=np.abs(a-b).max(axis=2)
ys,xs=np.nonzero(d>1e-5)
print('n bad px',len(ys),'of',W*H)
if len(ys):
    tiles=set(zip((ys//16).tolist(),(xs//16).tolist()))
    print('bad tiles',len(tiles), sorted(tiles)[:20])
    for k in range(min(5,len(ys))):
        y,x=ys[k],xs[k]; print(y,x,'gpu',a[y,x],'ref',b[y,x],'u8',u[y,x],o['rgba8'][y,x])
    # per-tile pattern: which pixel positions within tiles
    print('rows in tile', sorted(set((ys%16).tolist())), 'cols', sorted(set((xs%16).tolist())))
