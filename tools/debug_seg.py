import sys, os, numpy as np
os.environ['PM_DEBUG_SEG']='1'
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import __graft_entry__ as ge
pm=ge.load_package()
W=8192
scene=pm.build_scene(pm.SCENE_TIGER,W,W)
n=int(scene[:4].view(np.uint32)[0]); ix=int(scene[4:8].view(np.uint32)[0])
it=scene[ix:ix+32*n].view(np.uint32).reshape(n,8)
nseg=np.where(it[:,0]==3, it[:,3], np.where(it[:,0]==4, it[:,3]-1, 0))
pref=np.concatenate([[0],np.cumsum(nseg)])
r=pm.PietRenderer(0); r.drawable_size_will_change(W,W); r.init_scene(scene)
for _ in range(2):
    r.draw(); st=r.sync(); print(st.ms_bin, st.ms_fine)
for blk in [int(x) for x in sys.argv[1:]]:
    g=blk*256; i=int(np.searchsorted(pref,g,side='right')-1); print('block',blk,'item',i,'tag',it[i,0],'npts',it[i,3],'seg',g-pref[i])
