import sys, time, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import wrt_b200 as W, scenes, util
os.environ["WRT_KD_PROFILE"]="1"
sc=scenes.synthetic_torus_scene(n=2237,width=64,height=64,n_spheres=100000)
t=time.time(); hs=util.host_scene(W,sc); print('C5 host_scene (KD build incl. array hand-over): %.2f s, cores %d' % (time.time()-t, os.cpu_count()))
