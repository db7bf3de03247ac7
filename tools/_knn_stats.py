import sys, os, ctypes as C
sys.path.insert(0, '/root/repo')
import torch, pmb200, numpy as np
n=4194304; k=int(sys.argv[1]); W,H=1920,1080
m = pmb200.PhotonMapper(n_photons=n)
sc = pmb200.default_scene(sz_img=H); sc.cam_ox = -(W-H)/2.0; m.set_scene(sc)
m.init_random_numbers(); m.set_record_capacity(int(2.6*n)); m.clear_map(); m.trace(0.0, media=False, records=True, no_map=True); m.knn_build(0)
rgbf = torch.zeros((H,W,4),dtype=torch.float32,device='cuda')
L = pmb200.lib()
out = (C.c_ulonglong*8)()
L.pm_debug_knn_stats(out, 1)
m.render_knn(W,H,0.0,False,k,float('inf'),1e-4,1e-2,rgbf=rgbf); m.sync()
L.pm_debug_knn_stats(out, 1)
q = out[4]
print("k",k,"queries",q,"leaves/q",out[0]/q,"node steps/q",out[1]/q,"cand passed/q",out[2]/q,"merges/q",out[3]/q)
