import sys, os, ctypes as C
sys.path.insert(0, '/root/repo')
import torch, pmb200, numpy as np
L = pmb200.lib()
out = (C.c_ulonglong*8)()
def run(name, pts, q, k):
  for curve in (0, 1):
    m = pmb200.PhotonMapper(n_photons=16); m.knn_set_curve(curve); name = name[:24] + (' Z' if curve == 0 else ' H')
    tp = torch.from_numpy(pts).cuda(); tq = torch.from_numpy(q).cuda()
    m.knn_build_points(0, tp, tp, pts.shape[0]); m.sync()
    rgb = torch.empty((q.shape[0],4),dtype=torch.float32,device='cuda')
    L.pm_debug_knn_stats(out, 1)
    m.knn_radiance(0, tq, q.shape[0], k, float('inf'), rgb); m.sync()
    L.pm_debug_knn_stats(out, 1)
    nq = out[4]
    print(f"{name:28s} n={pts.shape[0]:9d} k={k:3d} leaves/q {out[0]/nq:6.1f} nodes/q {out[1]/nq:6.1f} passed/q {out[2]/nq:7.1f} merges/q {out[3]/nq:5.1f}")
    m.close()
rng = np.random.default_rng(0)
n = 4_000_000
# uniform on the wall x=1.5 only
p = np.zeros((n,4),np.float32); p[:,0]=1.5; p[:,1]=rng.uniform(-1.5,1.5,n); p[:,2]=rng.uniform(0,6,n)
q = np.zeros((100000,4),np.float32); q[:,0]=1.5; q[:,1]=rng.uniform(-1.4,1.4,100000); q[:,2]=rng.uniform(0.1,5.9,100000)
for k in (16,100): run("one wall, uniform", p, q, k)
# uniform in volume
p = np.zeros((n,4),np.float32); p[:,:3]=rng.uniform([-1.5,-1.5,0],[1.5,1.5,6],(n,3))
q = np.zeros((100000,4),np.float32); q[:,:3]=rng.uniform([-1.4,-1.4,0.1],[1.4,1.4,5.9],(100000,3))
for k in (16,100): run("volume, uniform", p, q, k)
# traced photons, wall hits inside the box only vs all
sys.argv=[sys.argv[0]]
m = pmb200.PhotonMapper(n_photons=2_000_000)
m.init_random_numbers(); m.set_record_capacity(6_000_000); m.clear_map(); m.trace(0.0, media=False, records=True, no_map=True)
from pmb200 import dist as pd
pos_p, pow_p, _, cnt = m.record_buffers(0)
pos = pd.device_tensor(pos_p, cnt*4, "<f4").cpu().numpy().reshape(cnt,4).copy()
meta = pos[:,3].copy().view(np.uint32); wall = (((meta>>5)&3).astype(np.int32)-1)==1
pw = pos[wall].copy(); pw[:,3]=0
inside = (np.abs(pw[:,0])<=1.5001)&(np.abs(pw[:,1])<=1.5001)&(pw[:,2]>=0)&(pw[:,2]<=6.0001)
print("wall records", len(pw), "inside box", inside.sum())
qi = pw[inside][rng.integers(0, inside.sum(), 100000)].copy()
for k in (16,100): run("traced, all wall records", pw, qi, k)
for k in (16,100): run("traced, inside box only", pw[inside].copy(), qi, k)
