#!/usr/bin/env python
"""tools/knn_stats.py -- traversal statistics of the Mode B renderer (needs a library built with -DPM_KNN_STATS):
    make -C cuda-photon-mapper_b200 PM_EXTRA_NVCCFLAGS=-DPM_KNN_STATS && python tools/knn_stats.py [photons] [k] [media]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmb200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4194304
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
media = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
W, H = 1920, 1080
m = pmb200.PhotonMapper(n_photons=n)
sc = pmb200.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0; m.set_scene(sc)
m.init_random_numbers(); m.set_record_capacity(int(2.6 * n)); m.clear_map(); m.trace(0.0, media=media, records=True, no_map=True)
m.knn_build(0)
if media: m.knn_build(1)
rgbf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
L = pmb200.lib(); out = (C.c_ulonglong * 8)()
L.pm_debug_knn_stats(out, 1)
m.render_knn(W, H, 0.0, media, k, float("inf"), 1e-4, 1e-2, rgbf=rgbf); m.sync()
L.pm_debug_knn_stats(out, 1)
q = max(out[4], 1)
print("k %d media %d: searches %d (%.2f per query issued), leaves/search %.1f, node steps %.1f, passed %.1f, merges %.2f, hinted %.3f, retries %.3f"
      % (k, media, out[4], out[4] / (W * H * (11 if media else 1)), out[0] / q, out[1] / q, out[2] / q, out[3] / q, out[6] / q, out[5] / q))
