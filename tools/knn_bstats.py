#!/usr/bin/env python
"""tools/knn_bstats.py photons k media -- counters of the batched k-NN renderer (needs a -DPM_KNN_BSTATS build: PMB200_LIB)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pmb200
n, k, media = int(sys.argv[1]), int(sys.argv[2]), bool(int(sys.argv[3]))
W, H = 1920, 1080
m = pmb200.PhotonMapper(n_photons=n)
sc = pmb200.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
m.set_scene(sc); m.init_random_numbers(); m.set_record_capacity(int(2.6 * n))
m.clear_map(); m.trace(0.0, media=media, records=True, no_map=True)
m.knn_build(0)
if media: m.knn_build(1)
rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
L = pmb200.lib()
m.knn_set_batched(True)
out = (C.c_ulonglong * 16)()
L.pm_debug_knn_bstats(out, 1)
m.render_knn(W, H, 0.0, media, k, float("inf"), 1e-4, 1e-2, rgba=rgba); m.sync()
L.pm_debug_knn_bstats(out, 1)
s = [int(x) for x in out]
names = ["batches", "walks", "node visits", "leaves listed", "leaves needed", "boundary overflows", "seeds", "one-by-one lanes", "cyc walks", "cyc selections",
         "cyc seeds", "cyc one-by-one", "cyc radiance", "candidates", "cyc tiles"]
b = max(s[0], 1)
for i, nm in enumerate(names):
    print("%-18s %14d   per batch %10.1f" % (nm, s[i], s[i] / b))
