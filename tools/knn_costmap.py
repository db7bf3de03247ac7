#!/usr/bin/env python
"""tools/knn_costmap.py -- per-pixel cycle count of the Mode B renderer (library built with -DPM_KNN_STATS), saved as PNG + stats."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmb200, numpy as np
from PIL import Image
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4194304
k = int(sys.argv[2]) if len(sys.argv) > 2 else 50
media = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
W, H = 1920, 1080
m = pmb200.PhotonMapper(n_photons=n)
sc = pmb200.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0; m.set_scene(sc)
m.init_random_numbers(); m.set_record_capacity(int(2.6 * n)); m.clear_map(); m.trace(0.0, media=media, records=True, no_map=True)
m.knn_build(0)
if media: m.knn_build(1)
rgbf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
m.render_knn(W, H, 0.0, media, k, float("inf"), 1e-4, 1e-2, rgbf=rgbf); m.sync()
c = rgbf[..., 3].cpu().numpy()
print("cycles/pixel: mean %.0f median %.0f p99 %.0f p99.9 %.0f max %.0f; sum of top 0.1%% = %.1f%% of all" % (
    c.mean(), np.median(c), np.percentile(c, 99), np.percentile(c, 99.9), c.max(), 100 * np.sort(c.ravel())[-c.size // 1000:].sum() / c.sum()))
ys, xs = np.unravel_index(np.argsort(c.ravel())[-8:], c.shape)
print("slowest pixels (x,y,cycles):", [(int(x), int(y), int(c[y, x])) for x, y in zip(xs, ys)])
img = np.clip(np.log10(np.maximum(c, 1)) / np.log10(c.max()) * 255, 0, 255).astype(np.uint8)
Image.fromarray(img).resize((960, 540)).save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "knn_costmap.png"))
