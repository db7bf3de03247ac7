#!/usr/bin/env python
"""tools/trace_timeline.py -- where a short trace launch spends its time (pm_trace_profile stamps), for a rank's share at N GPUs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pmb200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = 16777216
m = pmb200.PhotonMapper(n_photons=n)
m.set_stream(torch.cuda.current_stream().cuda_stream)
m.set_photon_range(0, n // N)
m.init_random_numbers()
for _ in range(3):
    m.clear_map(); m.trace(0.0, media=True)
m.trace_profile(True)
m.clear_map(); m.trace(0.0, media=True); m.sync()
p = m.get_trace_profile().astype(np.int64)
p = p[p[:, 0] > 0]
t0 = p[:, 0].min()
us = lambda x: x / 1e3
print("CTAs %d" % len(p))
print("start spread          %.1f us" % us(p[:, 0].max() - t0))
print("smem zeroed (mean)    %.1f us after own start" % us((p[:, 1] - p[:, 0]).mean()))
w = p[:, 8:40]
wv = np.where(w > 0, w, 0)
print("first warp done       mean %.1f us  min %.1f" % (us((np.where(w > 0, w, 1 << 62).min(1) - t0).mean()), us(np.where(w > 0, w, 1 << 62).min() - t0)))
for i in range(32):
    col = w[:, i]
    if (col > 0).any(): print("  warp %2d done  mean %.1f us  max %.1f" % (i, us((col[col > 0] - t0).mean()), us(col.max() - t0)))
print("all warps done [2]    mean %.1f us  max %.1f us" % (us((p[:, 2] - t0).mean()), us(p[:, 2].max() - t0)))
print("flushed [3]           mean %.1f us  max %.1f us   flush alone %.1f us" % (us((p[:, 3] - t0).mean()), us(p[:, 3].max() - t0), us((p[:, 3] - p[:, 2]).mean())))
