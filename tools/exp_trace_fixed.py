#!/usr/bin/env python
"""tools/exp_trace_fixed.py -- where the fixed (non-scaling) cost of the fused trace goes: CUDA-event time vs photon count,
and the per-CTA / per-warp %globaltimer stamps of pm_trace_profile (development aid, one GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pmb200

NP = 16777216
m = pmb200.PhotonMapper(n_photons=NP)
m.set_stream(torch.cuda.current_stream().cuda_stream)
m.init_random_numbers()

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for first, last, tag in ((0, NP, "all 16M"), (0, NP // 2, "rank 0 of 2"), (0, NP // 4, "rank 0 of 4"), (0, NP // 8, "rank 0 of 8"),
                         (NP // 8, NP // 4, "rank 1 of 8"), (7 * NP // 8, NP, "rank 7 of 8"), (0, NP // 16, "1/16"), (NP // 2, NP // 2 + 4096, "4096 photons")):
    m.set_photon_range(first, last)
    ms = t(lambda: m.trace(0.0, media=True))
    ms_nm = t(lambda: m.trace(0.0, media=True, no_map=True))
    ms_s = t(lambda: m.trace(0.0, media=False))
    m.trace_profile(True)
    m.trace(0.0, media=True); m.sync()
    p = m.get_trace_profile().astype(np.int64)
    m.trace_profile(False)
    p = p[p[:, 0] > 0]
    t0 = p[:, 0].min()
    init, walk, flush = (p[:, 1] - p[:, 0]) / 1e3, (p[:, 2] - p[:, 1]) / 1e3, (p[:, 3] - p[:, 2]) / 1e3
    wd = (p[:, 8:40] - p[:, 1:2]) / 1e3    # warp-done times relative to the CTA's walk start, us
    end = (p[:, 3].max() - t0) / 1e3
    print("%-14s media %.4f ms | no-map %.4f | surface-only %.4f || kernel span %.1f us: start skew %.1f, init %.1f, walk mean %.1f max %.1f (CTA %d), "
          "flush mean %.1f max %.1f | surface warps done: mean %.1f, p95 %.1f, max %.1f; CTA0 warp6 %.1f; medium warps done mean %.1f max %.1f"
          % (tag, ms, ms_nm, ms_s, end, (p[:, 0].max() - t0) / 1e3, init.mean(), walk.mean(), walk.max(), int(walk.argmax()), flush.mean(), flush.max(),
             wd[:, 6:].mean(), np.percentile(wd[:, 6:], 95), wd[:, 6:].max(), wd[0, 6], wd[:, :6].mean(), wd[:, :6].max()))
m.set_photon_range(0, NP)
print("clear   %.4f ms" % t(lambda: m.clear_map()))
print("build   %.4f ms" % t(lambda: m.build_map()))
W, H = 1920, 1080
sc = pmb200.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
m.set_scene(sc)
rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
for d in (1, 2, 4, 8):
    print("render 1/%d of the rows, uchar4 only  %.4f ms" % (d, t(lambda: m.render_device(W, H, 0.0, False, True, rgba=rgba, y0=0, y1=H // d))))
