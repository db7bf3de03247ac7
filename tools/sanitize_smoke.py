#!/usr/bin/env python
"""tools/sanitize_smoke.py -- small end-to-end run of every kernel, meant to be wrapped in compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
(the reference has data races by design -- non-atomic += on the grid, shared MWC state; this build must have none)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmb200

n, w, h = 20000, 96, 64
m = pmb200.PhotonMapper(n_photons=n)
sc = pmb200.default_scene(sz_img=64); sc.cam_ox = -16.0
m.set_scene(sc)
m.init_random_numbers()
m.set_record_capacity(16 * n)
m.clear_map()
m.trace(0.3, media=True, records=True)
m.build_map()
u8, f32 = m.render(w, h, 0.3, True, True)
m.knn_build(0); m.knn_build(1)
rgba = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda"); rgbf = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
m.render_knn(w, h, 0.3, True, 50, float("inf"), 1e-4, 1e-2, rgba=rgba, rgbf=rgbf)
m.render_knn(w, h, 0.3, False, 100, 0.01, 1e-4, 1e-2, rgba=rgba, rgbf=rgbf, y0=1, y1=h, y_step=3)
q = torch.rand((64, 4), device="cuda") * 3 - 1.5
idx = torch.empty((64, 20), dtype=torch.int32, device="cuda"); d2 = torch.empty((64, 20), device="cuda"); cnt = torch.empty(64, dtype=torch.int32, device="cuda")
m.knn_query(0, q, 64, 20, float("inf"), idx, d2, cnt)
rgb = torch.empty((64, 4), device="cuda")
m.knn_radiance_cone(q, 64, 100, 0.7, 50.0, rgb)
m.init_random_numbers_philox(7)
m.clear_map(); m.trace(0.0, media=True); m.build_map()
m.trace(0.0, media=True, split=True)                      # the two-launch form of the trace
m.set_volume_warps(3); m.trace(0.0, media=True)           # fused, another warp split
m.set_volume_warps(7)
m.trace(0.0, media=True, one_phase=True); m.trace(0.0, media=True, exact_medium=True)   # the cross-check forms of the Mode A trace
# the two-phase walk with queues that fill and drain: many photons on ONE CTA (8 000 per warp: ~240 queued for the state machine)
m2 = pmb200.PhotonMapper(n_photons=200000)
m2.init_random_numbers(); m2.set_trace_sms(1)
m2.clear_map(); m2.trace(0.0, media=True); m2.trace(1.3, media=False); m2.build_map()
m2.sync(); m2.close()
pin = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
tk = [m.frame_async(w, h, pin[i & 1], t=0.1 * i, emit=True, interp=True, media=True) for i in range(2)]   # pipelined frames
for t_ in tk:
    m.frame_wait(t_)
m.sync()
print("sanitize smoke ok: map sum %.4f, frame mean %.5f, knn frame mean %.5f" % (float(m.get_map().sum()), float(f32[..., :3].mean()), float(rgbf[..., :3].mean())))
m.close()
