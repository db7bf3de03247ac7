#!/usr/bin/env python
"""tools/time_knn.py -- CUDA-event timing of the Mode B stages on one GPU (development aid / bench helper).
    python tools/time_knn.py [photons] [k] [media 0|1]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmb200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4194304
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
media = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
W, H = 1920, 1080
m = pmb200.PhotonMapper(n_photons=n)
m.set_stream(torch.cuda.current_stream().cuda_stream)
sc = pmb200.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
m.set_scene(sc)
m.init_random_numbers()
m.set_record_capacity(int(2.6 * n))
def ev(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def trace():
    m.clear_map(); m.trace(0.0, media=media, records=True, no_map=True)
out = {"photons": n, "k": k, "media": media}
out["trace_records_ms"] = ev(trace)
out["build_surface_ms"] = ev(lambda: m.knn_build(0))
ns, lv = m.knn_size(0)
out["surface_points"], out["surface_levels"] = ns, lv
# algorithmic bytes of the build: 16 read pos + 8 write key,idx + 4 passes x (4 hist read + 16 scatter r/w) + 32 permute + boxes
out["build_surface_GBps_alg"] = (m.record_buffers(0)[3] * (16 + 8 + 4 * 20) + ns * 32) / (out["build_surface_ms"] * 1e-3) / 1e9
if media:
    out["build_volume_ms"] = ev(lambda: m.knn_build(1))
    out["volume_points"] = m.knn_size(1)[0]
rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
rgbf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
nq = W * H * (11 if media else 1)
if len(sys.argv) > 4 and sys.argv[4] == "both":      # round 1's warp-per-pixel renderer next to the batched one
    m.knn_set_batched(False)
    ms0 = ev(lambda: m.render_knn(W, H, 0.0, media, k, float("inf"), 1e-4, 1e-2, rgba=rgba, rgbf=rgbf), reps=2)
    out["render_knn_warp_per_pixel_ms"] = ms0
    ref = rgbf.clone()
    m.knn_set_batched(True)
ms = ev(lambda: m.render_knn(W, H, 0.0, media, k, float("inf"), 1e-4, 1e-2, rgba=rgba, rgbf=rgbf), reps=2)
if len(sys.argv) > 4 and sys.argv[4] == "both":
    out["max_rel_diff_between_renderers"] = float(((rgbf - ref).abs().max() / ref.abs().max()).item())
out["render_knn_ms"] = ms
out["queries_per_s"] = nq / (ms * 1e-3)
out["frame_ms"] = out["trace_records_ms"] + out["build_surface_ms"] + out.get("build_volume_ms", 0.0) + ms
print(json.dumps(out))
