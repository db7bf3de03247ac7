#!/usr/bin/env python
"""tools/ab_knn.py LIB [LIB ...] -- the Mode B gather of config 4 (16 M photons, k = 50, media, 1080p) with several builds of the
library in ONE process (development aid): ms per render and whether the float frames are bit-identical to the first build's."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmb200

n, k, media, W, H = 16777216, 50, True, 1920, 1080
ref = None
for path in sys.argv[1:]:
    pmb200.LIB_PATH, pmb200._lib = os.path.abspath(path), None
    m = pmb200.PhotonMapper(n_photons=n)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    sc = pmb200.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
    m.set_scene(sc); m.init_random_numbers(); m.set_record_capacity(int(2.6 * n))
    m.clear_map(); m.trace(0.0, media=media, records=True, no_map=True)
    m.knn_build(0); m.knn_build(1)
    rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    rgbf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    def render(): m.render_knn(W, H, 0.0, media, k, float("inf"), 1e-4, 1e-2, rgba=rgba, rgbf=rgbf)
    render(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); render(); render(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 2
    if ref is None: ref = rgbf.clone()
    same = bool(torch.equal(rgbf.view(torch.int32), ref.view(torch.int32)))
    print(json.dumps({"lib": os.path.basename(path), "render_knn_ms": ms, "queries_per_s": W * H * 11 / (ms * 1e-3), "frame_bit_identical_to_first": same}), flush=True)
    m.close(); del m, rgba, rgbf
    torch.cuda.empty_cache()
