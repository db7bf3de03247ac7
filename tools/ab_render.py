#!/usr/bin/env python
"""tools/ab_render.py LIB [LIB ...] -- the Mode A render kernel (1080p, media on, 16 M-photon map) with several builds of the library in
one process (development aid): ms per render and whether the frames are bit-identical to the first build's."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmb200
n, W, H = 16777216, 1920, 1080
ref = None
for path in sys.argv[1:]:
    pmb200.LIB_PATH, pmb200._lib = os.path.abspath(path), None
    m = pmb200.PhotonMapper(n_photons=n)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    sc = pmb200.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
    m.set_scene(sc); m.set_energy_scale(10000.0 / n); m.init_random_numbers()
    m.clear_map(); m.trace(0.0, media=True); m.build_map()
    rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    rgbf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    out = {"lib": os.path.basename(path)}
    for interp in (False, True):
        def render(): m.render_device(W, H, 0.0, interp, True, rgba=rgba, rgbf=rgbf)
        for _ in range(3): render()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): render()
        b.record(); torch.cuda.synchronize()
        out["render_ms_interp%d" % interp] = a.elapsed_time(b) / 20
        key = "f%d" % interp
        if ref is None or key not in ref: ref = dict(ref or {}, **{key: (rgbf.clone(), rgba.clone())})
        out["bit_identical_interp%d" % interp] = bool(torch.equal(rgbf.view(torch.int32), ref[key][0].view(torch.int32)) and torch.equal(rgba, ref[key][1]))
    print(json.dumps(out), flush=True)
    m.close(); del m
