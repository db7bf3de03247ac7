#!/usr/bin/env python
"""tools/time_trace.py -- CUDA-event timing of pm_trace variants on one GPU (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmb200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16777216
m = pmb200.PhotonMapper(n_photons=n)
m.set_stream(torch.cuda.current_stream().cuda_stream)
m.init_random_numbers()
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("surface only, map      %.3f ms" % t(lambda: m.trace(0.0, media=False)))
print("surface only, no map   %.3f ms" % t(lambda: m.trace(0.0, media=False, no_map=True)))
print("media, map, split      %.3f ms" % t(lambda: m.trace(0.0, media=True, split=True)))
for w in (4, 5, 6, 7, 8):
    m.set_volume_warps(w)
    print("media, map, fused w=%d  %.3f ms" % (w, t(lambda: m.trace(0.0, media=True))))
m.set_volume_warps(4)
print("media, no map          %.3f ms" % t(lambda: m.trace(0.0, media=True, no_map=True)))
print("clear                  %.3f ms" % t(lambda: m.clear_map()))
print("build                  %.3f ms" % t(lambda: m.build_map()))
