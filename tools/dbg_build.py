#!/usr/bin/env python
"""tools/dbg_build.py -- per-kernel CUDA-event times of the map build (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmb200
m = pmb200.PhotonMapper(n_photons=1 << 20)
m.set_stream(torch.cuda.current_stream().cuda_stream)
m.init_random_numbers(); m.clear_map(); m.trace(0.0, media=True)
for _ in range(3): m.build_map()
m.enable_timing(True)
for _ in range(50): m.build_map()
print({k: v[0] / v[1] * 1e3 for k, v in m.timings().items() if v[1]}, "us")
