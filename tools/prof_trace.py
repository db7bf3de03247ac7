#!/usr/bin/env python
"""tools/prof_trace.py -- a few fused Mode A traces (16M photons, media on) for an ncu capture (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmb200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16777216
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = pmb200.PhotonMapper(n_photons=n)
m.set_stream(torch.cuda.current_stream().cuda_stream)
m.init_random_numbers()
for _ in range(reps):
    m.clear_map(); m.trace(0.0, media=True); m.build_map()
m.sync()
