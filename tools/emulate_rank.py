#!/usr/bin/env python
"""tools/emulate_rank.py -- one rank's share of an N-GPU Mode A frame on ONE GPU, without the exchange (development aid):
photon range n/N, row band H/N, pipelined pm_frame_device; prints device ms/frame and the host time it takes to enqueue a frame."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmb200

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n, W, H = 16777216, 1920, 1080
m = pmb200.PhotonMapper(n_photons=n)
m.set_stream(torch.cuda.current_stream().cuda_stream)
m.set_photon_range(0, n // N)
m.init_random_numbers()
m.set_row_band(0, H // N)
rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
for sms in (0, 148, 140, 132, 124, 116, 108, 100):
    m.set_trace_sms(sms)
    for _ in range(20): m.frame_device(W, H, rgba=rgba, media=True)
    m.sync(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 300
    a.record()
    t0 = time.perf_counter()
    for _ in range(reps): m.frame_device(W, H, rgba=rgba, media=True)
    t1 = time.perf_counter()
    m.sync(); torch.cuda.synchronize()   # m.sync waits for both of the library's streams
    b.record(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("N=%d sms=%3d  wall %.4f ms/frame  host enqueue %.4f ms/frame" % (N, sms, (t2 - t0) * 1e3 / reps, (t1 - t0) * 1e3 / reps), flush=True)
