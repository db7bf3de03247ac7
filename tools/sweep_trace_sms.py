#!/usr/bin/env python
"""tools/sweep_trace_sms.py -- trace-kernel time against the CTA cap (pm_set_trace_sms) for the photon counts a rank gets at N = 1, 4, 8
(development aid: looks for partition-dependent cliffs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmb200

def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for n in (2097152, 4194304, 16777216):
    m = pmb200.PhotonMapper(n_photons=n)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    m.init_random_numbers()
    row = []
    for sms in range(96, 149, 2 if n < 16777216 else 4):
        m.set_trace_sms(sms)
        row.append("%d:%.4f" % (sms, t(lambda: m.trace(0.0, media=True), 10 if n < 16777216 else 5)))
    print("n=%d media  " % n + " ".join(row), flush=True)
    if n < 16777216:
        row = []
        for sms in range(96, 149, 4):
            m.set_trace_sms(sms)
            row.append("%d:%.4f" % (sms, t(lambda: m.trace(0.0, media=False))))
        print("n=%d surf   " % n + " ".join(row), flush=True)
    m.close()
