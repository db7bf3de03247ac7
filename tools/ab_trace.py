#!/usr/bin/env python
"""tools/ab_trace.py LIB... -- CUDA-event time of the fused Mode A trace (16M and 2M photons, media on) for several builds of the
library (PMB200_LIB), each in its own process (development aid)."""
import sys, os, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import torch, pmb200
    m = pmb200.PhotonMapper(n_photons=16777216)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    m.init_random_numbers()
    if os.environ.get("AB_VOL_WARPS"): m.set_volume_warps(int(os.environ["AB_VOL_WARPS"]))
    def t(fn, reps=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    out = {}
    for tag, (f, l) in (("16M", (0, 16777216)), ("2M", (0, 2097152))):
        m.set_photon_range(f, l)
        out[tag + " media"] = round(t(lambda: m.trace(0.0, media=True)), 4)
        out[tag + " surface"] = round(t(lambda: m.trace(0.0, media=False)), 4)
    print(json.dumps(out))
else:
    for lib in sys.argv[1:] or [""]:
        env = dict(os.environ)
        if lib: env["PMB200_LIB"] = os.path.abspath(lib)
        r = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True)
        print("%-34s %s" % (lib or "default", r.stdout.strip() or r.stderr.strip()[-300:]))
