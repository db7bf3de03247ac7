#!/usr/bin/env python
"""tools/show_bench.py -- the few numbers of bench.py JSON lines that matter while iterating."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "unreadable:", ex); continue
    print(f, "N=%d value %.4f ms  latency %.4f  e2e %.4f" % (d["n_gpus"], d["value"], d.get("frame_latency_ms", 0), d["e2e"]["value"]))
    print("   stages", {k: round(v, 4) for k, v in d["stages_ms"].items() if isinstance(v, float)})
    print("   kernels", {k: round(v["avg_ms"], 4) for k, v in d["kernels"].items()})
    for k in ("config5_progressive", "mode_b_config4", "peer_exchange_ok"):
        if k in d: print("   ", k, d[k])
