#!/usr/bin/env python
"""tools/dbg_group.py -- same-device group frames with a short peer timeout (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pmb200
media = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 2
N, W, H = 200000, 320, 200
g = pmb200.PhotonGroup([0] * nr, n_photons=N)
for r in range(nr):
    g.rank(r).peer_set_timeout(1.0)
g.init_random_numbers()
u8 = np.zeros((H, W, 4), np.uint8)
for f in range(4):
    t0 = time.time()
    g.frame(W, H, 0.0, True, False, bool(media), out_u8=u8)
    st = []
    for r in range(nr):
        try:
            g.rank(r).peer_status(); st.append("ok")
        except pmb200.PmError as ex:
            st.append("TIMEOUT")
    print("media %d ranks %d frame %d: %.3f s %s" % (media, nr, f, time.time() - t0, st), flush=True)
g.close()
