#!/usr/bin/env python
"""tests/golden/make_fullsize_golden.py -- photon maps of BASELINE configs 2 and 4 from the SEQUENTIAL oracle (oracle/pm_oracle.c,
pinned bit-exact to the reference's own routines by tests/test_oracle_vs_ref.py), at their stated sizes:

    config 2: default participating-media scene, 1 048 576 photons      config 4: the same scene, 16 777 216 photons

Two maps per config: `_map` is the reference's literal result, every deposit added to a float voxel in photon order -- at 16M
photons those FP32 sums have grown so large that late deposits are rounded away; `_map_exact` sums the SAME FP32 deposit values in
float64 (the oracle's shadow grid), i.e. what the deposits add up to without the rounding and saturation of the float voxels.  One CPU run (about 3 s and 35 s); stored in tests/golden/fullsize_maps.npz (with the MWC state the medium walk leaves
behind) so that the GPU tests need not re-run it.
Run from the repo root:  python tests/golden/make_fullsize_golden.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oraclelib

orc = oraclelib.Oracle()
out = {}
for name, n in (("config2", 1 << 20), ("config4", 1 << 24)):
    sc = orc.default_scene()
    t0 = time.time()
    table, st = orc.mwc_table(n)
    acc = np.zeros((32, 32, 32, 3), np.float64)
    grid, _, st2 = orc.emit(sc, table, 0, n, 0.0, True, rng=st, shadow64=acc)
    print(name, n, "photons: %.1f s" % (time.time() - t0), "sum", grid.sum(dtype=np.float64), "state", st2)
    out[name + "_map"] = grid
    out[name + "_state"] = np.array(st2, np.uint32)
    out[name + "_map_exact"] = acc.astype(np.float32)
    d = np.abs(grid.astype(np.float64) - acc)
    print("   sequential FP32 vs exact sum: max |diff| / max |map| = %.3e, rel-L1 = %.3e" % (d.max() / np.abs(acc).max(), d.sum() / np.abs(acc).sum()))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fullsize_maps.npz"), **out)
