#!/usr/bin/env python
"""tests/golden/make_golden.py -- mints the golden vectors under tests/golden/ from the REFERENCE'S OWN code.

The reference ships no tests or fixtures, so the vectors are generated here by running its device
routines, compiled unmodified as host C++ (oracle/_ref/libpmref_host.so, built by oracle/build_ref.sh from
/root/reference/photonMappingKernel.cu:1-1521), sequentially in photon-index / pixel-index order.
Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The committed .npz files are what tests/test_oracle_golden.py pins oracle/pm_oracle.c against on any machine.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refhost import RefHost  # noqa: E402

N = 4096
W = H = 48
CFG1_PLANES = np.array([[0, 1e9], [1, -1.5], [0, -1e9], [1, 1e9], [2, 1e9]], np.float32)


def main():
    r = RefHost()
    out = {}
    # RNG known-answer vectors: raw MWC draws, randFloat(1.0), the table in device (x,y,z) order, state after it
    r.set_rng(6548, 316)
    out["mwc_u32"] = np.array([r.get_random() for _ in range(16)], np.uint32)
    r.set_rng(6548, 316)
    out["mwc_randfloat"] = np.array([r.rand_float(1.0) for _ in range(16)], np.float32)
    r.set_rng(6548, 316)
    r.init_table(N)
    table = r.get_table(N)
    state = r.get_rng()
    out["table_head"] = table[:32].copy()
    out["table_sha256"] = np.frombuffer(hashlib.sha256(table.tobytes()).digest(), np.uint8)
    out["state_after_table"] = np.array(state, np.uint32)

    cases = []
    for scene in ("default", "cfg1"):
        for t in (0.0, 0.7):
            for media in (0, 1):
                r.reset()
                if scene == "cfg1":
                    r.set_scene(nr_objects=(1, 5), planes=CFG1_PLANES, sz_img=W)
                else:
                    r.set_scene(sz_img=W)
                r.set_table(table); r.set_rng(*state); r.clear_grid()
                rec = r.emit(0, N, t, False, bool(media), max_records=32 * N)
                grid = r.get_grid()
                key = "%s_t%.1f_m%d" % (scene, t, media)
                cases.append(key)
                nz = np.flatnonzero(np.abs(grid).reshape(-1, 3).sum(1) != 0)
                out[key + "_grid_idx"] = nz.astype(np.int32)
                out[key + "_grid_val"] = grid.reshape(-1, 3)[nz]
                out[key + "_rec_count"] = np.array([len(rec)], np.int64)
                out[key + "_rec_sha256"] = np.frombuffer(hashlib.sha256(rec.tobytes()).digest(), np.uint8)
                out[key + "_rec_head"] = rec[:48].copy()
                out[key + "_rng_after"] = np.array(r.get_rng(), np.uint32)
                out[key + "_spheres"] = r.get_scene()["spheres"]
                for interp in (0, 1):
                    img = r.render_f32(W, H, t, bool(interp), bool(media))
                    out[key + "_img_i%d" % interp] = img
                if media == 0 and scene == "default" and t == 0.0:
                    out[key + "_u8"] = r.render_u8(W, H, t, False, False)   # host cast: wraps where the float value is negative
    out["cases"] = np.array(cases)
    out["meta"] = np.array([N, W, H], np.int64)
    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(cases), "cases")


if __name__ == "__main__":
    main()
