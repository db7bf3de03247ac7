import numpy as np


def bits_equal(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def copy_scene(dst_cls, src):
    """Copy a pm_scene between the oracle's and the product's ctypes mirrors (same layout)."""
    import ctypes as C
    d = dst_cls()
    assert C.sizeof(d) == C.sizeof(src)
    C.memmove(C.byref(d), C.byref(src), C.sizeof(src))
    return d


def cfg1_scene(scene):
    """BASELINE config 1 / SURVEY 8(d): single glass sphere + floor; the other walls are pushed past the
    raytrace range (offset 1e9 > 999999.9) which disables them without touching code."""
    scene.n_spheres = 1
    for i, off in enumerate([1e9, -1.5, -1e9, 1e9, 1e9]):
        scene.planes[i][1] = off
    return scene


def grid_from_records(rec):
    """Exact (float64) photon map from a record list: the reference's deposit rules (storePhoton / splatEnergy /
    storeVolumePhoton, PMK:1059-1183) applied in double precision; order-independent up to double rounding."""
    from oracle.oraclelib import Oracle  # noqa: F401  (documentation: records come from the oracle or the product)
    g = np.zeros((32, 32, 32, 3), np.float64)

    def vox(p):
        v = np.empty(3, np.int64)
        v[0] = int(((float(p[0]) + 1.5) / 3.0) * 32)
        v[1] = int(((float(p[1]) + 1.5) / 3.0) * 32)
        v[2] = int((float(p[2]) / 6.0) * 32)
        return np.clip(v, 0, 31)

    def win(v):
        return max(v - 3, 0), min(v + 3, 32)

    for r in rec:
        v = vox(r["loc"])
        e = r["energy"].astype(np.float64)
        if r["kind"] == 1:
            g[v[0], v[1], v[2]] += e
            continue
        if r["type"] != 1:
            continue
        g[v[0], v[1], v[2]] += e
        rid = int(r["id"])
        rng = [range(*win(v[0])), range(*win(v[1])), range(*win(v[2]))]
        if rid in (0, 2):
            rng[0] = [31 if rid == 0 else 0]
        elif rid in (1, 3):
            rng[1] = [0 if rid == 1 else 31]
        elif rid == 4:
            rng[2] = [31]
        else:
            continue
        e05 = (r["energy"] * np.float32(0.05)).astype(np.float64)
        for i in rng[0]:
            for j in rng[1]:
                for k in rng[2]:
                    if (i, j, k) == (v[0], v[1], v[2]):
                        continue
                    d = np.float32(np.sqrt(np.float32((v[0] - i) ** 2 + (v[1] - j) ** 2 + (v[2] - k) ** 2)))
                    g[i, j, k] += e05 * np.float64(np.float32(1.0) / d)
    return g
