"""CPU check of the identity the CUDA kernels rely on to drop the double-precision division from
getVoxelCoordinates (PMK:260-267):  trunc(((p + 1.5) / 3.0) * 32) == sign(t) * floor(32|t|/3), t = (double)p + 1.5,
and trunc((p / 6.0) * 32) == sign(p) * floor(32|p|/6).  numpy float64 arithmetic is IEEE, like the device's."""
import numpy as np


def ref_x(p):
    return np.trunc(((p.astype(np.float64) + 1.5) / 3.0) * 32).astype(np.int64)


def ref_z(p):
    return np.trunc((p.astype(np.float64) / 6.0) * 32).astype(np.int64)


def fast(t, c):
    at = np.abs(t)
    a = at * (32.0 / c)
    k = np.trunc(a)
    r = 32.0 * at - c * k
    k = k + (r >= c) - (r < 0.0)
    return np.where(t < 0, -k, k).astype(np.int64)


def fast_x(p):
    return fast(p.astype(np.float64) + 1.5, 3.0)


def fast_z(p):
    return fast(p.astype(np.float64), 6.0)


def _samples():
    rng = np.random.default_rng(7)
    out = [rng.uniform(-2, 8, 2_000_000).astype(np.float32),
           rng.normal(0, 1e3, 500_000).astype(np.float32),
           (rng.uniform(-1, 1, 500_000) * 10.0 ** rng.uniform(-30, 5, 500_000)).astype(np.float32)]
    # floats adjacent to every voxel boundary of both mappings, +-4 ulps
    k = np.arange(-40, 80, dtype=np.float64)
    for edges in (k * 3.0 / 32.0 - 1.5, k * 6.0 / 32.0):
        e = edges.astype(np.float32)
        for _ in range(4):
            out.append(e.copy())
            e = np.nextafter(e, np.float32(np.inf))
        e = edges.astype(np.float32)
        for _ in range(4):
            e = np.nextafter(e, np.float32(-np.inf))
            out.append(e.copy())
    return np.concatenate(out)


def test_voxel_identity():
    p = _samples()
    p = p[np.abs(p) < 3e4]          # the kernels fall back to the literal form beyond |32t/3| >= 2^20
    assert np.array_equal(ref_x(p), fast_x(p))
    assert np.array_equal(ref_z(p), fast_z(p))
