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


def clamped_x(p):
    """voxel_x_clamped (csrc/pm_math.cuh): FP32 + integer arithmetic only"""
    u = np.float32(32.0) * p                      # exact
    u = (u.astype(np.float64) + 2.0 ** -48).astype(np.float32)   # fma(32, p, 2^-48): one rounding (the f64 sum of two f32 is exact
    #                                                              unless their exponents are > 29 apart, where RN is u either way)
    u = np.clip(u, np.float32(-48.0), np.float32(48.0))
    k = ((np.floor(u).astype(np.int64) + 48) * 43691) >> 17
    return np.minimum(k, 31)


def clamped_z(p):
    u = np.clip(np.float32(16.0) * p, np.float32(0.0), np.float32(96.0))
    k = (np.floor(u).astype(np.int64) * 43691) >> 17
    return np.minimum(k, 31)


def rng_exponents():
    return np.arange(20, 149, dtype=np.float64)


def test_clamped_voxel_matches_literal_form():
    p = _samples()
    big = np.float32([1e9, -1e9, 3e38, -3e38, 1e-40, -1e-40, 0.0, -0.0, -2.0 ** -53, -2.0 ** -52, -2.0 ** -54, 2.0 ** -53,
                      -1.5, np.nextafter(np.float32(-1.5), np.float32(0)), np.nextafter(np.float32(-1.5), np.float32(-2))])
    tiny = (-(2.0 ** -rng_exponents())).astype(np.float32)
    big = np.concatenate([big, tiny, -tiny])
    p = np.concatenate([p, big])
    assert p.dtype == np.float32

    def sat(f, p):   # the literal double form with the device's saturating double -> int conversion (cvt.rzi.s32.f64)
        return np.clip(np.trunc(f(p)), -2.0 ** 31, 2.0 ** 31 - 1).astype(np.int64)
    lit_x = lambda p: ((p.astype(np.float64) + 1.5) / 3.0) * 32
    lit_z = lambda p: (p.astype(np.float64) / 6.0) * 32
    assert np.array_equal(np.clip(sat(lit_x, p), 0, 31), clamped_x(p))
    assert np.array_equal(np.clip(sat(lit_z, p), 0, 31), clamped_z(p))
    # every float in the two voxel ranges' neighbourhood that is a multiple of 2^-12: all boundaries, both sides
    q = (np.arange(-3 * 4096, 8 * 4096, dtype=np.int64) / 4096.0).astype(np.float32)
    assert np.array_equal(np.clip(ref_x(q), 0, 31), clamped_x(q))
    assert np.array_equal(np.clip(ref_z(q), 0, 31), clamped_z(q))


def test_div65535_constant_reciprocal_is_correctly_rounded():
    """strided subset of oracle/check_div65535.c (the exhaustive run, stride 1, reports 0 mismatches of 2^32)"""
    import os, subprocess
    exe = os.path.join(os.path.dirname(__file__), "..", "oracle", "check_div65535")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "check_div65535"])
    out = subprocess.run([exe, "61"], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == 0 and int(out[1]) > 70_000_000


def test_voxel_wide_one_conversion_form_strided():
    """strided subset of oracle/check_voxel_wide.c: the render kernel's one-conversion voxel_x_wide / voxel_z_wide against the
    two-conversion form and the reference's literal double form (the exhaustive run, stride 1, reports 0 0 of 2^32 patterns)"""
    import os, subprocess
    exe = os.path.join(os.path.dirname(__file__), "..", "oracle", "check_voxel_wide")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "check_voxel_wide"])
    out = subprocess.run([exe, "61"], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == 0 and int(out[1]) == 0 and int(out[2]) > 70_000_000


def _div3_trunc(n):
    return np.where(n >= 0, (n * 43691) >> 17, -((-n * 43691) >> 17))


def wide_x(p):
    """voxel_x_wide (csrc/pm_math.cuh): clamp(voxel, -3, 36) in FP32 + integer arithmetic, NaN -> 0"""
    with np.errstate(invalid="ignore"):
        u = np.float32(32.0) * p
        u = (u.astype(np.float64) + 2.0 ** -48).astype(np.float32)
        u = np.where(np.isnan(u), np.float32(-48.0), np.clip(u, np.float32(-57.0), np.float32(60.0)))
        n = np.where(u >= -48.0, np.floor(u), np.ceil(u)).astype(np.int64) + 48
    return _div3_trunc(n)


def wide_z(p):
    with np.errstate(invalid="ignore"):
        u = np.float32(16.0) * p
        u = np.where(np.isnan(u), np.float32(0.0), np.clip(u, np.float32(-9.0), np.float32(108.0)))
        n = np.where(u >= 0.0, np.floor(u), np.ceil(u)).astype(np.int64)
    return _div3_trunc(n)


def test_wide_voxel_matches_literal_form():
    p = _samples()
    tiny = (-(2.0 ** -rng_exponents())).astype(np.float32)
    extra = np.float32([1e9, -1e9, 3e38, -3e38, np.inf, -np.inf, np.nan, 0.0, -0.0, -1.5, -1.59375, -1.78125, 2.0, 1.875, 1.96875,
                        -0.5625, -1.6875, 6.75, 6.5625])
    q = (np.arange(-4 * 4096, 9 * 4096, dtype=np.int64) / 4096.0).astype(np.float32)
    p = np.concatenate([p, tiny, -tiny, extra, q])

    def sat(v):   # cvt.rzi.s32.f64: saturating, NaN -> 0
        with np.errstate(invalid="ignore"):
            return np.where(np.isnan(v), 0.0, np.clip(np.trunc(v), -2.0 ** 31, 2.0 ** 31 - 1)).astype(np.int64)
    with np.errstate(invalid="ignore", over="ignore"):
        lit_x = sat(((p.astype(np.float64) + 1.5) / 3.0) * 32)
        lit_z = sat((p.astype(np.float64) / 6.0) * 32)
    assert np.array_equal(np.clip(lit_x, -3, 36), wide_x(p))
    assert np.array_equal(np.clip(lit_z, -3, 36), wide_z(p))
