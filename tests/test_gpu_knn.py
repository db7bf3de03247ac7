"""GPU parity tests of the Mode B pipeline (Morton keys, radix sort, 32-wide LBVH, k-nearest-photon search) through the
C-ABI, against the brute-force oracle (oracle/knn_oracle.c).  Bars: sorted keys / permutation bit-exact (stable sort);
k-NN index sets, distances and counts bit-exact; radiance estimate within 2e-5 relative (the product sums the k
powers in FP32 lane order, the oracle in double)."""
import numpy as np
import pytest

from tests.util import copy_scene

pytestmark = pytest.mark.gpu


def _points(n, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        p = rng.uniform([-1.5, -1.5, 0], [1.5, 1.5, 6], (n, 3))
    elif kind == "walls":       # photons live on 2-D manifolds
        p = rng.uniform([-1.5, -1.5, 0], [1.5, 1.5, 6], (n, 3))
        ax = rng.integers(0, 3, n)
        p[np.arange(n), ax] = np.where(ax == 2, 6.0, np.where(rng.random(n) < 0.5, -1.5, 1.5))
    elif kind == "lattice":     # many exact distance ties
        p = rng.integers(0, 8, (n, 3)) * 0.25 + np.array([-1.0, -1.0, 1.0])
    elif kind == "same":
        p = np.tile(np.array([[0.25, -0.5, 3.0]]), (n, 1))
    elif kind == "outside":     # beyond the Morton box: clamped keys, still exact
        p = rng.normal(0, 30, (n, 3))
    out = np.zeros((n, 4), np.float32)
    out[:, :3] = p
    return out


def _build(pm, which, pts, power=None, curve=1):
    import torch
    m = pm.PhotonMapper(n_photons=16)
    m.knn_set_curve(curve)
    tp = torch.from_numpy(pts).cuda()
    tw = torch.from_numpy(power).cuda() if power is not None else None
    m.knn_build_points(which, tp, tw, pts.shape[0])
    m.sync()
    return m, tp, tw


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n,kind", [(1, "uniform"), (33, "uniform"), (4097, "walls"), (200000, "uniform"), (50000, "outside")])
def test_sort_keys_and_stable_sort(pm, oracle, n, kind, curve):
    """Morton (curve 0) and Hilbert (curve 1, default) keys bit-exact; the radix sort equals a stable sort."""
    pts = _points(n, kind, 1)
    m, _, _ = _build(pm, 0, pts, curve=curve)
    keys, perm = m.knn_sorted(0, n)
    okeys = oracle.hilbert30(pts) if curve else oracle.morton30(pts)
    operm = oracle.stable_sort_perm(okeys)
    assert np.array_equal(perm, operm)
    assert np.array_equal(keys, okeys[operm])
    assert m.knn_size(0)[0] == n
    m.close()


def test_tree_boxes_are_tight_and_nested(pm, oracle):
    n = 70001
    pts = _points(n, "walls", 2)
    m, _, _ = _build(pm, 0, pts)
    _, perm = m.knn_sorted(0, n)
    sp = pts[perm][:, :3]
    n_pts, levels = m.knn_size(0)
    assert levels == 3          # 2188 leaves -> 69 nodes -> 3 nodes (<= 32: top level)
    prev = None
    for lv in range(levels):
        b = m.knn_level(0, lv)
        cnt = b.shape[1]
        if lv == 0:
            assert cnt == (n + 31) // 32
            for e in (0, 1, cnt // 2, cnt - 1):
                chunk = sp[32 * e: 32 * e + 32]
                assert np.array_equal(b[:3, e], chunk.min(0)) and np.array_equal(b[3:, e], chunk.max(0))
        else:
            assert cnt == (prev.shape[1] + 31) // 32
            for e in range(cnt):
                ch = prev[:, 32 * e: 32 * e + 32]
                assert np.array_equal(b[:3, e], ch[:3].min(1)) and np.array_equal(b[3:, e], ch[3:].max(1))
        prev = b
    m.close()


def _check_knn(pm, oracle, pts, queries, k, max_r2=np.inf, curve=1):
    import torch
    m, tp, _ = _build(pm, 0, pts, curve=curve)
    nq = queries.shape[0]
    tq = torch.from_numpy(queries).cuda()
    idx = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    d2 = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
    m.knn_query(0, tq, nq, k, float(max_r2), idx, d2, cnt)
    m.sync()
    oidx, od2, ocnt = oracle.knn_bruteforce(pts, queries, k, max_r2)
    assert np.array_equal(cnt.cpu().numpy(), ocnt)
    assert np.array_equal(idx.cpu().numpy(), oidx)
    assert d2.cpu().numpy().tobytes() == od2.tobytes()
    m.close()


@pytest.mark.parametrize("n,kind,k", [(1, "uniform", 1), (20, "uniform", 50), (32, "same", 8), (33, "uniform", 33),
                                      (5000, "lattice", 50), (5000, "lattice", 100), (60000, "walls", 50),
                                      (60000, "uniform", 100), (60000, "uniform", 128), (3000, "outside", 64),
                                      (150000, "walls", 17)])
def test_knn_bit_exact(pm, oracle, n, kind, k):
    pts = _points(n, kind, 3)
    rng = np.random.default_rng(4)
    q = _points(600, "uniform", 5)
    q[:50, :3] = pts[rng.integers(0, n, 50), :3]          # queries exactly on photons (d2 = 0 ties)
    q[50:60, :3] = rng.normal(0, 100, (10, 3))            # far outside
    q[60, :3] = np.nan                                    # NaN query finds nothing
    _check_knn(pm, oracle, pts, q, k, curve=(n + k) % 2)  # both sort curves get exercised


@pytest.mark.parametrize("max_r2", [0.0, 1e-3, 0.05])
def test_knn_radius_limited(pm, oracle, max_r2):
    pts = _points(40000, "walls", 6)
    q = _points(500, "walls", 7)
    q[:20, :3] = pts[:20, :3]
    _check_knn(pm, oracle, pts, q, 50, max_r2)


@pytest.mark.parametrize("media", [False, True])
def test_knn_on_traced_photons(pm, oracle, media):
    """Whole Mode B front end: trace with records -> build both maps -> k-NN + radiance estimate at wall points."""
    import torch
    from pmb200 import dist as pd
    n = 30000
    osc = oracle.default_scene()
    m = pm.PhotonMapper(n_photons=n)
    m.set_scene(copy_scene(pm.Scene, osc))
    m.init_random_numbers()
    m.set_record_capacity(16 * n)
    m.clear_map()
    m.trace(0.0, media=media, records=True, no_map=True)
    rng = np.random.default_rng(8)
    for which in ((0, 1) if media else (0,)):
        m.knn_build(which)
        pos_p, pow_p, _, cnt = m.record_buffers(which)
        pos = pd.device_tensor(pos_p, cnt * 4, "<f4").cpu().numpy().reshape(cnt, 4).copy()
        pw = pd.device_tensor(pow_p, cnt * 4, "<f4").cpu().numpy().reshape(cnt, 4).copy()
        opos = pos.copy()
        if which == 0:   # the surface map keeps wall hits only: meta type field == 1
            meta = pos[:, 3].copy().view(np.uint32)
            is_wall = ((meta >> 5) & 3).astype(np.int32) - 1 == 1
            opos[~is_wall, :3] = np.nan
            assert m.knn_size(0)[0] == int(is_wall.sum())
        else:
            assert m.knn_size(1)[0] == 3 * n
        q = opos[rng.integers(0, cnt, 400)].copy()
        q = q[~np.isnan(q[:, 0])]
        q[:, :3] += rng.normal(0, 0.02, (q.shape[0], 3)).astype(np.float32)
        nq, k = q.shape[0], 50
        tq = torch.from_numpy(q).cuda()
        idx = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        d2 = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        c = torch.empty(nq, dtype=torch.int32, device="cuda")
        m.knn_query(which, tq, nq, k, float("inf"), idx, d2, c)
        rgb = torch.empty((nq, 4), dtype=torch.float32, device="cuda")
        m.knn_radiance(which, tq, nq, k, float("inf"), rgb)
        m.sync()
        oidx, od2, ocnt = oracle.knn_bruteforce(opos, q, k)
        assert np.array_equal(idx.cpu().numpy(), oidx) and np.array_equal(c.cpu().numpy(), ocnt)
        assert d2.cpu().numpy().tobytes() == od2.tobytes()
        est = oracle.knn_estimate(pw, oidx, od2, ocnt, volume=(which == 1))
        got = rgb.cpu().numpy()
        scale = np.abs(est).max()
        assert np.abs(got[:, :3] - est).max() <= 2e-5 * scale
    m.close()


@pytest.mark.parametrize("batched", [True, False])
@pytest.mark.parametrize("media", [False, True])
def test_render_knn_vs_oracle(pm, oracle, media, batched):
    """Mode B frame (pm_render_knn) against the oracle composition: eye-ray geometry from pm_oracle.c, brute-force k-NN
    estimates from knn_oracle.c, composited as documented in include/pmb200.h.  Tolerance 5e-5 of the frame maximum
    (FP32 power sums vs double).  batched: one lane per pixel, the tree walked once per 8 x 4 tile (pm_knn_set_batched);
    otherwise the warp-per-pixel renderer (the default)."""
    import torch
    from pmb200 import dist as pd
    n, w, h, k = 20000, 64, 48, 50
    w_s, w_v = 2.0e-4, 3.0e-2
    osc = oracle.default_scene(sz_img=48)
    osc.cam_ox = -8.0
    m = pm.PhotonMapper(n_photons=n)
    m.set_scene(copy_scene(pm.Scene, osc))
    m.init_random_numbers()
    m.set_record_capacity(16 * n)
    m.knn_set_batched(batched)
    m.clear_map()
    m.trace(0.0, media=media, records=True, no_map=True)
    m.knn_build(0)
    if media:
        m.knn_build(1)
    rgbf = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    rgba = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    m.render_knn(w, h, 0.0, media, k, float("inf"), w_s, w_v, rgba=rgba, rgbf=rgbf, y0=0, y1=20)
    m.render_knn(w, h, 0.0, media, k, float("inf"), w_s, w_v, rgba=rgba, rgbf=rgbf, y0=20, y1=h, y_step=2)   # interleaved rows
    m.render_knn(w, h, 0.0, media, k, float("inf"), w_s, w_v, rgba=rgba, rgbf=rgbf, y0=21, y1=h, y_step=2)
    m.sync()

    def records(which):
        pos_p, pow_p, _, cnt = m.record_buffers(which)
        pos = pd.device_tensor(pos_p, cnt * 4, "<f4").cpu().numpy().reshape(cnt, 4).copy()
        pw = pd.device_tensor(pow_p, cnt * 4, "<f4").cpu().numpy().reshape(cnt, 4).copy()
        return pos, pw

    hit, march = oracle.eye_geometry(osc, w, h, 0.0)
    spos, spow = records(0)
    meta = spos[:, 3].copy().view(np.uint32)
    spos[((meta >> 5) & 3).astype(np.int32) - 1 != 1, :3] = np.nan
    q = np.zeros((h * w, 4), np.float32); q[:, :3] = hit.reshape(-1, 8)[:, 3:6]
    idx, d2, cnt = oracle.knn_bruteforce(spos, q, k)
    surf = oracle.knn_estimate(spow, idx, d2, cnt, volume=False) * np.float32(w_s)
    on_wall = (hit.reshape(-1, 8)[:, 0] != 0) & (hit.reshape(-1, 8)[:, 1] == 1)
    surf[~on_wall] = 0
    want = np.zeros((h * w, 3), np.float32)
    if media:
        vpos, vpow = records(1)
        for s in range(10):
            q[:, :3] = march.reshape(-1, 10, 3)[:, s]
            idx, d2, cnt = oracle.knn_bruteforce(vpos, q, k)
            want += oracle.knn_estimate(vpow, idx, d2, cnt, volume=True) * np.float32(w_v)
        want += surf * np.float32(0.15)
    else:
        want += surf
    got = rgbf.cpu().numpy().reshape(-1, 4)
    scale = np.abs(want).max()
    assert scale > 0
    assert np.abs(got[:, :3] - want).max() <= 5e-5 * scale
    assert np.all(got[:, 3] == 1.0)
    u8 = rgba.cpu().numpy().reshape(-1, 4)
    expect_u8 = np.clip(np.nan_to_num(got[:, :3].astype(np.float64) * 255.0, nan=0.0), 0, 255).astype(np.uint8)
    assert np.array_equal(u8[:, :3], expect_u8) and np.all(u8[:, 3] == 0)
    m.close()


def test_legacy_cone_filter_estimator(pm, oracle):
    """the fixed-radius cone-filter estimate of the legacy file ("photonMappingKernel - Copy.cu":191-208,
    sqRadius 0.7, exposure 50) over the k nearest surface photons within the radius; against oracle/knn_oracle.c."""
    import torch
    from pmb200 import dist as pd
    n, k, sq_radius, exposure = 30000, 100, 0.7, 50.0
    osc = oracle.default_scene()
    m = pm.PhotonMapper(n_photons=n)
    m.set_scene(copy_scene(pm.Scene, osc))
    m.init_random_numbers()
    m.set_record_capacity(16 * n)
    m.clear_map()
    m.trace(0.0, media=False, records=True, no_map=True)
    m.knn_build(0)
    pos_p, pow_p, dir_p, cnt = m.record_buffers(0)
    pos = pd.device_tensor(pos_p, cnt * 4, "<f4").cpu().numpy().reshape(cnt, 4).copy()
    pw = pd.device_tensor(pow_p, cnt * 4, "<f4").cpu().numpy().reshape(cnt, 4).copy()
    dr = pd.device_tensor(dir_p, cnt * 4, "<f4").cpu().numpy().reshape(cnt, 4).copy()
    meta = pos[:, 3].copy().view(np.uint32)
    typ = ((meta >> 5) & 3).astype(np.int32) - 1
    wid = ((meta >> 7) & 15).astype(np.int32) - 1
    opos = pos.copy(); opos[typ != 1, :3] = np.nan
    rng = np.random.default_rng(12)
    inside = (typ == 1) & (np.abs(pos[:, 0]) <= 1.6) & (np.abs(pos[:, 1]) <= 1.6) & (pos[:, 2] >= 0) & (pos[:, 2] <= 6.1)
    pick = rng.choice(np.flatnonzero(inside), 500, replace=False)
    q = pos[pick].copy()
    q[:, 3] = wid[pick].astype(np.float32)          # query = a point on that wall + its wall id
    q[:10, 3] = 7.0                                   # unknown wall id -> zero
    tq = torch.from_numpy(q).cuda()
    rgb = torch.empty((500, 4), dtype=torch.float32, device="cuda")
    m.knn_radiance_cone(tq, 500, k, sq_radius, exposure, rgb)
    m.sync()
    idx, d2, c = oracle.knn_bruteforce(opos, q, k, sq_radius)
    normals = np.zeros((5, 3), np.float32)
    for i in range(5):
        normals[i, int(osc.planes[i][0])] = -1.0 if osc.planes[i][1] > 0 else 1.0
    want = oracle.knn_cone_estimate(pos, dr, pw, idx, d2, c, q[:, 3].astype(np.int32), normals, exposure)
    got = rgb.cpu().numpy()
    assert np.array_equal(got[:, 3], want[:, 3])
    assert np.all(got[:10] == 0)
    assert np.abs(got[:, :3] - want[:, :3]).max() <= 2e-5 * np.abs(want[:, :3]).max()
    assert want[10:, 3].mean() > 10
    m.close()
