"""Full-size property tests (BASELINE.json sizes) and the statistical check against the reference's own CUDA kernels.
The oracle cannot run 16M photons in seconds, so these use size-independent properties: determinism, shard invariance,
sortedness / bijectivity of the sort, and brute-force verification of a random SUBSET of k-NN queries."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mode_a_full_size_determinism_and_shard_invariance(pm):
    """16M photons, media on: two runs give bit-identical maps; 8 photon shards summed give the same accumulators."""
    n = 16777216
    m = pm.PhotonMapper(n_photons=n)
    m.init_random_numbers()
    st = m.get_mwc_state()
    m.emit(0.0, media=True)
    g1 = m.get_map()
    acc1 = m.get_accumulators()
    m.set_mwc_state(*st)
    m.emit(0.0, media=True)
    assert g1.tobytes() == m.get_map().tobytes()
    m.clear_map()
    for r in range(8):
        m.set_mwc_state(*st)
        m.set_photon_range(n * r // 8, n * (r + 1) // 8)
        m.trace(0.0, media=True)
    m.build_map()
    head = pm.ACC_HIT_ENTRIES + 32 * 32 * 32 * 3
    fold = lambda a: np.concatenate([a[:head], a[head:].reshape(-1, 32 * 32 * 32).sum(0)])
    assert np.array_equal(fold(acc1), fold(m.get_accumulators()))
    assert g1.tobytes() == m.get_map().tobytes()
    # energy bookkeeping: every photon deposits its three volume photons (5e-5 * (9 + 8 + 7) each, all three channels)
    vol = acc1[head:].sum() / 2.0 ** 36
    assert abs(vol - n * 5e-5 * 24) <= 1e-6 * n * 5e-5 * 24
    m.close()


def test_mode_b_full_size_properties(pm, oracle):
    """4M photons (BASELINE config 3): sorted keys are sorted, the permutation is a bijection onto the kept records, and
    64 random k=100 queries agree bit-exactly with brute force over all 9.6M wall photons."""
    import torch
    from pmb200 import dist as pd
    n = 4194304
    m = pm.PhotonMapper(n_photons=n)
    m.init_random_numbers()
    m.set_record_capacity(int(2.6 * n))
    m.clear_map()
    m.trace(0.0, media=False, records=True, no_map=True)
    m.knn_build(0)
    npts, levels = m.knn_size(0)
    pos_p, pow_p, _, cnt = m.record_buffers(0)
    assert 2 * n < npts <= cnt and levels == 4
    keys, perm = m.knn_sorted(0, npts)
    assert np.all(keys[1:] >= keys[:-1])
    assert len(np.unique(perm)) == npts and perm.max() < cnt
    pos = pd.device_tensor(pos_p, cnt * 4, "<f4").cpu().numpy().reshape(cnt, 4).copy()
    meta = pos[:, 3].copy().view(np.uint32)
    is_wall = ((meta >> 5) & 3).astype(np.int32) - 1 == 1
    assert int(is_wall.sum()) == npts and np.all(is_wall[perm])
    opos = pos.copy(); opos[~is_wall, :3] = np.nan
    rng = np.random.default_rng(11)
    q = opos[perm[rng.integers(0, npts, 64)]].copy()
    q[:, :3] += rng.normal(0, 0.01, (64, 3)).astype(np.float32)
    k = 100
    tq = torch.from_numpy(q).cuda()
    idx = torch.empty((64, k), dtype=torch.int32, device="cuda"); d2 = torch.empty((64, k), dtype=torch.float32, device="cuda")
    c = torch.empty(64, dtype=torch.int32, device="cuda")
    m.knn_query(0, tq, 64, k, float("inf"), idx, d2, c)
    m.sync()
    oidx, od2, ocnt = oracle.knn_bruteforce(opos, q, k)
    assert np.array_equal(idx.cpu().numpy(), oidx) and d2.cpu().numpy().tobytes() == od2.tobytes()
    assert np.all(c.cpu().numpy() == k) and np.all(np.diff(d2.cpu().numpy(), axis=1) >= 0)
    m.close()


def test_structural_agreement_with_reference_cuda_kernel(pm, oracle):
    """The reference's own CUDA kernels (photonMappingKernel.cu recompiled for sm_100a, oracle/_ref) on the same table.
    On a B200 its non-atomic `+=` (PMK:1068, :1158, :1177) loses most concurrent deposits -- measured here: its voxel map
    holds only ~27 % of the energy the sequential execution (and this build) deposits -- so its frame is the same picture,
    darker.  Hence a STRUCTURAL check: its map never holds more energy than ours, and the two frames correlate (>= 0.6;
    measured 0.75-0.8 -- the lost updates are concentrated on the brightest voxels and ours saturates there, so the
    reference's frame is not a uniformly scaled copy)."""
    import torch
    path = os.path.join(ROOT, "oracle", "_ref", "libpmref_cuda_10000.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libpmref_cuda_10000.so not built")
    n, w, h = 10000, 512, 512
    m = pm.PhotonMapper(n_photons=n)
    m.init_random_numbers()
    table = m.get_random_table()
    L = C.CDLL(path)
    assert L.refcu_capacity() == n
    assert L.refcu_set_table(table.ctypes.data_as(C.c_void_p), n) == 0
    assert L.refcu_set_szimg(512) == 0
    ref = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    rgrid = np.zeros((32, 32, 32, 3), np.float32)
    for media in (0, 1):
        L.refcu_emit(C.c_float(0.0), 0, media)
        L.refcu_render(C.c_void_p(ref.data_ptr()), w, h, C.c_float(0.0), 0, media)
        torch.cuda.synchronize()
        assert L.refcu_get_grid(rgrid.ctypes.data_as(C.c_void_p)) == 0
        a = ref.cpu().numpy()[..., :3].astype(np.float64)
        m.set_energy_scale(1.0)
        m.emit(0.0, media=bool(media))
        ours = m.get_map()
        assert 0.05 * ours.sum() < rgrid.sum() <= 1.001 * ours.sum(), (media, rgrid.sum(), ours.sum())
        u8, _ = m.render(w, h, 0.0, False, bool(media), want_f32=False)
        b = u8[..., :3].astype(np.float64)
        corr = np.corrcoef(a.ravel(), b.ravel())[0, 1]
        assert corr >= 0.6, (media, corr)      # measured 0.75-0.8: the races hit the brightest voxels hardest
    m.close()
