"""Full-size tests (BASELINE.json sizes): the oracle at 16M photons where it is affordable (the MWC stream and table: 0.3 s; the first
and last 64k photons of the job; the photon maps of configs 2 and 4 from tests/golden/fullsize_maps.npz, one 35 s CPU run of the
sequential oracle, script beside it), size-independent properties elsewhere (determinism, shard invariance, sortedness /
bijectivity of the sort, brute-force verification of a random SUBSET of k-NN queries), and the statistical check against the
reference's own CUDA kernels."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


N16 = 16777216
GOLDEN = os.path.join(ROOT, "tests", "golden", "fullsize_maps.npz")


def test_mwc_table_and_state_at_16M_vs_serial_stream(pm, oracle):
    """All 16 777 216 rows of the direction table and the generator state after them, against the oracle's serial loop
    (PMK:1483-1498): 50 M steps, i.e. every level of the jump tables (1024^2 = 2^20 steps per entry of the third).  A second table
    continues the stream (steps 50 M .. 100 M), as progressive passes do."""
    m = pm.PhotonMapper(n_photons=N16)
    m.init_random_numbers()
    ref, st = oracle.mwc_table(N16)
    got = m.get_random_table()
    assert got.tobytes() == ref.tobytes()
    assert m.get_mwc_state() == st
    m.init_random_numbers()
    ref2, st2 = oracle.mwc_table(N16, *st)
    assert m.get_random_table().tobytes() == ref2.tobytes() and m.get_mwc_state() == st2
    # a sharded fill (rank 7 of 8) produces the same rows of the same stream
    m2 = pm.PhotonMapper(n_photons=N16)
    m2.set_photon_range(7 * N16 // 8, N16)
    m2.init_random_numbers()
    assert m2.get_random_table()[7 * N16 // 8:].tobytes() == ref[7 * N16 // 8:].tobytes() and m2.get_mwc_state() == st
    m.close(); m2.close()


@pytest.mark.parametrize("which", ["first", "last"])
def test_records_of_first_and_last_64k_photons_of_a_16M_job(pm, oracle, which):
    """Photon records (surface, shadow and the three medium-walk deposits with their nine MWC draws per photon) of photons
    [0, 65536) and [16M - 65536, 16M) of the 16M-photon job, byte-identical to the sequential oracle.  The last photon's medium
    draws sit 201 M steps into the stream (3 x 16M table draws + 9 x 16M): the oracle gets there step by step (pmo_mwc_skip)."""
    n0, n1 = (0, 65536) if which == "first" else (N16 - 65536, N16)
    table, st = oracle.mwc_table(N16)
    osc = oracle.default_scene()
    st0 = oracle.mwc_skip(9 * n0, *st)
    _, orec, _ = oracle.emit(osc, table, n0, n1, 0.0, True, rng=st0, max_records=16 * 65536, want_grid=False)
    m = pm.PhotonMapper(n_photons=N16)
    m.init_random_numbers()
    m.set_photon_range(n0, n1)
    m.set_record_capacity(16 * 65536)
    m.clear_map()
    m.trace(0.0, media=True, records=True)
    rec = m.get_records()
    assert len(rec) == len(orec) and rec.tobytes() == orec.tobytes()
    m.close()


@pytest.mark.parametrize("config,n,w,h", [("config2", 1 << 20, 1024, 1024), ("config4", N16, 1920, 1080)])
def test_photon_map_and_frame_at_baseline_size_vs_sequential_oracle(pm, oracle, config, n, w, h):
    """BASELINE configs 2 (1M photons, 1024^2) and 4 (16M photons, 1920x1080), media on, against the sequential oracle at the
    stated size (tests/golden/fullsize_maps.npz).  Measured and asserted:
      * vs the same FP32 deposits summed in float64 (the oracle's shadow grid): |diff| <= 1e-6 max|map| -- the product's sums are exact;
      * vs the reference's literal result (every deposit added to a float voxel in photon order): 1.2e-3 max|map| at 1M photons and
        6.1e-2 at 16M -- that error is the REFERENCE's: its own voxel sums have outgrown FP32 (the oracle's two maps differ by exactly
        as much), which is why the small-size tolerance of 2e-5 (tests/test_gpu_parity.py) does not carry over;
      * the frame rendered from the product's map vs the oracle's render of the exact-sum map on three row bands: rel-L1 <= 1e-5,
        PSNR >= 80 dB; vs the literal map's frame: rel-L1 <= 2e-2 (measured 1e-2 at 16M)."""
    z = np.load(GOLDEN)
    seq, chunked = z[config + "_map"], z[config + "_map_exact"]
    m = pm.PhotonMapper(n_photons=n)
    sc = pm.default_scene(sz_img=h); sc.cam_ox = -(w - h) / 2.0
    m.set_scene(sc)
    m.init_random_numbers()
    m.emit(0.0, media=True)
    ours = m.get_map()
    assert m.get_mwc_state() == tuple(int(x) for x in z[config + "_state"])
    mx = float(np.abs(chunked).max())
    d_exact = float(np.abs(ours.astype(np.float64) - chunked).max()) / mx
    d_seq = float(np.abs(ours.astype(np.float64) - seq).max()) / mx
    d_ref = float(np.abs(seq.astype(np.float64) - chunked).max()) / mx
    print("%s: |ours - exact| %.2e, |ours - sequential| %.2e, |sequential - exact| %.2e (of max|map|)" % (config, d_exact, d_seq, d_ref))
    assert d_exact <= 1e-6
    assert d_seq <= (2e-3 if n == 1 << 20 else 7e-2) and d_seq <= d_ref * 1.01 + 1e-6
    u8, f32 = m.render(w, h, 0.0, False, True)
    osc = oracle.default_scene(sz_img=h); osc.cam_ox = -(w - h) / 2.0
    worst_l1, worst_psnr, l1_seq = 0.0, 1e9, 0.0
    for y0 in (0, h // 2 - 20, h - 40):
        img, _ = oracle.render(osc, chunked, w, h, 0.0, False, True, y0=y0, y1=y0 + 40, want_u8=False)
        a, b = f32[y0:y0 + 40, :, :3].astype(np.float64), img[y0:y0 + 40].astype(np.float64)
        worst_l1 = max(worst_l1, np.abs(a - b).sum() / np.abs(b).sum())
        mse = ((a - b) ** 2).mean()
        worst_psnr = min(worst_psnr, 10 * np.log10(b.max() ** 2 / mse) if mse > 0 else 1e9)
        img2, _ = oracle.render(osc, seq, w, h, 0.0, False, True, y0=y0, y1=y0 + 40, want_u8=False)
        l1_seq = max(l1_seq, np.abs(a - img2[y0:y0 + 40]).sum() / np.abs(img2[y0:y0 + 40]).sum())
    print("%s frame: rel-L1 %.2e, PSNR %.1f dB vs the exact-sum map's frame; rel-L1 %.2e vs the literal map's frame" % (config, worst_l1, worst_psnr, l1_seq))
    assert worst_l1 <= 1e-5 and worst_psnr >= 80.0
    assert l1_seq <= 2e-2
    m.close()


@pytest.mark.parametrize("config,n,w,h,cfg1", [("config1", 65536, 256, 256, True), ("config2", 1 << 20, 1024, 1024, False)])
def test_image_distance_to_the_reference_device_build(pm, config, n, w, h, cfg1):
    """north_star: "images within a stated relative-L1/PSNR tolerance of the reference".  The reference kernel itself on this GPU,
    made deterministic where it can be -- its three racy `photons[..] += ..` (PMK:1068, :1158, :1177) as atomicAdd, -fmad=false
    (oracle/build_ref.sh P4) -- on the same direction table.  What is left between it and this build is the reference's DEVICE
    arithmetic: rsqrtf in normalize is rsqrt.approx (2 ulp) where the host-compiled reference and the oracle use 1/sqrtf, and
    because 88% of the wall bounces are decided by the last bit of the hit point a few photons bounce differently.
    Both maps go through the same renderer (bit-exact to the oracle), float frames compared.  Measured on B200 / asserted:
        surface only : map rel-L1 2.3e-3 / 1.7e-3 (config 1 / 2), frame rel-L1 7.6e-4 / 3.0e-4, PSNR 79.7 / 91.0 dB
                       -> asserted rel-L1 <= 3e-3, PSNR >= 70 dB, deposited energy within 0.1%
        media on     : the reference's medium walk draws from a global MWC state that all its threads race on (PMK:1026-1036), so
                       its volume photons are a different random sample every run: frame rel-L1 1.0e-1 / 9.3e-3, PSNR 31.8 / 56.3 dB
                       -> asserted rel-L1 <= 0.2 / 0.05, PSNR >= 25 / 40 dB (statistical agreement only; a second run gave 1.9e-2, 50.4 dB)."""
    import importlib.util
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpmref_cuda_atomic_%d.so" % n)):
        pytest.skip("oracle/_ref/libpmref_cuda_atomic_%d.so not built" % n)
    spec = importlib.util.spec_from_file_location("ref_device_distance", os.path.join(ROOT, "tools", "ref_device_distance.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    d = mod.distances(n, w, h, cfg1, False)
    print(config, "surface", d)
    assert d["frame_rel_l1"] <= 3e-3 and d["frame_psnr_db"] >= 70.0 and d["map_rel_l1"] <= 5e-3
    assert abs(d["map_energy_ratio"] - 1.0) <= 1e-3
    d = mod.distances(n, w, h, cfg1, True)
    print(config, "media", d)
    assert d["frame_rel_l1"] <= (0.2 if cfg1 else 0.05) and d["frame_psnr_db"] >= (25.0 if cfg1 else 40.0)
    assert abs(d["map_energy_ratio"] - 1.0) <= 1e-3


def test_config1_at_its_stated_size(pm, oracle):
    """BASELINE config 1 as stated: single sphere + floor in the medium, 65 536 photons, 256 x 256, k = 50.  Mode A: records
    byte-identical, map equal to the exact sum of the oracle's deposits (3e-7), float frame from the oracle's map bit-exact.  Mode B: the k = 50 index
    sets of 2048 eye-ray hit points bit-exact against brute force over all wall photons."""
    import torch
    from pmb200 import dist as pd
    from tests.util import cfg1_scene, copy_scene
    n, w, h, k = 65536, 256, 256, 50
    osc = cfg1_scene(oracle.default_scene(sz_img=h))
    table, st = oracle.mwc_table(n)
    exact = np.zeros((32, 32, 32, 3), np.float64)
    ogrid, orec, _ = oracle.emit(osc, table, 0, n, 0.0, True, rng=st, max_records=16 * n, shadow64=exact)
    oimg, ou8 = oracle.render(osc, ogrid, w, h, 0.0, False, True)
    m = pm.PhotonMapper(n_photons=n, scene=copy_scene(pm.Scene, osc))
    m.init_random_numbers()
    m.set_record_capacity(16 * n)
    m.clear_map(); m.trace(0.0, media=True, records=True); m.build_map()
    rec = m.get_records()
    assert rec.tobytes() == orec.tobytes()
    # every photon of this scene lands on the one floor: 65 536 sequential FP32 adds into voxels that reach 3e4 leave the reference's
    # own sums 5e-5 off (measured), so the tight bar is against the same deposits summed in double
    ours = m.get_map().astype(np.float64)
    assert np.abs(ours - exact).max() <= 3e-7 * np.abs(exact).max()
    assert np.abs(ours - ogrid).max() <= 2e-4 * np.abs(ogrid).max()
    m.set_map(ogrid)
    u8, f32 = m.render(w, h, 0.0, False, True)
    assert f32[..., :3].tobytes() == oimg.tobytes() and np.array_equal(u8, ou8)
    # Mode B, k = 50: queries = eye-ray hit points of every 32nd pixel
    m.knn_build(0)
    npts, _ = m.knn_size(0)
    pos_p, _, _, cnt = m.record_buffers(0)
    pos = pd.device_tensor(pos_p, cnt * 4, "<f4").cpu().numpy().reshape(cnt, 4).copy()
    meta = pos[:, 3].copy().view(np.uint32)
    opos = pos.copy(); opos[((meta >> 5) & 3).astype(np.int32) - 1 != 1, :3] = np.nan
    hit, _ = oracle.eye_geometry(osc, w, h, 0.0)
    q = np.zeros((2048, 4), np.float32)
    q[:, :3] = hit.reshape(-1, 8)[::32, 3:6]
    tq = torch.from_numpy(q).cuda()
    idx = torch.empty((2048, k), dtype=torch.int32, device="cuda"); d2 = torch.empty((2048, k), dtype=torch.float32, device="cuda")
    c = torch.empty(2048, dtype=torch.int32, device="cuda")
    m.knn_query(0, tq, 2048, k, float("inf"), idx, d2, c)
    m.sync()
    oidx, od2, ocnt = oracle.knn_bruteforce(opos, q, k)
    assert np.array_equal(idx.cpu().numpy(), oidx) and d2.cpu().numpy().tobytes() == od2.tobytes() and np.array_equal(c.cpu().numpy(), ocnt)
    m.close()


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


@pytest.mark.parametrize("t,light", [(0.0, None), (0.7, (0.3, -0.9, 2.5)), (0.0, (1.45, 1.45, 5.9))])
def test_medium_walk_filter_gives_the_exact_counts_at_16M(pm, t, light):
    """The Mode A medium walk computes deposit points approximately and redoes photons near a voxel boundary exactly
    (csrc/pm_trace.cu volume_photon_fast).  Its accumulators must equal, bit for bit, those of the walk that evaluates every point with
    the reference's arithmetic (PM_TRACE_EXACT_MEDIUM -- the routine the record tests pin to the oracle) on all 16M photons = 50M
    deposits, for the default light and for lights that put the deposits elsewhere (the last one: a corner, most points clamp)."""
    n = 16777216
    m = pm.PhotonMapper(n_photons=n)
    if light is not None:
        sc = m.get_scene()
        for i in range(3):
            sc.light[i] = light[i]
        m.set_scene(sc)
    m.init_random_numbers()
    st = m.get_mwc_state()
    m.clear_map()
    m.trace(t, media=True)
    fast = m.get_accumulators()
    m.set_mwc_state(*st)
    m.clear_map()
    m.trace(t, media=True, exact_medium=True)
    exact = m.get_accumulators()
    assert np.array_equal(fast, exact)
    assert exact[pm.ACC_HIT_ENTRIES + 32 * 32 * 32 * 3:].sum() != 0
    m.close()


@pytest.mark.parametrize("n,t,media,variant", [(16777216, 0.0, True, "default"), (16777216, 1.3, False, "default"), (1 << 20, 0.7, True, "shifted"),
                                               (1 << 20, 2.3, True, "default"), (1 << 20, 0.0, True, "light_low"), (1 << 20, 0.4, False, "big_spheres"),
                                               (100003, 0.0, True, "default"), (37, 0.0, True, "default")])
def test_two_phase_walk_gives_the_machine_s_deposits(pm, n, t, media, variant):
    """Mode A takes fresh photons through the common path in lock-step and queues the rest for the general state machine
    (csrc/pm_trace.cu, phase F / phase G).  Its accumulators must equal, bit for bit, those of the machine alone (PM_TRACE_ONE_PHASE)
    and those of the records instantiation -- the one the record tests pin to the oracle, photon by photon."""
    m = pm.PhotonMapper(n_photons=n)
    sc = m.get_scene()
    if variant == "shifted":
        for i, off in enumerate((1.2, -1.1, -1.7, 1.6, 5.5)):
            sc.planes[i][1] = off
    elif variant == "light_low":      # the light next to the floor and a wall: most shadow rays leave through them
        sc.light[0], sc.light[1], sc.light[2] = -1.3, -1.35, 0.4
    elif variant == "big_spheres":    # spheres that reach through walls: their shadow tests cannot be skipped
        sc.animate = 0
        sc.spheres[0][0], sc.spheres[0][1], sc.spheres[0][2], sc.spheres[0][3] = 1.2, -1.0, 3.0, 0.7
        sc.spheres[1][0], sc.spheres[1][1], sc.spheres[1][2], sc.spheres[1][3] = -0.5, 0.9, 5.6, 0.9
    m.set_scene(sc)
    m.init_random_numbers()
    st = m.get_mwc_state()
    accs = []
    for kw in ({}, {"one_phase": True}, {"records": True}):
        if "records" in kw:
            if n > (1 << 20):
                continue
            m.set_record_capacity(8 * n)
        m.set_mwc_state(*st)
        m.clear_map()
        m.trace(t, media=media, **kw)
        accs.append(m.get_accumulators())
    assert np.array_equal(accs[0], accs[1])
    assert np.abs(accs[1][:pm.ACC_HIT_ENTRIES]).sum() > 0
    if len(accs) == 3:
        head = pm.ACC_HIT_ENTRIES + 32 * 32 * 32 * 3
        assert np.array_equal(accs[0][:head], accs[2][:head])
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
