"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test drives libpmb200.so through its C-ABI
(pmb200 = ctypes over include/pmb200.h) and checks it against the CPU oracle (oracle/pm_oracle.c) on the same
seeded inputs.  Bars (BASELINE.json north_star):
  * integer / byte / index work (MWC stream, photon records' ids, uchar4 frame): bit-exact;
  * photon positions / power: north_star allows 1e-5 relative; we assert BIT-exact (same FP32 operation order);
  * float framebuffer for an identical photon map: bit-exact;
  * photon map (float sums over thousands of deposits, summed in a different order than the sequential
    reference): |diff| <= 2e-5 * max|map| -- the tolerance is the FP32 rounding of the REFERENCE's sequential
    sums, the product's sums are exact int64 fixed point;
  * rendered frame from the product's own map vs the oracle's: relative L1 <= 1e-4, PSNR >= 60 dB.
"""
import numpy as np
import pytest

from tests.util import bits_equal, cfg1_scene, copy_scene, grid_from_records

pytestmark = pytest.mark.gpu

MAP_TOL = 2e-5


def _mapper(pm, n, scene=None):
    m = pm.PhotonMapper(n_photons=n)
    if scene is not None:
        m.set_scene(scene)
    return m


def test_mwc_table_bit_exact(pm, oracle):
    """launch_init_random_numbers_kernel replacement: parallel jump-ahead == the serial MWC loop (PMK:1483-1498)."""
    for n in (3, 1000, 10000, 300001):
        m = _mapper(pm, n)
        m.init_random_numbers()
        tab = m.get_random_table()
        ref, st = oracle.mwc_table(n)
        assert bits_equal(tab, ref), n
        assert m.get_mwc_state() == st
        # a second call continues the stream, as the reference's global m_w/m_z would
        m.init_random_numbers()
        ref2, st2 = oracle.mwc_table(n, *st)
        assert bits_equal(m.get_random_table(), ref2)
        assert m.get_mwc_state() == st2
        m.close()


@pytest.mark.parametrize("media", [False, True])
@pytest.mark.parametrize("t", [0.0, 0.7])
@pytest.mark.parametrize("scene_name", ["default", "cfg1", "backwall5", "smoke", "shifted", "light_outside", "narrow"])
def test_trace_records_and_map(pm, oracle, media, t, scene_name):
    """Stage 1: photon records bit-exact (position, direction, power, object ids, order), photon map within MAP_TOL."""
    n = 20000
    osc = oracle.default_scene()
    if scene_name == "cfg1":
        cfg1_scene(osc)
    elif scene_name == "backwall5":
        # back wall at z = 5 (as in the legacy variant, "photonMappingKernel - Copy.cu":28): its hits fall in voxel slab 26, not on
        # the map boundary slab 31 that splatEnergy hard-codes -> exercises the per-photon off-slab fallback of store_photon
        osc.planes[4][1] = 5.0
    elif scene_name == "smoke":
        osc.n_spheres = 3      # spheres[2], the large smoke sphere of the reference's screenshots (PMK:65)
    elif scene_name == "shifted":
        # asymmetric wall offsets: still the reference's object layout, so the trace kernel's branch-free one-division-per-axis wall test
        # (pm_math.cuh ray_walls_std) runs on offsets other than +-1.5 / 6
        for i, off in enumerate((1.2, -1.1, -1.7, 1.6, 5.5)):
            osc.planes[i][1] = off
    elif scene_name == "light_outside":
        # the light beyond the x = +1.5 wall: primary rays see BOTH x walls in front (the nearer one must win), some start on no side
        osc.light[0] = 2.5
    elif scene_name == "narrow":
        # x walls 0.8 apart: fails the separation condition of ray_walls_std, so the generic per-wall instantiation must be selected
        osc.planes[0][1] = 0.4
        osc.planes[2][1] = -0.4
    m = _mapper(pm, n, copy_scene(pm.Scene, osc))
    m.init_random_numbers()
    table, st = oracle.mwc_table(n)
    m.set_record_capacity(16 * n)
    m.clear_map()
    m.trace(t, media=media, records=True)
    m.build_map()
    rec = m.get_records()
    grid = m.get_map()
    ogrid, orec, ost = oracle.emit(osc, table, 0, n, t, media, rng=st, max_records=16 * n)
    assert len(rec) == len(orec)
    assert rec.tobytes() == orec.tobytes()
    assert m.get_mwc_state() == ost
    scale = np.abs(ogrid).max()
    assert np.abs(grid - ogrid).max() <= MAP_TOL * scale
    m.close()


def test_map_matches_exact_sum_of_records(pm, oracle):
    """The product's map is the correctly rounded exact sum of the deposits (int64 fixed point), so it must agree
    with a float64 accumulation of the oracle's records to ~1 ulp -- much tighter than against the reference's
    own sequential FP32 sums."""
    n = 3000
    osc = oracle.default_scene()
    m = _mapper(pm, n, copy_scene(pm.Scene, osc))
    m.init_random_numbers()
    table, st = oracle.mwc_table(n)
    m.emit(0.0, media=True)
    grid = m.get_map()
    _, orec, _ = oracle.emit(osc, table, 0, n, 0.0, True, rng=st, max_records=16 * n)
    exact = grid_from_records(orec)
    scale = np.abs(exact).max()
    assert np.abs(grid - exact).max() <= 3e-7 * scale
    m.close()


@pytest.mark.parametrize("interp", [False, True])
@pytest.mark.parametrize("media", [False, True])
@pytest.mark.parametrize("t", [0.0, 1.3])
def test_render_bit_exact_on_injected_map(pm, oracle, interp, media, t):
    """Stages 3-5 on the oracle's photon map: float framebuffer and uchar4 frame bit-exact (all four flag
    combinations of the legacy ABI, static and animated scene)."""
    n, w, h = 20000, 256, 256
    osc = oracle.default_scene(sz_img=256)
    table, st = oracle.mwc_table(n)
    ogrid, _, _ = oracle.emit(osc, table, 0, n, t, media, rng=st)
    oimg, ou8 = oracle.render(osc, ogrid, w, h, t, interp, media)
    m = _mapper(pm, n, copy_scene(pm.Scene, osc))
    m.set_map(ogrid)
    u8, f32 = m.render(w, h, t, interp, media)
    assert bits_equal(f32[..., :3], oimg)
    assert np.all(f32[..., 3] == 1.0)
    assert np.array_equal(u8, ou8)
    m.close()


@pytest.mark.parametrize("interp", [False, True])
def test_smoke_sphere_scene_frame(pm, oracle, interp):
    """The smoke-sphere scene (n_spheres = 3, PMK:65) end to end: eye rays that end on the third sphere gather nothing (integrate /
    interpolateEnergy return 0 for spheres), the ray-march still sees the medium in front of it.  Float frame bit-exact on the
    oracle's map; product map within tolerance; the sphere covers a visible part of the frame."""
    n, w, h, t = 20000, 256, 256, 0.4
    osc = oracle.default_scene(sz_img=256)
    osc.n_spheres = 3
    table, st = oracle.mwc_table(n)
    ogrid, _, _ = oracle.emit(osc, table, 0, n, t, True, rng=st)
    oimg, ou8 = oracle.render(osc, ogrid, w, h, t, interp, True)
    hit, _ = oracle.eye_geometry(osc, w, h, t)
    on_smoke = (hit[..., 0] > 0) & (hit[..., 1] == 0) & (hit[..., 2] == 2)
    assert on_smoke.mean() > 0.05
    m = _mapper(pm, n, copy_scene(pm.Scene, osc))
    m.init_random_numbers()
    m.emit(t, media=True)
    assert np.abs(m.get_map() - ogrid).max() <= MAP_TOL * np.abs(ogrid).max()
    m.set_map(ogrid)
    u8, f32 = m.render(w, h, t, interp, True)
    assert bits_equal(f32[..., :3], oimg) and np.array_equal(u8, ou8)
    m.close()


def test_render_rect_frame_and_row_bands(pm, oracle):
    """Non-square frame with a camera offset, rendered in two row bands (the multi-GPU screen split)."""
    import torch
    n, w, h = 5000, 320, 180
    osc = oracle.default_scene(sz_img=180)
    osc.cam_ox = -70.0
    table, st = oracle.mwc_table(n)
    ogrid, _, _ = oracle.emit(osc, table, 0, n, 0.0, True, rng=st)
    oimg, ou8 = oracle.render(osc, ogrid, w, h, 0.0, False, True)
    m = _mapper(pm, n, copy_scene(pm.Scene, osc))
    m.set_map(ogrid)
    rgba = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    rgbf = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    m.render_device(w, h, 0.0, False, True, rgba=rgba, rgbf=rgbf, y0=0, y1=77)
    m.render_device(w, h, 0.0, False, True, rgba=rgba, rgbf=rgbf, y0=77, y1=h)
    m.sync()
    assert bits_equal(rgbf.cpu().numpy()[..., :3], oimg)
    assert np.array_equal(rgba.cpu().numpy(), ou8)
    m.close()


def test_full_frame_vs_oracle(pm, oracle):
    """Whole path (trace -> map -> render) against the whole oracle path: image tolerance from north_star
    (relative L1 <= 1e-4, PSNR >= 60 dB on the float frame; uchar4 frame differs in <= 0.1% of channel values, by 1)."""
    n, w, h = 50000, 256, 256
    osc = oracle.default_scene(sz_img=256)
    table, st = oracle.mwc_table(n)
    ogrid, _, _ = oracle.emit(osc, table, 0, n, 0.0, True, rng=st)
    oimg, ou8 = oracle.render(osc, ogrid, w, h, 0.0, False, True)
    m = _mapper(pm, n, copy_scene(pm.Scene, osc))
    m.init_random_numbers()
    u8 = np.empty((h, w, 4), np.uint8); f32 = np.empty((h, w, 4), np.float32)
    m.frame(w, h, 0.0, emit=True, interp=False, media=True, out_u8=u8, out_f32=f32)
    img = f32[..., :3]
    rel_l1 = np.abs(img - oimg).sum() / np.abs(oimg).sum()
    mse = np.mean((img.astype(np.float64) - oimg) ** 2)
    psnr = 10 * np.log10(float(oimg.max()) ** 2 / mse) if mse > 0 else np.inf
    assert rel_l1 <= 1e-4, rel_l1
    assert psnr >= 60.0, psnr
    d = np.abs(u8.astype(np.int32) - ou8.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() <= 1e-3
    m.close()


def test_photon_range_sharding_is_exact(pm, oracle):
    """Tracing [0,n) in one go and in three shards gives bit-identical accumulators (int64, order independent)
    and therefore a bit-identical map: the multi-GPU invariance the design relies on."""
    n = 30000
    m = _mapper(pm, n)
    m.init_random_numbers()
    st = m.get_mwc_state()
    m.clear_map(); m.trace(0.0, media=True); m.build_map()
    whole = m.get_accumulators()
    g_whole = m.get_map()
    m.set_mwc_state(*st)
    m.clear_map()
    for a, b in ((0, 7000), (7000, 7001), (7001, n)):
        m.set_mwc_state(*st)
        m.set_photon_range(a, b)
        m.trace(0.0, media=True)
    m.build_map()
    parts = m.get_accumulators()
    assert np.array_equal(_fold(pm, whole), _fold(pm, parts))
    assert bits_equal(g_whole, m.get_map())
    m.close()


def test_progressive_passes_accumulate(pm, oracle):
    """Progressive photon mapping (BASELINE config 5): several passes accumulate into the same map without clearing; every
    pass re-fills the direction table from the continuing MWC stream, as repeated display() frames of the reference would
    if the grid were not cleared.  The MWC state stays bit-exact, the map stays within MAP_TOL of the sequential oracle."""
    n, passes = 6000, 3
    osc = oracle.default_scene()
    m = _mapper(pm, n, copy_scene(pm.Scene, osc))
    m.clear_map()
    st = (6548, 316)
    ogrid = np.zeros((32, 32, 32, 3), np.float32)
    for _ in range(passes):
        m.init_random_numbers()
        m.trace(0.0, media=True)
        table, st = oracle.mwc_table(n, *st)
        ogrid, _, st = oracle.emit(osc, table, 0, n, 0.0, True, rng=st, grid=ogrid)
        assert m.get_mwc_state() == st
    m.set_energy_scale(1.0 / passes)
    m.build_map()
    grid = m.get_map()
    want = ogrid / np.float32(passes)
    assert np.abs(grid - want).max() <= MAP_TOL * np.abs(want).max()
    m.close()


def test_philox_table_and_trace(pm, oracle):
    """Counter-based RNG mode: the Philox table is bit-exact with the oracle's, and a trace over it matches the oracle
    consuming the same table (records bit-exact)."""
    n = 20000
    osc = oracle.default_scene()
    m = _mapper(pm, n, copy_scene(pm.Scene, osc))
    m.init_random_numbers_philox(0x5EED)
    table = oracle.philox_table(n, 0x5EED)
    assert bits_equal(m.get_random_table(), table)
    assert m.get_mwc_state() == (6548, 316)
    m.set_record_capacity(16 * n)
    m.clear_map(); m.trace(0.3, media=True, records=True); m.build_map()
    ogrid, orec, _ = oracle.emit(osc, table, 0, n, 0.3, True, rng=(6548, 316), max_records=16 * n)
    assert m.get_records().tobytes() == orec.tobytes()
    assert np.abs(m.get_map() - ogrid).max() <= MAP_TOL * np.abs(ogrid).max()
    m.close()


def test_sharded_table_generation(pm, oracle):
    """A rank of a multi-GPU job generates only its own rows of the direction table (plus rows 0..2, which every photon's
    medium walk reads): the rows, the MWC state and the traced records equal the oracle's for that photon range."""
    n, a, b = 20000, 7001, 13000
    osc = oracle.default_scene()
    m = _mapper(pm, n, copy_scene(pm.Scene, osc))
    m.set_photon_range(a, b)
    m.init_random_numbers()
    table, st = oracle.mwc_table(n)
    got = m.get_random_table()
    assert bits_equal(got[a:b], table[a:b]) and bits_equal(got[:3], table[:3])
    assert m.get_mwc_state() == st
    m.set_record_capacity(16 * (b - a))
    m.clear_map(); m.trace(0.0, media=True, records=True, no_map=True)
    _, st_shard = oracle.mwc_draws(9 * a, *st)
    _, orec, _ = oracle.emit(osc, table, a, b, 0.0, True, rng=st_shard, max_records=16 * (b - a), want_grid=False)
    assert m.get_records().tobytes() == orec.tobytes()
    m.set_photon_range(0, n)                       # widening the range after a sharded fill is refused
    with pytest.raises(pm.PmError, match="photon range"):
        m.trace(0.0)
    m.close()


def _fold(pm, acc):
    """Accumulators with the grey replicas summed (a CTA picks its replica by block index, so only the sum is
    shard-invariant)."""
    head = pm.ACC_HIT_ENTRIES + 32 * 32 * 32 * 3
    return np.concatenate([acc[:head], acc[head:].reshape(-1, 32 * 32 * 32).sum(0)])


def test_legacy_abi_frame(pm, oracle):
    """The three reference launchers in display() order (callbacksPBO.cpp:55-63): initRandomNumbers once, emit, render
    into a device uchar4 buffer.  Default context: 10 000 photons, 512x512, szImg 512."""
    import torch
    w = h = 512
    n = 10000
    osc = oracle.default_scene()
    table, st = oracle.mwc_table(n)
    pos = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    pm.launch_init_random_numbers_kernel()
    for t, interp, media in ((0.0, False, False), (0.25, True, True)):
        pm.launch_emit_photons_kernel(pos, w, h, t, interp, media)
        pm.launch_photon_mapping_kernel(pos, w, h, t, interp, media)
        ogrid, _, st2 = oracle.emit(osc, table, 0, n, t, media, rng=st)
        st = st2
        _, ou8 = oracle.render(osc, ogrid, w, h, t, interp, media)
        got = pos.cpu().numpy()
        d = np.abs(got.astype(np.int32) - ou8.astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() <= 2e-3, (t, d.max(), (d > 0).mean())
        assert np.all(got[..., 3] == 0)


@pytest.mark.parametrize("warps", [1, 3, 6, 16])
def test_fused_trace_equals_split_trace(pm, oracle, warps):
    """The warp-specialised trace kernel (medium walk on `warps` warps of every CTA, surface walk on the rest) and the two
    separate launches (PM_TRACE_SPLIT) produce identical accumulators, identical volume records and the same MWC state."""
    n = 50000
    m = _mapper(pm, n)
    m.init_random_numbers()
    st = m.get_mwc_state()
    m.clear_map(); m.trace(0.3, media=True, split=True)
    split = _fold(pm, m.get_accumulators())
    st_split = m.get_mwc_state()
    m.set_mwc_state(*st)
    m.set_volume_warps(warps)
    m.clear_map(); m.trace(0.3, media=True)
    assert np.array_equal(split, _fold(pm, m.get_accumulators()))
    assert m.get_mwc_state() == st_split
    # a partial range that does not divide by anything
    m.set_mwc_state(*st); m.set_photon_range(777, 31001)
    m.clear_map(); m.trace(0.3, media=True, split=True)
    split = _fold(pm, m.get_accumulators())
    m.set_mwc_state(*st)
    m.clear_map(); m.trace(0.3, media=True)
    assert np.array_equal(split, _fold(pm, m.get_accumulators()))
    m.close()


def test_pipelined_frames_equal_synchronous_frames(pm, oracle):
    """pm_frame_host_async / pm_frame_wait: five animated frames submitted back to back (copy of frame f under the trace of
    frame f+1, a ring of three device frame buffers) are byte-identical to the same frames through the synchronous pm_frame_host."""
    import torch
    w, h, n = 640, 360, 40000
    times = [0.0, 0.2, 0.4, 0.6, 0.8]
    ref = _mapper(pm, n)
    ref.init_random_numbers()
    want = []
    for t in times:
        u8 = np.empty((h, w, 4), np.uint8)
        ref.frame(w, h, t, emit=True, interp=True, media=True, out_u8=u8, out_f32=None)
        want.append(u8)
    ref.close()
    m = _mapper(pm, n)
    m.init_random_numbers()
    bufs = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    got, prev = [], None
    for f, t in enumerate(times):
        tk = m.frame_async(w, h, bufs[f & 1], t=t, emit=True, interp=True, media=True)
        if prev is not None:
            m.frame_wait(prev[0])
            got.append(bufs[prev[1]].numpy().copy())
        prev = (tk, f & 1)
    m.frame_wait(prev[0])
    got.append(bufs[prev[1]].numpy().copy())
    for f in range(len(times)):
        assert np.array_equal(got[f], want[f]), f
    with pytest.raises(pm.PmError):
        m.frame_wait(0)          # only the three most recent tickets can be waited for
    m.close()


def test_older_variant_launchers(pm):
    """launch_render_kernel (kernelPBO.cu:295 + render_kernel :268): host float3 pixels -> device uchar4, unscaled, with the
    device's float -> unsigned char conversion (cvt.rzi.u32.f32, low byte); launch_kernel (kernelPBO.cu:317, body commented
    out in the reference) leaves the buffer untouched."""
    import torch
    w, h = 37, 11
    rng = np.random.default_rng(3)
    px = rng.uniform(-50, 400, (h * w, 3)).astype(np.float32)
    px[:8] = [[0, 255, 256], [-1, -0.5, 0.999], [np.nan, np.inf, -np.inf], [300, 511.9, 512], [1e10, 4294967296.0, 4294967040.0],
              [254.999, 255.5, 1.5], [65535.7, 65536, 1e-30], [-1e-30, 2.0, 3.0]]
    pos = torch.full((h, w, 4), 77, dtype=torch.uint8, device="cuda")
    pm.launch_render_kernel(pos, w, h, 0.0, px)
    got = pos.cpu().numpy().reshape(-1, 4)
    with np.errstate(invalid="ignore"):
        u = np.where(np.isnan(px), 0.0, np.clip(np.trunc(px.astype(np.float64)), 0.0, 4294967295.0)).astype(np.uint64)
    assert np.array_equal(got[:, :3], (u & 0xFF).astype(np.uint8))
    assert np.all(got[:, 3] == 0)
    before = pos.clone()
    pm.launch_kernel(pos, w, h, 0.0)
    assert torch.equal(pos, before)


def test_fast_path_division_selftest(pm):
    """The trace kernel divides by the ray component without the compiler's range check (pm_math.cuh fdiv_fastpath).  On 2^32
    operand pairs of the domain it argues about: every quotient checkDistance would accept is bit-identical to the IEEE division,
    every other one is rejected as well."""
    m = _mapper(pm, 16)
    bad, acc = m.selftest_fdiv(1 << 32, seed=12345)
    assert bad == 0, m.L.pm_last_error(m.h).decode()
    assert acc > (1 << 28)      # the accepted range is well covered
    m.close()


@pytest.mark.parametrize("table", ["zeros", "wide", "tiny_scales", "with_nan", "huge_scales", "zero_scale_row", "mwc"])
def test_medium_walk_filter_on_hostile_tables(pm, table):
    """volume_photon_fast's preconditions are checked in the kernel: with a zero table (the reference's state before
    launch_init_random_numbers_kernel), wide rows, tiny / huge / zero scale rows (rows 0..2 scale every photon's draws), or NaNs, the
    default trace must still give the accumulators of the exact walk."""
    n = 8192
    rng = np.random.default_rng(7)
    m = _mapper(pm, n)
    if table == "mwc":
        m.init_random_numbers()   # the reference's own table: entries up to 65537 in magnitude (randFloat, PMK:1039-1052)
    elif table != "zeros":
        tab = rng.uniform(-1.0, 1.0, (n, 3)).astype(np.float32)
        if table == "huge_scales":
            tab[:3] *= 1e12
        elif table == "zero_scale_row":
            tab[1] = 0.0
            tab[0, 2] = 0.0
        if table == "wide":
            tab *= 50.0
        elif table == "tiny_scales":
            tab[:3] *= 1e-4
        elif table == "with_nan":
            tab[5::97, 1] = np.nan
            tab[7::101] = 0.0
        m.set_random_table(tab)
    st = m.get_mwc_state()
    m.clear_map(); m.trace(0.0, media=True)
    fast = m.get_accumulators()
    m.set_mwc_state(*st)
    m.clear_map(); m.trace(0.0, media=True, exact_medium=True)
    assert np.array_equal(fast, m.get_accumulators())
    m.close()
