"""Error behaviour of the extended C-ABI: bad arguments and call-order violations return a negative pm_status with a
message (the legacy three symbols abort like checkCUDAError instead, PMK:49-55: tests/test_gpu_cli.py runs them in a subprocess)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rc(pm, fn, *args):
    return getattr(pm.lib(), fn)(*args)


def test_argument_and_state_errors(pm):
    import torch
    L = pm.lib()
    m = pm.PhotonMapper(n_photons=1000)
    h = m.h
    assert L.pm_set_photon_count(h, 2) == -2                       # PM_ERR_ARG: fewer than the three rows the medium walk reads
    assert L.pm_set_photon_count(h, 10 ** 9) == -2                 # beyond the MWC jump range
    assert L.pm_set_photon_range(h, 10, 5) == -2 and L.pm_set_photon_range(h, 0, 1001) == -2
    assert L.pm_set_mwc_state(h, 0, 316) == -2                     # zero state is a fixed point of the generator
    assert b"MWC" in L.pm_last_error(h)
    sc = pm.default_scene(); sc.n_planes = 6
    assert L.pm_set_scene(h, C.byref(sc)) == -2
    sc = pm.default_scene(); sc.sz_img = 0
    assert L.pm_set_scene(h, C.byref(sc)) == -2
    buf = torch.zeros((8, 8, 4), dtype=torch.uint8, device="cuda")
    assert L.pm_render(h, 0.0, False, False, 8, 8, 0, 8, C.c_void_p(buf.data_ptr()), None) == -3   # PM_ERR_STATE: no map built yet
    assert b"pm_build_map" in L.pm_last_error(h)
    m.emit(0.0)
    assert L.pm_render(h, 0.0, False, False, 8, 8, 4, 2, C.c_void_p(buf.data_ptr()), None) == -2   # y0 > y1
    assert L.pm_render(h, 0.0, False, False, 8, 8, 0, 8, C.c_void_p(buf.data_ptr()), None) == 0
    assert L.pm_trace(h, 0.0, pm.PM_TRACE_RECORDS) == -3           # records without pm_set_record_capacity
    q = torch.zeros((4, 4), dtype=torch.float32, device="cuda"); out = torch.zeros((4, 4), dtype=torch.float32, device="cuda")
    assert L.pm_knn_radiance(h, 0, C.c_void_p(q.data_ptr()), 4, 0, 1.0, C.c_void_p(out.data_ptr())) == -2     # k = 0
    assert L.pm_knn_radiance(h, 0, C.c_void_p(q.data_ptr()), 4, 129, 1.0, C.c_void_p(out.data_ptr())) == -2   # k > 128
    assert L.pm_knn_radiance(h, 2, C.c_void_p(q.data_ptr()), 4, 8, 1.0, C.c_void_p(out.data_ptr())) == -2     # no such map
    assert L.pm_knn_set_curve(h, 5) == -2
    # an empty map answers every query with "nothing found"
    idx = torch.full((4, 8), 7, dtype=torch.int32, device="cuda"); d2 = torch.zeros((4, 8), dtype=torch.float32, device="cuda")
    cnt = torch.full((4,), 7, dtype=torch.int32, device="cuda")
    m.knn_query(0, q, 4, 8, float("inf"), idx, d2, cnt)
    m.sync()
    assert np.all(cnt.cpu().numpy() == 0) and np.all(idx.cpu().numpy() == -1) and np.all(np.isinf(d2.cpu().numpy()))
    # record overflow is reported, not silently truncated
    m.init_random_numbers()
    m.set_record_capacity(10)
    m.clear_map(); m.trace(0.0, records=True, no_map=True)
    assert L.pm_knn_build(h, 0) == -3 and b"overflow" in L.pm_last_error(h)
    m.close()


def test_null_handles(pm):
    L = pm.lib()
    assert L.pm_sync(None) == -2 and L.pm_destroy(None) == -2 and L.pm_clear_map(None) == -2
    assert L.pm_create(None, 0) == -2
    assert L.pm_last_error(None) == b"null context"


def test_knn_map_goes_stale_when_its_records_are_rewritten(pm):
    """A k-NN map built over the context's record buffers points into them; re-tracing with records (or resizing the buffers)
    invalidates it: queries return PM_ERR_STATE until pm_knn_build runs again (they used to read rewritten / freed memory)."""
    import torch
    L = pm.lib()
    m = pm.PhotonMapper(n_photons=5000)
    m.init_random_numbers()
    m.set_record_capacity(40000)
    m.clear_map(); m.trace(0.0, records=True, no_map=True)
    m.knn_build(0)
    q = torch.zeros((4, 4), dtype=torch.float32, device="cuda")
    out = torch.zeros((4, 4), dtype=torch.float32, device="cuda")
    assert L.pm_knn_radiance(m.h, 0, C.c_void_p(q.data_ptr()), 4, 8, C.c_float(float("inf")), C.c_void_p(out.data_ptr())) == 0
    m.clear_map(); m.trace(0.1, records=True, no_map=True)          # rewrites the buffers the map points into
    assert L.pm_knn_radiance(m.h, 0, C.c_void_p(q.data_ptr()), 4, 8, C.c_float(float("inf")), C.c_void_p(out.data_ptr())) == -3
    assert b"rebuild" in L.pm_last_error(m.h)
    m.knn_build(0)
    assert L.pm_knn_radiance(m.h, 0, C.c_void_p(q.data_ptr()), 4, 8, C.c_float(float("inf")), C.c_void_p(out.data_ptr())) == 0
    m.set_record_capacity(50000)                                     # frees them
    assert L.pm_knn_radiance(m.h, 0, C.c_void_p(q.data_ptr()), 4, 8, C.c_float(float("inf")), C.c_void_p(out.data_ptr())) == -3
    m.close()
