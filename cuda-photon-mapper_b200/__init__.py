"""pmb200 -- host-side Python mirror of the C-ABI in include/pmb200.h (ctypes; no compute happens here).

The product is libpmb200.so (hand-written sm_100a CUDA behind a C ABI).  This module only loads it and
marshals pointers: numpy / torch buffers in, status codes out.  It fails loudly if the library is missing
or no CUDA device is present -- there is no CPU fallback and the oracle under oracle/ is never imported here.

Reference call surface mirrored (simplePBO.cpp:125-182, callbacksPBO.cpp:47-101):
    initRandomNumbers()            -> PhotonMapper.init_random_numbers() / launch_init_random_numbers_kernel()
    simpleRunCuda(interp, media)   -> PhotonMapper.emit(t, media)        / launch_emit_photons_kernel(...)
    photonMappingCuda(interp, media)-> PhotonMapper.render(...)           / launch_photon_mapping_kernel(...)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMB200_LIB") or os.path.join(_HERE, "libpmb200.so")   # PMB200_LIB: development override

PM_TRACE_MEDIA, PM_TRACE_RECORDS, PM_TRACE_NO_MAP, PM_TRACE_SPLIT, PM_TRACE_EXACT_MEDIUM, PM_TRACE_ONE_PHASE = 1, 2, 4, 8, 16, 32
GRID_N = 32
ACC_HIT_ENTRIES = 5 * 32 * 32 * 4
ACC_ENTRIES = ACC_HIT_ENTRIES + 32 * 32 * 32 * 3 + 32 * 32 * 32   # pm_layout.h: hit + vox rgb + grey

RECORD_DTYPE = np.dtype([("type", "<i4"), ("id", "<i4"), ("index", "<i4"), ("kind", "<i4"),
                         ("loc", "<f4", 3), ("dir", "<f4", 3), ("energy", "<f4", 3)])


class Scene(C.Structure):
    """pm_scene (include/pmb200_types.h)."""
    _fields_ = [("n_spheres", C.c_int32), ("n_planes", C.c_int32),
                ("spheres", (C.c_float * 4) * 3), ("planes", (C.c_float * 2) * 5),
                ("light", C.c_float * 3), ("sz_img", C.c_int32),
                ("cam_ox", C.c_float), ("cam_oy", C.c_float), ("animate", C.c_int32)]


class PmError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libpmb200.so (once).  Raises if it has not been built: the product never falls back to CPU code."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PmError(f"{LIB_PATH} is missing: build it with `make -C {_HERE}` (or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, f32, u32, b = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32, C.c_bool
        sig = {
            "pm_create": (i32, [C.POINTER(vp), i32]), "pm_destroy": (i32, [vp]),
            "pm_last_error": (C.c_char_p, [vp]), "pm_version": (C.c_char_p, []),
            "pm_default_context": (vp, []), "pm_set_stream": (i32, [vp, vp]), "pm_sync": (i32, [vp]),
            "pm_scene_default": (None, [C.POINTER(Scene)]), "pm_set_scene": (i32, [vp, C.POINTER(Scene)]),
            "pm_get_scene": (i32, [vp, C.POINTER(Scene)]),
            "pm_position_objects": (i32, [C.POINTER(Scene), f32, C.POINTER(Scene)]),
            "pm_trace_plan": (i32, [C.POINTER(Scene), f32, C.POINTER(i32), C.POINTER(C.c_uint32)]),
            "pm_set_photon_count": (i32, [vp, i64]), "pm_set_photon_range": (i32, [vp, i64, i64]),
            "pm_set_energy_scale": (i32, [vp, f32]),
            "pm_init_random_table": (i32, [vp]), "pm_init_random_table_philox": (i32, [vp, C.c_uint64]), "pm_set_random_table_host": (i32, [vp, vp, i64]),
            "pm_get_random_table_host": (i32, [vp, vp, i64]),
            "pm_set_mwc_state": (i32, [vp, u32, u32]), "pm_get_mwc_state": (i32, [vp, C.POINTER(u32), C.POINTER(u32)]),
            "pm_clear_map": (i32, [vp]), "pm_trace": (i32, [vp, f32, C.c_uint]), "pm_set_volume_warps": (i32, [vp, i32]),
            "pm_accumulators": (i32, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
            "pm_get_accumulators_host": (i32, [vp, vp]),
            "pm_build_map": (i32, [vp]), "pm_get_map_host": (i32, [vp, vp]), "pm_set_map_host": (i32, [vp, vp]),
            "pm_map_device": (i32, [vp, C.POINTER(vp)]),
            "pm_set_record_capacity": (i32, [vp, i64]), "pm_record_count": (i32, [vp, C.POINTER(i64)]),
            "pm_get_records_host": (i32, [vp, vp, i64]),
            "pm_record_buffers": (i32, [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]),
            "pm_knn_set_curve": (i32, [vp, i32]), "pm_knn_set_batched": (i32, [vp, b]), "pm_knn_build": (i32, [vp, i32]), "pm_knn_build_points": (i32, [vp, i32, vp, vp, i64, b]),
            "pm_knn_size": (i32, [vp, i32, C.POINTER(i64), C.POINTER(C.c_int32)]),
            "pm_knn_query": (i32, [vp, i32, vp, i64, i32, f32, vp, vp, vp]),
            "pm_knn_radiance": (i32, [vp, i32, vp, i64, i32, f32, vp]),
            "pm_knn_radiance_cone": (i32, [vp, vp, i64, i32, f32, f32, vp]),
            "pm_render_knn": (i32, [vp, f32, b, i32, i32, i32, i32, i32, f32, f32, f32, vp, vp]),
            "pm_render_knn_rows": (i32, [vp, f32, b, i32, i32, i32, i32, i32, i32, f32, f32, f32, vp, vp]),
            "pm_render_knn_host": (i32, [vp, f32, b, i32, i32, i32, f32, f32, f32, vp, vp]),
            "pm_knn_sorted_host": (i32, [vp, i32, vp, vp, i64]),
            "pm_knn_level_host": (i32, [vp, i32, i32, C.POINTER(i64), vp]),
            "pm_render": (i32, [vp, f32, b, b, i32, i32, i32, i32, vp, vp]),
            "pm_render_host": (i32, [vp, f32, b, b, i32, i32, vp, vp]),
            "pm_frame_host": (i32, [vp, f32, b, b, b, i32, i32, vp, vp]),
            "launch_render_kernel": (None, [vp, C.c_uint, C.c_uint, f32, vp]),
            "launch_kernel": (None, [vp, C.c_uint, C.c_uint, f32, vp, vp]),
            "pm_frame_host_async": (i32, [vp, f32, b, b, b, i32, i32, vp, C.POINTER(C.c_int64)]),
            "pm_frame_wait": (i32, [vp, C.c_int64]), "pm_reserve_frame": (i32, [vp, i32, i32]),
            "pm_frame_device": (i32, [vp, f32, b, b, b, i32, i32, vp, vp]), "pm_set_trace_sms": (i32, [vp, i32]),
            "pm_launch_count": (i64, [vp]),
            "pm_enable_timing": (i32, [vp, b]), "pm_kernel_count": (i32, []), "pm_kernel_name": (C.c_char_p, [i32]),
            "pm_get_timings": (i32, [vp, vp, vp]),
            "pm_peer_export": (i32, [vp, vp]), "pm_peer_connect": (i32, [vp, i32, i32, vp]),
            "pm_peer_connect_local": (i32, [vp, i32, i32, C.POINTER(vp)]), "pm_peer_disconnect": (i32, [vp]),
            "pm_peer_info": (i32, [vp, C.POINTER(i32), C.POINTER(i32)]), "pm_peer_barrier": (i32, [vp]), "pm_peer_status": (i32, [vp]),
            "pm_peer_set_timeout": (i32, [vp, C.c_double]),
            "pm_shared_alloc": (i32, [vp, C.c_size_t, C.POINTER(vp), vp]), "pm_shared_open": (i32, [vp, vp, C.POINTER(vp)]),
            "pm_shared_close": (i32, [vp, vp, b]), "pm_set_row_band": (i32, [vp, i32, i32]),
            "pm_group_create": (i32, [C.POINTER(vp), C.POINTER(i32), i32]), "pm_group_destroy": (i32, [vp]), "pm_group_size": (i32, [vp]),
            "pm_group_context": (vp, [vp, i32]), "pm_group_last_error": (C.c_char_p, [vp]),
            "pm_group_set_scene": (i32, [vp, C.POINTER(Scene)]), "pm_group_set_photon_count": (i32, [vp, i64]),
            "pm_group_set_energy_scale": (i32, [vp, f32]), "pm_group_init_random_table": (i32, [vp]),
            "pm_group_frame_host": (i32, [vp, f32, b, b, b, i32, i32, vp, vp]),
            "pm_group_frame_host_async": (i32, [vp, f32, b, b, b, i32, i32, vp, C.POINTER(C.c_int64)]),
            "pm_group_frame_wait": (i32, [vp, C.c_int64]),
            "pm_trace_profile": (i32, [vp, b]), "pm_get_trace_profile_host": (i32, [vp, vp, i64, C.POINTER(i64)]),
            "pm_selftest_fdiv": (i32, [vp, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
            "launch_init_random_numbers_kernel": (None, []),
            "launch_emit_photons_kernel": (None, [vp, C.c_uint, C.c_uint, f32, b, b]),
            "launch_photon_mapping_kernel": (None, [vp, C.c_uint, C.c_uint, f32, b, b]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def default_scene(sz_img=512, animate=1):
    s = Scene()
    lib().pm_scene_default(C.byref(s))
    s.sz_img = sz_img
    s.animate = animate
    return s


def trace_plan(scene, t=0.0):
    """(two_phase, [shadow_need of walls 0..4]): what a Mode A trace of `scene` at time t will do (pm_trace_plan; host side only)."""
    tp = C.c_int32()
    need = (C.c_uint32 * 5)()
    rc = lib().pm_trace_plan(C.byref(scene), t, C.byref(tp), need)
    if rc != 0:
        raise PmError(f"pmb200 error {rc}")
    return bool(tp.value), [int(x) for x in need]


def _ptr(a):
    """Device or host pointer of a numpy array, a torch tensor, an int address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class PhotonMapper:
    """One pm_context: a photon mapper bound to one GPU and one stream."""

    def __init__(self, device=-1, n_photons=10000, scene=None, _borrowed=None):
        self.L = lib()
        if _borrowed is not None:       # a rank of a PhotonGroup: the group owns the context
            self.h, self.n_photons, self._owned = C.c_void_p(_borrowed), n_photons, False
            return
        self._owned = True
        h = C.c_void_p()
        rc = self.L.pm_create(C.byref(h), device)
        if rc != 0:
            raise PmError({-4: "no CUDA device: pmb200 has no CPU fallback"}.get(rc, f"pm_create failed ({rc})"))
        self.h = h
        self.n_photons = 0
        self.set_photon_count(n_photons)
        if scene is not None:
            self.set_scene(scene)

    def close(self):
        if getattr(self, "h", None):
            if self._owned:
                self.L.pm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise PmError(f"pmb200 error {rc}: {self.L.pm_last_error(self.h).decode()}")

    # -- configuration -----------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._ck(self.L.pm_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        self._ck(self.L.pm_sync(self.h))

    def set_scene(self, scene):
        self._ck(self.L.pm_set_scene(self.h, C.byref(scene)))

    def get_scene(self):
        s = Scene()
        self._ck(self.L.pm_get_scene(self.h, C.byref(s)))
        return s

    def set_photon_count(self, n):
        self._ck(self.L.pm_set_photon_count(self.h, n))
        self.n_photons = n

    def set_photon_range(self, first, last):
        self._ck(self.L.pm_set_photon_range(self.h, first, last))

    def set_energy_scale(self, s):
        self._ck(self.L.pm_set_energy_scale(self.h, s))

    # -- random table ------------------------------------------------------------------------------
    def init_random_numbers(self):
        """initRandomNumbers() (simplePBO.cpp:125) == launch_init_random_numbers_kernel on this context."""
        self._ck(self.L.pm_init_random_table(self.h))

    def init_random_numbers_philox(self, seed=0x5EED):
        """Counter-based table: row i from Philox4x32-10(counter i, key seed)."""
        self._ck(self.L.pm_init_random_table_philox(self.h, seed))

    def set_random_table(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32)
        self._ck(self.L.pm_set_random_table_host(self.h, _ptr(xyz), xyz.shape[0]))

    def get_random_table(self, n=None):
        n = self.n_photons if n is None else n
        out = np.empty((n, 3), np.float32)
        self._ck(self.L.pm_get_random_table_host(self.h, _ptr(out), n))
        return out

    def set_mwc_state(self, w, z):
        self._ck(self.L.pm_set_mwc_state(self.h, w, z))

    def get_mwc_state(self):
        w, z = C.c_uint32(), C.c_uint32()
        self._ck(self.L.pm_get_mwc_state(self.h, C.byref(w), C.byref(z)))
        return w.value, z.value

    # -- stage 1 -------------------------------------------------------------------------------------
    def clear_map(self):
        self._ck(self.L.pm_clear_map(self.h))

    def trace(self, t=0.0, media=False, records=False, no_map=False, split=False, exact_medium=False, one_phase=False):
        flags = ((PM_TRACE_MEDIA if media else 0) | (PM_TRACE_RECORDS if records else 0) | (PM_TRACE_NO_MAP if no_map else 0)
                 | (PM_TRACE_SPLIT if split else 0) | (PM_TRACE_EXACT_MEDIUM if exact_medium else 0)
                 | (PM_TRACE_ONE_PHASE if one_phase else 0))
        self._ck(self.L.pm_trace(self.h, t, flags))

    def set_volume_warps(self, warps):
        self._ck(self.L.pm_set_volume_warps(self.h, warps))

    def accumulators(self):
        """(device pointer, number of int64 entries) of the exact accumulators."""
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.pm_accumulators(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def get_accumulators(self):
        out = np.empty(ACC_ENTRIES, np.int64)
        self._ck(self.L.pm_get_accumulators_host(self.h, _ptr(out)))
        return out

    def build_map(self):
        self._ck(self.L.pm_build_map(self.h))

    def emit(self, t=0.0, media=False):
        """simpleRunCuda() (simplePBO.cpp:166): clear the map, trace every photon, finalise the map."""
        self.clear_map()
        self.trace(t, media)
        self.build_map()

    def get_map(self):
        g = np.empty((32, 32, 32, 3), np.float32)
        self._ck(self.L.pm_get_map_host(self.h, _ptr(g)))
        return g

    def set_map(self, grid):
        grid = np.ascontiguousarray(grid, np.float32)
        assert grid.size == 32 * 32 * 32 * 3
        self._ck(self.L.pm_set_map_host(self.h, _ptr(grid)))

    def map_device_ptr(self):
        p = C.c_void_p()
        self._ck(self.L.pm_map_device(self.h, C.byref(p)))
        return p.value

    # -- records -------------------------------------------------------------------------------------
    def set_record_capacity(self, n):
        self._ck(self.L.pm_set_record_capacity(self.h, n))

    def record_count(self):
        n = C.c_int64()
        self._ck(self.L.pm_record_count(self.h, C.byref(n)))
        return n.value

    def get_records(self):
        n = self.record_count()
        out = np.zeros(n, RECORD_DTYPE)
        self._ck(self.L.pm_get_records_host(self.h, _ptr(out), n))
        return out

    def record_buffers(self, which=0):
        """(pos_meta ptr, power_index ptr, dir ptr or None, count) of the surface (0) or volume (1) record set."""
        a, b, c, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        self._ck(self.L.pm_record_buffers(self.h, which, C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
        return a.value, b.value, c.value, n.value

    # -- Mode B: photon-map build + k-nearest-photon search ---------------------------------------------------
    def knn_set_curve(self, curve):
        """0 = Morton (Z-order) sort key, 1 = Hilbert (default)."""
        self._ck(self.L.pm_knn_set_curve(self.h, curve))

    def knn_set_batched(self, on=True):
        self._ck(self.L.pm_knn_set_batched(self.h, on))

    def knn_build(self, which=0):
        self._ck(self.L.pm_knn_build(self.h, which))

    def knn_build_points(self, which, pos4, power4, n, records=False):
        """pos4 / power4: DEVICE float4 arrays (torch tensors or addresses) that outlive the map.  records=True: the rows
        are photon records (e.g. all-gathered from several GPUs); the surface map then keeps wall hits only."""
        self._ck(self.L.pm_knn_build_points(self.h, which, _ptr(pos4), _ptr(power4), n, records))

    def knn_size(self, which=0):
        n, lv = C.c_int64(), C.c_int32()
        self._ck(self.L.pm_knn_size(self.h, which, C.byref(n), C.byref(lv)))
        return n.value, lv.value

    def knn_query(self, which, queries4, nq, k, max_r2, idx, d2, cnt):
        self._ck(self.L.pm_knn_query(self.h, which, _ptr(queries4), nq, k, max_r2, _ptr(idx), _ptr(d2), _ptr(cnt)))

    def knn_radiance(self, which, queries4, nq, k, max_r2, rgb4):
        self._ck(self.L.pm_knn_radiance(self.h, which, _ptr(queries4), nq, k, max_r2, _ptr(rgb4)))

    def render_knn(self, w, h, t=0.0, media=False, k=50, max_r2=float("inf"), w_surface=1.0, w_volume=1.0, rgba=None, rgbf=None,
                   y0=0, y1=None, y_step=1):
        """Mode B frame into caller-owned DEVICE buffers: rows y0, y0+y_step, ... < y1."""
        y1 = h if y1 is None else y1
        self._ck(self.L.pm_render_knn_rows(self.h, t, media, w, h, y0, y1, y_step, k, max_r2, w_surface, w_volume, _ptr(rgba), _ptr(rgbf)))

    def knn_radiance_cone(self, queries4, nq, k, sq_radius, exposure, rgb4):
        """Legacy fixed-radius cone-filter estimate; queries4 = (x, y, z, wall id)."""
        self._ck(self.L.pm_knn_radiance_cone(self.h, _ptr(queries4), nq, k, sq_radius, exposure, _ptr(rgb4)))

    def knn_sorted(self, which, n):
        keys = np.empty(n, np.uint32); perm = np.empty(n, np.uint32)
        self._ck(self.L.pm_knn_sorted_host(self.h, which, _ptr(keys), _ptr(perm), n))
        return keys, perm

    def knn_level(self, which, level):
        n = C.c_int64()
        self._ck(self.L.pm_knn_level_host(self.h, which, level, C.byref(n), None))
        boxes = np.empty((6, n.value), np.float32)
        self._ck(self.L.pm_knn_level_host(self.h, which, level, C.byref(n), _ptr(boxes)))
        return boxes

    # -- stages 3-5 -------------------------------------------------------------------------------------
    def render_device(self, w, h, t=0.0, interp=False, media=False, rgba=None, rgbf=None, y0=0, y1=None):
        """photonMappingCuda() (simplePBO.cpp:130) into caller-owned DEVICE buffers (torch tensors or addresses)."""
        y1 = h if y1 is None else y1
        self._ck(self.L.pm_render(self.h, t, interp, media, w, h, y0, y1, _ptr(rgba), _ptr(rgbf)))

    def render(self, w, h, t=0.0, interp=False, media=False, want_u8=True, want_f32=True, out_u8=None, out_f32=None):
        """Render through HOST buffers; returns (rgba uint8 [h,w,4], rgbf float32 [h,w,4])."""
        u8 = out_u8 if out_u8 is not None else (np.empty((h, w, 4), np.uint8) if want_u8 else None)
        f32 = out_f32 if out_f32 is not None else (np.empty((h, w, 4), np.float32) if want_f32 else None)
        self._ck(self.L.pm_render_host(self.h, t, interp, media, w, h, _ptr(u8), _ptr(f32)))
        return u8, f32

    def frame(self, w, h, t=0.0, emit=True, interp=False, media=False, out_u8=None, out_f32=None):
        """One display() frame (callbacksPBO.cpp:47-101) through host buffers."""
        self._ck(self.L.pm_frame_host(self.h, t, emit, interp, media, w, h, _ptr(out_u8), _ptr(out_f32)))

    def frame_async(self, w, h, out_u8, t=0.0, emit=True, interp=False, media=False):
        """pm_frame_host_async: enqueue a frame whose uchar4 image lands in out_u8 (pinned host memory); returns the ticket."""
        tk = C.c_int64()
        self._ck(self.L.pm_frame_host_async(self.h, t, emit, interp, media, w, h, _ptr(out_u8), C.byref(tk)))
        return tk.value

    def frame_wait(self, ticket):
        self._ck(self.L.pm_frame_wait(self.h, ticket))

    def frame_device(self, w, h, rgba=None, rgbf=None, t=0.0, emit=True, interp=False, media=False):
        """pm_frame_device: one pipelined frame into DEVICE buffers (this rank's row band; peer barrier when ranks are connected)."""
        self._ck(self.L.pm_frame_device(self.h, t, emit, interp, media, w, h, _ptr(rgba), _ptr(rgbf)))

    def set_trace_sms(self, sms):
        self._ck(self.L.pm_set_trace_sms(self.h, sms))

    def launch_count(self):
        return self.L.pm_launch_count(self.h)

    def enable_timing(self, on=True):
        self._ck(self.L.pm_enable_timing(self.h, on))

    def timings(self):
        """{kernel name: (total ms, launches)} since the last call (synchronises the stream)."""
        n = self.L.pm_kernel_count()
        ms = np.zeros(n, np.float64); cnt = np.zeros(n, np.int64)
        self._ck(self.L.pm_get_timings(self.h, _ptr(ms), _ptr(cnt)))
        return {self.L.pm_kernel_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n)}


    # -- multi-GPU: peers (pm_peer_*) ---------------------------------------------------------
    def peer_export(self):
        """64-byte CUDA IPC handle of this context's exchange block (accumulators + flags), as bytes."""
        buf = C.create_string_buffer(64)
        self._ck(self.L.pm_peer_export(self.h, buf))
        return buf.raw

    def peer_connect(self, rank, world, handles):
        """handles: list of `world` 64-byte handles (entry [rank] is ignored).  pm_build_map then sums the accumulators of all
        ranks over peer memory; every rank must step through the frames in lockstep."""
        blob = b"".join(h if h is not None else b"\0" * 64 for h in handles)
        assert len(blob) == 64 * world
        self._ck(self.L.pm_peer_connect(self.h, rank, world, blob))

    def peer_disconnect(self):
        self._ck(self.L.pm_peer_disconnect(self.h))

    def peer_barrier(self):
        self._ck(self.L.pm_peer_barrier(self.h))

    def peer_status(self):
        self._ck(self.L.pm_peer_status(self.h))

    def peer_set_timeout(self, seconds):
        self._ck(self.L.pm_peer_set_timeout(self.h, seconds))

    def shared_alloc(self, nbytes):
        """(device pointer, 64-byte handle) of device memory other ranks can map with shared_open."""
        p, buf = C.c_void_p(), C.create_string_buffer(64)
        self._ck(self.L.pm_shared_alloc(self.h, nbytes, C.byref(p), buf))
        return p.value, buf.raw

    def shared_open(self, handle):
        p = C.c_void_p()
        self._ck(self.L.pm_shared_open(self.h, handle, C.byref(p)))
        return p.value

    def shared_close(self, ptr, opened):
        self._ck(self.L.pm_shared_close(self.h, C.c_void_p(ptr), opened))

    def set_row_band(self, y0=-1, y1=-1):
        self._ck(self.L.pm_set_row_band(self.h, y0, y1))

    def trace_profile(self, on=True):
        self._ck(self.L.pm_trace_profile(self.h, on))

    def get_trace_profile(self):
        """[CTAs, 48] uint64 %globaltimer stamps of the last trace launch (see pmb200.h)."""
        out = np.zeros(48 * 1024, np.uint64)
        n = C.c_int64()
        self._ck(self.L.pm_get_trace_profile_host(self.h, _ptr(out), out.size, C.byref(n)))
        return out[:n.value].reshape(-1, 48)

    def selftest_fdiv(self, pairs, seed=1):
        """(violations, accepted): the trace kernel's branch-free wall division against the IEEE one (see pmb200.h)."""
        bad, acc = C.c_uint64(), C.c_uint64()
        self._ck(self.L.pm_selftest_fdiv(self.h, pairs, seed, C.byref(bad), C.byref(acc)))
        return bad.value, acc.value


class PhotonGroup:
    """pm_group: n GPUs behind one object, for a single-process host program (one worker thread per GPU inside the library)."""

    def __init__(self, devices, n_photons=10000, scene=None):
        self.L = lib()
        g = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        rc = self.L.pm_group_create(C.byref(g), arr, len(devices))
        if rc != 0:
            raise PmError({-4: "no CUDA device: pmb200 has no CPU fallback"}.get(rc, f"pm_group_create failed ({rc})"))
        self.g, self.n = g, len(devices)
        self.set_photon_count(n_photons)
        if scene is not None:
            self.set_scene(scene)

    def _ck(self, rc):
        if rc != 0:
            raise PmError(f"pmb200 group error {rc}: {self.L.pm_group_last_error(self.g).decode()}")

    def close(self):
        if getattr(self, "g", None):
            self.L.pm_group_destroy(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rank(self, r):
        """The context of rank r as a (borrowed) PhotonMapper, for inspection."""
        return PhotonMapper(_borrowed=self.L.pm_group_context(self.g, r), n_photons=self.n_photons)

    def set_scene(self, scene):
        self._ck(self.L.pm_group_set_scene(self.g, C.byref(scene)))

    def set_photon_count(self, n):
        self._ck(self.L.pm_group_set_photon_count(self.g, n))
        self.n_photons = n

    def set_energy_scale(self, s):
        self._ck(self.L.pm_group_set_energy_scale(self.g, s))

    def init_random_numbers(self):
        self._ck(self.L.pm_group_init_random_table(self.g))

    def frame(self, w, h, t=0.0, emit=True, interp=False, media=False, out_u8=None, out_f32=None):
        self._ck(self.L.pm_group_frame_host(self.g, t, emit, interp, media, w, h, _ptr(out_u8), _ptr(out_f32)))

    def frame_async(self, w, h, out_u8, t=0.0, emit=True, interp=False, media=False):
        tk = C.c_int64()
        self._ck(self.L.pm_group_frame_host_async(self.g, t, emit, interp, media, w, h, _ptr(out_u8), C.byref(tk)))
        return tk.value

    def frame_wait(self, ticket):
        self._ck(self.L.pm_group_frame_wait(self.g, ticket))


# -- the reference's three launchers, verbatim names (process-global default context) --------------------
def launch_init_random_numbers_kernel():
    lib().launch_init_random_numbers_kernel()


def launch_emit_photons_kernel(pos, image_width, image_height, animTime, interpolateFlag, participatingMediaFlag):
    lib().launch_emit_photons_kernel(_ptr(pos), image_width, image_height, animTime, interpolateFlag, participatingMediaFlag)


def launch_photon_mapping_kernel(pos, image_width, image_height, animTime, interpolateFlag, participatingMediaFlag):
    lib().launch_photon_mapping_kernel(_ptr(pos), image_width, image_height, animTime, interpolateFlag, participatingMediaFlag)


# -- the older variant's two launchers (kernelPBO.cu:295, :317); dead in the reference (declarations and calls commented out) ---
def launch_render_kernel(pos, image_width, image_height, time, pixelData):
    lib().launch_render_kernel(_ptr(pos), image_width, image_height, time, _ptr(pixelData))


def launch_kernel(pos, image_width, image_height, time, numPhotons=None, photons=None):
    lib().launch_kernel(_ptr(pos), image_width, image_height, time, _ptr(numPhotons), _ptr(photons))
