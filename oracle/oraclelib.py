"""oracle/oraclelib.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/libpm_oracle.so (the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpm_oracle.so")

RECORD_DTYPE = np.dtype([("type", "<i4"), ("id", "<i4"), ("index", "<i4"), ("kind", "<i4"),
                         ("loc", "<f4", 3), ("dir", "<f4", 3), ("energy", "<f4", 3)])


class Scene(C.Structure):
    """Mirror of pm_scene (include/pmb200_types.h)."""
    _fields_ = [("n_spheres", C.c_int32), ("n_planes", C.c_int32),
                ("spheres", (C.c_float * 4) * 3), ("planes", (C.c_float * 2) * 5),
                ("light", C.c_float * 3), ("sz_img", C.c_int32),
                ("cam_ox", C.c_float), ("cam_oy", C.c_float), ("animate", C.c_int32)]

    def copy(self):
        s = Scene()
        C.memmove(C.byref(s), C.byref(self), C.sizeof(Scene))
        return s

    def spheres_np(self):
        return np.array([[self.spheres[i][j] for j in range(4)] for i in range(3)], np.float32)

    def planes_np(self):
        return np.array([[self.planes[i][j] for j in range(2)] for i in range(5)], np.float32)


def available():
    return os.path.exists(LIB_PATH)


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


class Oracle:
    def __init__(self):
        self.lib = L = C.CDLL(LIB_PATH)
        L.pmo_mwc_next.restype = C.c_uint32
        L.pmo_rand_float.restype = C.c_float
        L.pmo_rand_float.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        L.pmo_mwc_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.pmo_mwc_skip.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        L.pmo_set_shadow_grid64.argtypes = [C.c_void_p]
        L.pmo_philox_table.argtypes = [C.c_uint64, C.c_void_p, C.c_int]
        L.pmo_position_objects.argtypes = [C.POINTER(Scene), C.c_float]
        L.pmo_emit.restype = C.c_long
        L.pmo_emit.argtypes = [C.POINTER(Scene), C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        L.pmo_render.argtypes = [C.POINTER(Scene), C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.pmo_raytrace.restype = C.c_int
        L.pmo_eye_geometry.argtypes = [C.POINTER(Scene), C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.pmo_morton30_many.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.pmo_hilbert30_many.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.pmo_stable_sort_perm.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.pmo_knn_bruteforce.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pmo_knn_cone_estimate.argtypes = [C.c_void_p] * 8 + [C.c_float, C.c_long, C.c_int, C.c_void_p]
        L.pmo_knn_estimate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]

    # -- scene -----------------------------------------------------------------------------------
    def default_scene(self, sz_img=512, animate=1):
        s = Scene()
        self.lib.pmo_scene_default(C.byref(s))
        s.sz_img = sz_img
        s.animate = animate
        return s

    def position_objects(self, scene, t):
        s = scene.copy()
        self.lib.pmo_position_objects(C.byref(s), C.c_float(t))
        return s

    # -- RNG --------------------------------------------------------------------------------------
    def mwc_table(self, n, w=6548, z=316):
        """Returns (table[n,3], (w,z) state after the 3n draws)."""
        st = np.array([w, z], np.uint32)
        tab = np.zeros((n, 3), np.float32)
        self.lib.pmo_mwc_table(C.c_void_p(st.ctypes.data), C.c_void_p(st.ctypes.data + 4), _p(tab), n)
        return tab, (int(st[0]), int(st[1]))

    def mwc_skip(self, n, w=6548, z=316):
        """(w, z) after n serial draws."""
        st = np.array([w, z], np.uint32)
        self.lib.pmo_mwc_skip(C.c_void_p(st.ctypes.data), C.c_void_p(st.ctypes.data + 4), C.c_long(n))
        return int(st[0]), int(st[1])

    def mwc_draws(self, n, w=6548, z=316):
        st = np.array([w, z], np.uint32)
        out = np.zeros(n, np.uint32)
        for i in range(n):
            out[i] = self.lib.pmo_mwc_next(C.c_void_p(st.ctypes.data), C.c_void_p(st.ctypes.data + 4))
        return out, (int(st[0]), int(st[1]))

    def philox(self, ctr, key):
        c = np.array(ctr, np.uint32); k = np.array(key, np.uint32); o = np.zeros(4, np.uint32)
        self.lib.pmo_philox4x32_10(_p(c), _p(k), _p(o))
        return o

    def philox_table(self, n, seed=0x5EED):
        tab = np.zeros((n, 3), np.float32)
        self.lib.pmo_philox_table(C.c_uint64(seed), _p(tab), n)
        return tab

    # -- stage 1 -----------------------------------------------------------------------------------
    def emit(self, scene, table, n0, n1, t=0.0, media=False, rng=(6548, 316), grid=None, max_records=0,
             want_grid=True, shadow64=None):
        """Returns (grid, records, rng_state_after).  grid is accumulated into (not cleared) if given.  shadow64: a float64
        [32,32,32,3] array that receives the same deposits summed in double."""
        self.lib.pmo_set_shadow_grid64(_p(shadow64) if shadow64 is not None else None)
        table = np.ascontiguousarray(table, np.float32)
        st = np.array(rng, np.uint32)
        if grid is None and want_grid:
            grid = np.zeros((32, 32, 32, 3), np.float32)
        rec = np.zeros(max_records, RECORD_DTYPE) if max_records else None
        cnt = self.lib.pmo_emit(C.byref(scene), t, _p(table), n0, n1, int(media),
                                C.c_void_p(st.ctypes.data), C.c_void_p(st.ctypes.data + 4),
                                _p(grid) if want_grid else None, _p(rec), max_records)
        self.lib.pmo_set_shadow_grid64(None)
        if rec is not None:
            assert cnt <= max_records, (cnt, max_records)
            rec = rec[:cnt]
        return grid, rec, (int(st[0]), int(st[1]))

    # -- stages 3-5 --------------------------------------------------------------------------------
    def render(self, scene, grid, w, h, t=0.0, interp=False, media=False, y0=0, y1=None, want_u8=True):
        grid = np.ascontiguousarray(grid, np.float32)
        y1 = h if y1 is None else y1
        rgb = np.zeros((h, w, 3), np.float32)
        u8 = np.zeros((h, w, 4), np.uint8) if want_u8 else None
        self.lib.pmo_render(C.byref(scene), t, _p(grid), w, h, y0, y1, int(interp), int(media), _p(rgb), _p(u8))
        return rgb, u8

    def eye_geometry(self, scene, w, h, t=0.0):
        """(hit[h,w,8] = (hit, type, id, px, py, pz, 0, 0), march[h,w,10,3]) -- eye-ray geometry for the Mode B oracle."""
        hit = np.zeros((h, w, 8), np.float32); march = np.zeros((h, w, 10, 3), np.float32)
        self.lib.pmo_eye_geometry(C.byref(scene), t, w, h, 0, h, _p(hit), _p(march))
        return hit, march

    # -- probes --------------------------------------------------------------------------------------
    def voxel(self, p):
        p = np.asarray(p, np.float32); v = np.zeros(3, np.int32)
        self.lib.pmo_voxel(_p(p), _p(v))
        return v

    def raytrace(self, scene, ray, org):
        ray = np.asarray(ray, np.float32); org = np.asarray(org, np.float32)
        d = C.c_float(0); ty = C.c_int(0); ix = C.c_int(0)
        hit = self.lib.pmo_raytrace(C.byref(scene), _p(ray), _p(org), C.byref(d), C.byref(ty), C.byref(ix))
        return bool(hit), d.value, ty.value, ix.value

    def integrate_volume(self, grid, p):
        grid = np.ascontiguousarray(grid, np.float32)
        p = np.asarray(p, np.float32); c = np.zeros(3, np.float32)
        self.lib.pmo_integrate_volume(_p(grid), _p(p), _p(c))
        return c

    def gather(self, grid, p, type_, id_, interp=False):
        grid = np.ascontiguousarray(grid, np.float32)
        p = np.asarray(p, np.float32); c = np.zeros(3, np.float32)
        self.lib.pmo_gather(_p(grid), _p(p), C.c_int(type_), C.c_int(id_), C.c_int(int(interp)), _p(c))
        return c

    # -- Mode B oracle (oracle/knn_oracle.c) ------------------------------------------------------------------
    def morton30(self, pos4):
        pos4 = np.ascontiguousarray(pos4, np.float32)
        keys = np.empty(pos4.shape[0], np.uint32)
        self.lib.pmo_morton30_many(_p(pos4), pos4.shape[0], _p(keys))
        return keys

    def hilbert30(self, pos4):
        pos4 = np.ascontiguousarray(pos4, np.float32)
        keys = np.empty(pos4.shape[0], np.uint32)
        self.lib.pmo_hilbert30_many(_p(pos4), pos4.shape[0], _p(keys))
        return keys

    def stable_sort_perm(self, keys):
        keys = np.ascontiguousarray(keys, np.uint32)
        perm = np.empty(keys.shape[0], np.uint32)
        self.lib.pmo_stable_sort_perm(_p(keys), keys.shape[0], _p(perm))
        return perm

    def knn_bruteforce(self, pos4, queries4, k, max_r2=np.inf):
        pos4 = np.ascontiguousarray(pos4, np.float32); queries4 = np.ascontiguousarray(queries4, np.float32)
        nq = queries4.shape[0]
        idx = np.empty((nq, k), np.int32); d2 = np.empty((nq, k), np.float32); cnt = np.empty(nq, np.int32)
        self.lib.pmo_knn_bruteforce(_p(pos4), pos4.shape[0], _p(queries4), nq, k, C.c_float(max_r2), _p(idx), _p(d2), _p(cnt))
        return idx, d2, cnt

    def knn_estimate(self, power4, idx, d2, cnt, volume):
        power4 = np.ascontiguousarray(power4, np.float32)
        nq, k = idx.shape
        out = np.empty((nq, 3), np.float32)
        self.lib.pmo_knn_estimate(_p(power4), _p(np.ascontiguousarray(idx)), _p(np.ascontiguousarray(d2)), _p(np.ascontiguousarray(cnt)),
                                  nq, k, int(volume), _p(out))
        return out

    def knn_cone_estimate(self, pos_meta4, dir4, power4, idx, d2, cnt, wall, normals, exposure):
        nq, k = idx.shape
        out = np.empty((nq, 4), np.float32)
        a = [np.ascontiguousarray(x, t) for x, t in ((pos_meta4, np.float32), (dir4, np.float32), (power4, np.float32), (idx, np.int32),
                                                      (d2, np.float32), (cnt, np.int32), (wall, np.int32), (normals, np.float32))]
        self.lib.pmo_knn_cone_estimate(*[_p(x) for x in a], C.c_float(exposure), nq, k, _p(out))
        return out
