"""oracle/refhost.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libpmref_host.so.

That library is the reference's own device routines (photonMappingKernel.cu:1-1521) compiled as host
C++ by oracle/build_ref.sh.  Only tests/, tests/golden/make_golden.py and bench.py's reference /
cpu_baseline leg may import this module; the product path never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpmref_host.so")

RECORD_DTYPE = np.dtype([("type", "<i4"), ("id", "<i4"), ("index", "<i4"), ("kind", "<i4"),
                         ("loc", "<f4", 3), ("dir", "<f4", 3), ("energy", "<f4", 3)])

DEFAULT_PLANES = np.array([[0, 1.5], [1, -1.5], [0, -1.5], [1, 1.5], [2, 6.0]], dtype=np.float32)
DEFAULT_SPHERES = np.array([[1.0, -1.0, 1.0, 0.4], [-0.6, -1.0, 4.5, 0.4], [0.0, 0.0, 1.5, 1.0]], dtype=np.float32)
DEFAULT_LIGHT = np.array([0.0, 1.4, 3.5], dtype=np.float32)


def available():
    return os.path.exists(LIB_PATH)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class RefHost:
    """Thin, stateful wrapper (the reference keeps all state in globals; so does this)."""

    def __init__(self):
        self.lib = L = C.CDLL(LIB_PATH)
        L.ref_emit.restype = C.c_long
        L.ref_emit.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_long]
        L.ref_emit_omp.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
        L.ref_render_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
        L.ref_render_f32_omp.argtypes = L.ref_render_f32.argtypes
        L.ref_render_u8.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_float, C.c_int, C.c_int]
        L.ref_set_rng.argtypes = [C.c_uint, C.c_uint]
        L.ref_get_random.restype = C.c_uint
        L.ref_rand_float.restype = C.c_float
        L.ref_rand_float.argtypes = [C.c_float]
        L.ref_position_objects.argtypes = [C.c_float]
        L.ref_raytrace.restype = C.c_int
        self.reset()

    # -- state -------------------------------------------------------------------------------
    def reset(self):
        self.set_scene()
        self.set_rng(6548, 316)
        self.clear_grid()

    def capacity(self):
        return self.lib.ref_capacity()

    def omp_threads(self):
        return self.lib.ref_omp_threads()

    def omp_set_threads(self, n):
        """Use n OpenMP threads from now on (a launcher may have exported OMP_NUM_THREADS=1)."""
        if hasattr(self.lib, "ref_omp_set_threads"):
            self.lib.ref_omp_set_threads(int(n))

    def set_scene(self, nr_objects=(2, 5), planes=None, spheres=None, light=None, sz_img=512):
        n = np.array(nr_objects, dtype=np.int32)
        p = np.ascontiguousarray(DEFAULT_PLANES if planes is None else planes, dtype=np.float32)
        s = np.ascontiguousarray(DEFAULT_SPHERES if spheres is None else spheres, dtype=np.float32)
        l = np.ascontiguousarray(DEFAULT_LIGHT if light is None else light, dtype=np.float32)
        self.lib.ref_set_scene(n.ctypes.data_as(C.c_void_p), _fp(p), _fp(s), _fp(l), C.c_int(sz_img))

    def get_scene(self):
        n = np.zeros(2, np.int32); p = np.zeros((5, 2), np.float32); s = np.zeros((3, 4), np.float32)
        l = np.zeros(3, np.float32); sz = C.c_int(0)
        self.lib.ref_get_scene(n.ctypes.data_as(C.c_void_p), _fp(p), _fp(s), _fp(l), C.byref(sz))
        return dict(nr_objects=n, planes=p, spheres=s, light=l, sz_img=sz.value)

    def position_objects(self, t):
        self.lib.ref_position_objects(C.c_float(t))

    def set_rng(self, w, z):
        self.lib.ref_set_rng(w, z)

    def get_rng(self):
        a = np.zeros(2, np.uint32)
        self.lib.ref_get_rng(a.ctypes.data_as(C.c_void_p))
        return int(a[0]), int(a[1])

    def get_random(self):
        return self.lib.ref_get_random()

    def rand_float(self, mx):
        return self.lib.ref_rand_float(mx)

    def init_table(self, n):
        self.lib.ref_init_table(C.c_int(n))

    def init_table_native(self, n):
        self.lib.ref_init_table_native(C.c_int(n))

    def set_table(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self.lib.ref_set_table(_fp(xyz), C.c_int(xyz.shape[0]))

    def get_table(self, n):
        a = np.zeros((n, 3), np.float32)
        self.lib.ref_get_table(_fp(a), C.c_int(n))
        return a

    def clear_grid(self):
        self.lib.ref_clear_grid()

    def get_grid(self):
        g = np.zeros((32, 32, 32, 3), np.float32)
        self.lib.ref_get_grid(_fp(g))
        return g

    def set_grid(self, g):
        g = np.ascontiguousarray(g, dtype=np.float32)
        assert g.shape == (32, 32, 32, 3)
        self.lib.ref_set_grid(_fp(g))

    # -- the two hot loops ---------------------------------------------------------------------
    def emit(self, n0, n1, t=0.0, interp=False, media=False, max_records=0):
        rec = np.zeros(max_records, RECORD_DTYPE) if max_records else None
        cnt = self.lib.ref_emit(n0, n1, t, int(interp), int(media),
                                rec.ctypes.data_as(C.c_void_p) if rec is not None else None, max_records)
        if rec is not None:
            assert cnt <= max_records, (cnt, max_records)
            return rec[:cnt]
        return cnt

    def emit_omp(self, n0, n1, t=0.0, interp=False, media=False):
        self.lib.ref_emit_omp(n0, n1, t, int(interp), int(media))

    def render_f32(self, w, h, t=0.0, interp=False, media=False, ox=0.0, oy=0.0, omp=False):
        img = np.zeros((h, w, 3), np.float32)
        fn = self.lib.ref_render_f32_omp if omp else self.lib.ref_render_f32
        fn(img.ctypes.data_as(C.c_void_p), w, h, ox, oy, t, int(interp), int(media))
        return img

    def render_u8(self, w, h, t=0.0, interp=False, media=False):
        img = np.zeros((h, w, 4), np.uint8)
        self.lib.ref_render_u8(img.ctypes.data_as(C.c_void_p), w, h, t, int(interp), int(media))
        return img

    # -- single-routine probes -------------------------------------------------------------------
    def voxel(self, p):
        p = np.asarray(p, np.float32); v = np.zeros(3, np.int32)
        self.lib.ref_voxel(_fp(p), v.ctypes.data_as(C.c_void_p))
        return v

    def raytrace(self, ray, org):
        ray = np.asarray(ray, np.float32); org = np.asarray(org, np.float32)
        d = C.c_float(0); ty = C.c_int(0); ix = C.c_int(0)
        hit = self.lib.ref_raytrace(_fp(ray), _fp(org), C.byref(d), C.byref(ty), C.byref(ix))
        return bool(hit), d.value, ty.value, ix.value

    def integrate_volume(self, p):
        p = np.asarray(p, np.float32); c = np.zeros(3, np.float32)
        self.lib.ref_integrate_volume(_fp(p), _fp(c))
        return c

    def gather(self, p, type_, id_, interp=False):
        p = np.asarray(p, np.float32); c = np.zeros(3, np.float32)
        self.lib.ref_gather(_fp(p), C.c_int(type_), C.c_int(id_), C.c_int(int(interp)), _fp(c))
        return c
