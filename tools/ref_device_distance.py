#!/usr/bin/env python
"""tools/ref_device_distance.py -- how far the product's photon map / frame is from the reference's own DEVICE arithmetic:
the reference CUDA kernel with atomic deposits and -fmad=false (oracle/_ref/libpmref_cuda_atomic_<N>.so, build_ref.sh P4) on the
same table, configs 1 and 2, media off (deterministic up to the order of the float atomics) and on (the reference's medium walk
draws from a racy global MWC state, so that part is statistical)."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, pmb200 as pm


def distances(n, w, h, cfg1, media):
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpmref_cuda_atomic_%d.so" % n))
    assert L.refcu_capacity() == n
    sc = pm.default_scene(sz_img=h)
    nobj = np.array([2, 5], np.int32)
    if cfg1:
        sc.n_spheres = 1; nobj[0] = 1
        for i, off in enumerate([1e9, -1.5, -1e9, 1e9, 1e9]):
            sc.planes[i][1] = off
    m = pm.PhotonMapper(n_photons=n, scene=sc)
    m.set_energy_scale(10000.0 / n)
    m.init_random_numbers()
    table = m.get_random_table()
    pl = np.array([[sc.planes[i][0], sc.planes[i][1]] for i in range(5)], np.float32)
    sp = np.array([[sc.spheres[i][j] for j in range(4)] for i in range(3)], np.float32)
    li = np.array([sc.light[i] for i in range(3)], np.float32)
    assert L.refcu_set_scene(nobj.ctypes.data_as(C.c_void_p), pl.ctypes.data_as(C.c_void_p), sp.ctypes.data_as(C.c_void_p), li.ctypes.data_as(C.c_void_p)) == 0
    assert L.refcu_set_table(table.ctypes.data_as(C.c_void_p), n) == 0
    assert L.refcu_set_szimg(h) == 0
    L.refcu_emit(C.c_float(0.0), 0, int(media))
    ref_u8 = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    L.refcu_render(C.c_void_p(ref_u8.data_ptr()), w, h, C.c_float(0.0), 0, int(media))
    torch.cuda.synchronize()
    rgrid = np.zeros((32, 32, 32, 3), np.float32)
    assert L.refcu_get_grid(rgrid.ctypes.data_as(C.c_void_p)) == 0
    m.emit(0.0, media=media)
    ours = m.get_map() / np.float32(10000.0 / n)          # unscaled, like the reference's
    u8, f32 = m.render(w, h, 0.0, False, media)
    m.set_map((rgrid * np.float32(10000.0 / n)).astype(np.float32))
    u8r, f32r = m.render(w, h, 0.0, False, media)         # the reference's map through the same (oracle-exact) renderer
    m.close()
    a, b = f32[..., :3].astype(np.float64), f32r[..., :3].astype(np.float64)
    mse = ((a - b) ** 2).mean()
    out = {"map_rel_l1": float(np.abs(ours.astype(np.float64) - rgrid).sum() / np.abs(rgrid).sum()),
           "map_energy_ratio": float(ours.sum(dtype=np.float64) / rgrid.sum(dtype=np.float64)),
           "frame_rel_l1": float(np.abs(a - b).sum() / np.abs(b).sum()),
           "frame_psnr_db": float(10 * np.log10(b.max() ** 2 / mse)) if mse > 0 else float("inf")}
    return out


if __name__ == "__main__":
    for name, n, w, h, cfg1 in (("config1", 65536, 256, 256, True), ("config2", 1048576, 1024, 1024, False)):
        for media in (False, True):
            print(name, "media" if media else "surface", json.dumps(distances(n, w, h, cfg1, media)))
