import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pmb200 as pm
N, W, H = 200000, 320, 200
def single(frames):
    m = pm.PhotonMapper(n_photons=N)
    sc = pm.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
    m.set_scene(sc); m.init_random_numbers()
    out = []
    for f in range(frames):
        u8 = np.zeros((H, W, 4), np.uint8)
        m.frame(W, H, 0.1 * f, True, False, True, out_u8=u8)
        out.append((u8, m.get_map()))
    m.close()
    return out
ref = single(5)
for mode in ("default-stream", "own-stream", "own-stream", "default-stream"):
    m = pm.PhotonMapper(n_photons=N)
    if mode == "own-stream":
        st = torch.cuda.Stream(); m.set_stream(st.cuda_stream)
    sc = pm.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
    m.set_scene(sc); m.init_random_numbers()
    u8 = [torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda") for _ in range(5)]
    torch.cuda.synchronize()
    maps = []
    for f in range(5):
        m.frame_device(W, H, rgba=u8[f], t=0.1 * f, emit=True, interp=False, media=True)
        if mode == "sync-each":
            m.sync(); maps.append(m.get_map())
    m.sync()
    for f in range(5):
        d = u8[f].cpu().numpy()
        print(mode, "frame", f, "pixels differing:", int((d != ref[f][0]).any(-1).sum()), "map equal" if maps and maps[f].tobytes() == ref[f][1].tobytes() else ("map differs" if maps else ""))
    m.close()
