"""Multi-GPU exchange through the C-ABI: pm_group_* (one process, one worker thread per GPU, peer access) and
pm_peer_* (one process per GPU, CUDA IPC).  The accumulators are exact integers, so the photon map, the summed accumulators
and the uchar4 / float frames must be BIT-identical to the single-GPU result for every number of ranks.

A group may name the same device twice: the whole protocol (double-buffered accumulators, in-kernel signal / wait, sum over
peer pointers, row bands copied into one host frame) then runs on ONE GPU -- that is what the driver's single-GPU box
exercises.  The cases with distinct devices and the CUDA-IPC processes need >= 2 GPUs (gpurun --gpus 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, W, H = 200000, 320, 200


def _gpus():
    import torch
    return torch.cuda.device_count()


def _single(pm, media, t, frames=1):
    m = pm.PhotonMapper(n_photons=N)
    sc = pm.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
    m.set_scene(sc)
    m.init_random_numbers()
    out = []
    for f in range(frames):
        u8 = np.zeros((H, W, 4), np.uint8); f32 = np.zeros((H, W, 4), np.float32)
        m.frame(W, H, t + 0.1 * f, True, False, media, out_u8=u8, out_f32=f32)
        out.append((u8, f32, m.get_map(), m.get_accumulators()))
    m.close()
    return out


def _group_devices():
    cases = [[0, 0], [0, 0, 0]]
    if _gpus() >= 2:
        cases += [[0, 1]]
    if _gpus() >= 4:
        cases += [[0, 1, 2, 3]]
    return cases


@pytest.mark.parametrize("media", [False, True])
def test_group_frames_bit_identical_to_single_gpu(pm, media):
    """pm_group_frame_host over several ranks == pm_frame_host on one context: map, summed accumulators, both frames; three
    consecutive frames (the accumulator buffers alternate)."""
    ref = _single(pm, media, 0.3, frames=3)
    for devices in _group_devices():
        g = pm.PhotonGroup(devices, n_photons=N)
        sc = pm.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
        g.set_scene(sc)
        g.init_random_numbers()
        for f in range(3):
            u8 = np.zeros((H, W, 4), np.uint8); f32 = np.zeros((H, W, 4), np.float32)
            g.frame(W, H, 0.3 + 0.1 * f, True, False, media, out_u8=u8, out_f32=f32)
            ru8, rf32, rmap, racc = ref[f]
            for r in range(len(devices)):
                ctx = g.rank(r)
                ctx.peer_status()
                assert ctx.get_map().tobytes() == rmap.tobytes(), (devices, f, r)
                assert np.array_equal(ctx.get_accumulators(), racc), (devices, f, r)
            assert np.array_equal(u8, ru8), (devices, f)
            assert f32.tobytes() == rf32.tobytes(), (devices, f)
        g.close()


def test_group_pipelined_frames(pm):
    """pm_group_frame_host_async / pm_group_frame_wait: frames submitted two deep land complete and bit-identical."""
    import torch
    ref = _single(pm, True, 0.0, frames=4)
    for devices in _group_devices()[:1] + _group_devices()[2:]:
        g = pm.PhotonGroup(devices, n_photons=N)
        sc = pm.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
        g.set_scene(sc)
        g.init_random_numbers()
        bufs = [torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
        pending = []
        got = []
        for f in range(4):
            tk = g.frame_async(W, H, bufs[f & 1], 0.1 * f, True, False, True)
            if pending:
                ptk, pf = pending.pop()
                g.frame_wait(ptk)
                got.append(bufs[pf & 1].numpy().copy())
            pending.append((tk, f))
        ptk, pf = pending.pop()
        g.frame_wait(ptk)
        got.append(bufs[pf & 1].numpy().copy())
        for f in range(4):
            assert np.array_equal(got[f], ref[f][0]), (devices, f)
        g.close()


def test_vox_section_is_exchanged_when_touched(pm):
    """A back wall at z = 5 lies off the map boundary: its hits go to the acc_vox section, which the exchange reads only when a
    rank flags it.  The summed map must still equal the single-GPU one."""
    def scene():
        sc = pm.default_scene(sz_img=H)
        sc.planes[4][1] = 5.0
        return sc
    m = pm.PhotonMapper(n_photons=N, scene=scene())
    m.init_random_numbers()
    for _ in range(2):          # the medium walk's MWC stream continues from frame to frame
        m.emit(0.0, True)
    rmap, racc = m.get_map(), m.get_accumulators()
    m.close()
    assert np.any(racc[5 * 32 * 32 * 4: 5 * 32 * 32 * 4 + 32 * 32 * 32 * 3] != 0), "the scene must reach the acc_vox section"
    g = pm.PhotonGroup([0, 0] if _gpus() < 2 else [0, 1], n_photons=N, scene=scene())
    g.init_random_numbers()
    u8 = np.zeros((H, W, 4), np.uint8)
    for _ in range(2):
        g.frame(W, H, 0.0, True, False, True, out_u8=u8)
    for r in range(2):
        assert g.rank(r).get_map().tobytes() == rmap.tobytes()
        assert np.array_equal(g.rank(r).get_accumulators(), racc)
    g.close()


def test_peer_wait_times_out_instead_of_hanging(pm):
    """A rank whose peers never arrive reports PM_ERR_STATE after the timeout; the GPU is not left spinning."""
    a, b = pm.PhotonMapper(n_photons=1000), pm.PhotonMapper(n_photons=1000)
    import ctypes as C
    L = pm.lib()
    members = (C.c_void_p * 2)(a.h, b.h)
    assert L.pm_peer_connect_local(a.h, 0, 2, members) == 0
    assert L.pm_peer_connect_local(b.h, 1, 2, members) == 0
    a.peer_set_timeout(0.2)
    a.init_random_numbers()
    a.clear_map(); a.trace(0.0); a.build_map()      # rank 1 never traces
    with pytest.raises(pm.PmError, match="timed out"):
        a.peer_status()
    a.peer_disconnect(); b.peer_disconnect()
    a.close(); b.close()


@pytest.mark.parametrize("world", [2, 4])
def test_ipc_ranks_bit_identical(pm, world):
    """One process per GPU, exchange blocks mapped through CUDA IPC handles (what bench.py does under torchrun)."""
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    ref = _single(pm, True, 0.0, frames=3)
    out = os.path.join(ROOT, "gpurun_out", "ipc_test_%d.npz" % world)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    env = dict(os.environ, PMB200_TEST_OUT=out)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "mp_peer_worker.py"), str(N), str(W), str(H), "3"]
    subprocess.check_call(cmd, env=env, cwd=ROOT, timeout=600)
    z = np.load(out)
    for f in range(3):
        assert np.array_equal(z["u8_%d" % f], ref[f][0]), f
        assert z["map_%d" % f].tobytes() == ref[f][2].tobytes(), f


def test_frame_device_pipeline_matches_serial_frames(pm):
    """pm_frame_device (clear + trace on one stream, map build on a second, render on a third; accumulators rotating through three
    buffers, gather tables through two)
    gives the frames of the one-stream pm_frame_host, bit for bit, also when several frames are enqueued back to back."""
    import torch
    ref = _single(pm, True, 0.0, frames=5)
    m = pm.PhotonMapper(n_photons=N)
    sc = pm.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
    m.set_scene(sc)
    m.init_random_numbers()
    u8 = [torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda") for _ in range(5)]
    f32 = [torch.zeros((H, W, 4), dtype=torch.float32, device="cuda") for _ in range(5)]
    torch.cuda.synchronize()
    for f in range(5):
        m.frame_device(W, H, rgba=u8[f], rgbf=f32[f], t=0.1 * f, emit=True, interp=False, media=True)
    m.sync()
    for f in range(5):
        assert np.array_equal(u8[f].cpu().numpy(), ref[f][0]), f
        assert f32[f].cpu().numpy().tobytes() == ref[f][1].tobytes(), f
    assert m.get_map().tobytes() == ref[4][2].tobytes()
    m.close()
