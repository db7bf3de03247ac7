"""One rank of tests/test_gpu_multi.py::test_ipc_ranks_bit_identical (launched by torchrun): traces its photon shard, connects
to the other ranks through CUDA IPC handles (exchanged with torch.distributed), renders its row band straight into rank 0's
frame buffer (mapped through IPC as well); rank 0 saves the frames and maps."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pmb200  # noqa: E402
from pmb200 import dist as pmdist  # noqa: E402


def main():
    n, w, h, frames = (int(x) for x in sys.argv[1:5])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = pmb200.PhotonMapper(device=local, n_photons=n)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    sc = pmb200.default_scene(sz_img=h); sc.cam_ox = -(w - h) / 2.0
    m.set_scene(sc)
    m.set_photon_range(*pmdist.photon_shard(n, rank, world))
    m.init_random_numbers()
    pmdist.connect_peers(m)
    frame_ptr, opened = pmdist.shared_frame(m, w * h * 4)
    y0, y1 = pmdist.row_band_uneven(h, rank, world)
    out = {}
    for f in range(frames):
        m.clear_map(); m.trace(0.1 * f, media=True); m.build_map()
        m.render_device(w, h, 0.1 * f, False, True, rgba=frame_ptr, y0=y0, y1=y1)
        m.peer_barrier()                      # every band has landed in rank 0's buffer
        m.peer_status()
        if rank == 0:
            out["u8_%d" % f] = pmdist.device_tensor(frame_ptr, w * h * 4, "|u1").cpu().numpy().reshape(h, w, 4)
            out["map_%d" % f] = m.get_map()
        m.peer_barrier()                      # rank 0 has read the frame before anybody overwrites it
    m.sync()
    dist.barrier()
    if rank == 0:
        np.savez(os.environ["PMB200_TEST_OUT"], **out)
    m.shared_close(frame_ptr, opened)
    m.peer_disconnect()
    m.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
