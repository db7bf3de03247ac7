"""world_size-2 gloo test (CPU) of the multi-GPU host logic in pmb200.dist: photon sharding, the exact int64
all-reduce of the accumulators, and the row-band frame gather.  The per-rank compute is done by the oracle (there is no
GPU here); what is under test is the partition + exchange logic that bench.py runs over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fixed_point_accumulators(rec):
    """Order-independent int64 accumulation of a record list keyed by clamped voxel (2^24 scale for surface
    deposits, 2^36 for volume deposits) -- the property the product's accumulators rely on."""
    acc = np.zeros((32 * 32 * 32, 3), np.int64)
    loc = rec["loc"].astype(np.float64)
    vx = np.clip(np.trunc(((loc[:, 0] + 1.5) / 3.0) * 32), 0, 31).astype(np.int64)
    vy = np.clip(np.trunc(((loc[:, 1] + 1.5) / 3.0) * 32), 0, 31).astype(np.int64)
    vz = np.clip(np.trunc((loc[:, 2] / 6.0) * 32), 0, 31).astype(np.int64)
    v = (vx * 32 + vy) * 32 + vz
    scale = np.where(rec["kind"] == 1, 2.0 ** 36, 2.0 ** 24)[:, None]
    np.add.at(acc, v, np.rint(rec["energy"].astype(np.float64) * scale).astype(np.int64))
    return acc


def _worker(rank, world, port, n, w, h, out_dir):
    sys.path.insert(0, ROOT)
    import pmb200
    from pmb200 import dist as pd
    from oracle.oraclelib import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    sc = orc.default_scene(sz_img=h)
    table, st = orc.mwc_table(n)
    first, last = pd.photon_shard(n, rank, world)
    # NB: the medium draws of photon i sit at stream position 9*i, so a shard starts from the state advanced by 9*first
    _, st_shard = orc.mwc_draws(9 * first, *st)
    _, rec, _ = orc.emit(sc, table, first, last, 0.0, True, rng=st_shard, max_records=32 * n, want_grid=False)
    acc = torch.from_numpy(fixed_point_accumulators(rec))
    pd.allreduce_accumulators(acc)
    # every rank renders its band of the frame of the FULL map and gathers
    grid, _, _ = orc.emit(sc, table, 0, n, 0.0, True, rng=st)
    y0, y1 = pd.row_band(h, rank, world)
    band, _ = orc.render(sc, grid, w, h, 0.0, False, True, y0=y0, y1=y1, want_u8=False)
    frame = torch.zeros((h, w, 3), dtype=torch.float32)
    frame[y0:y1] = torch.from_numpy(band[y0:y1])
    pd.gather_frame(frame, y0, y1)
    if rank == 0:
        np.save(os.path.join(out_dir, "acc.npy"), acc.numpy())
        np.save(os.path.join(out_dir, "frame.npy"), frame.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_partition_helpers():
    sys.path.insert(0, ROOT)
    from pmb200 import dist as pd      # pure Python: does not load libpmb200.so
    for n, world in ((10, 3), (16777216, 8), (7, 8), (1000003, 4)):
        shards = [pd.photon_shard(n, r, world) for r in range(world)]
        assert shards[0][0] == 0 and shards[-1][1] == n
        assert all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
    assert [pd.row_band(1080, r, 8) for r in (0, 7)] == [(0, 135), (945, 1080)]
    with pytest.raises(ValueError):
        pd.row_band(1080, 0, 7)
    bands = [pd.row_band_uneven(1080, r, 7) for r in range(7)]
    assert bands[0][0] == 0 and bands[-1][1] == 1080 and all(bands[i][1] == bands[i + 1][0] for i in range(6))
    assert max(b - a for a, b in bands) - min(b - a for a, b in bands) <= 1


def test_two_rank_exchange_is_exact(oracle, tmp_path):
    n, w, h = 6001, 40, 30
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, n, w, h, str(tmp_path)), nprocs=2, join=True)
    sc = oracle.default_scene(sz_img=h)
    table, st = oracle.mwc_table(n)
    grid, rec, _ = oracle.emit(sc, table, 0, n, 0.0, True, rng=st, max_records=32 * n)
    assert np.array_equal(np.load(tmp_path / "acc.npy"), fixed_point_accumulators(rec))
    img, _ = oracle.render(sc, grid, w, h, 0.0, False, True, want_u8=False)
    assert np.load(tmp_path / "frame.npy").tobytes() == img.tobytes()
