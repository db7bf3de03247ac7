"""pmb200.dist -- multi-GPU plumbing (one process per GPU, torch.distributed; NCCL over NVLink on the GPU box,
gloo in the CPU tests).  The path shards naturally:

  * photons: rank r traces the contiguous index range photon_shard(n, r, world) of the SAME random table, so the
    union of the shards is exactly the single-GPU photon set (the MWC stream is index-addressed);
  * one exchange: the int64 fixed-point accumulators are summed across ranks (all-reduce).  Integer addition is
    associative, so the summed accumulators -- and the float photon map built from them -- are bit-identical for
    every world size;
  * pixels: rank r renders the row band row_band(h, r, world); bands are all-gathered into the full frame.
"""
import torch
import torch.distributed as dist


def photon_shard(n_photons, rank, world):
    """Contiguous, disjoint, exhaustive split of [0, n_photons)."""
    return n_photons * rank // world, n_photons * (rank + 1) // world


def row_band(height, rank, world):
    """Contiguous row band [y0, y1) of rank `rank`; height must divide evenly so bands all-gather in place."""
    if height % world:
        raise ValueError("frame height %d does not split into %d equal row bands" % (height, world))
    rows = height // world
    return rank * rows, (rank + 1) * rows


def row_band_uneven(height, rank, world):
    """Contiguous row band of rank `rank` for any height (bands differ by at most one row)."""
    return height * rank // world, height * (rank + 1) // world


def connect_peers(mapper):
    """One process per GPU: exchange the CUDA IPC handles of every rank's exchange block with torch.distributed and connect
    them (pm_peer_connect), so that pm_build_map sums the accumulators over NVLink peer memory itself.  No-op for one rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return False
    world, rank = dist.get_world_size(), dist.get_rank()
    handles = [None] * world
    dist.all_gather_object(handles, mapper.peer_export())
    mapper.peer_connect(rank, world, handles)
    dist.barrier()      # nobody starts signalling before every rank has mapped every block
    return True


def shared_frame(mapper, nbytes):
    """A device buffer on rank 0 that every rank maps (CUDA IPC): ranks render their row band straight into it over NVLink.
    Returns (device pointer valid on this rank, opened) -- pass both to mapper.shared_close()."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return mapper.shared_alloc(nbytes)[0], False
    box = [None]
    ptr = None
    if dist.get_rank() == 0:
        ptr, box[0] = mapper.shared_alloc(nbytes)
    dist.broadcast_object_list(box, src=0)
    if dist.get_rank() != 0:
        return mapper.shared_open(box[0]), True
    return ptr, False


def shared_host_frames(count, nbytes, world):
    """`count` page-locked host frames of `nbytes` that EVERY rank process can write (POSIX shared memory, registered with CUDA in
    each process), plus an int64 [count, world] array of completion words in the same segment.  Every rank copies its own row band
    of a frame into it over its own PCIe link; rank 0 reads the assembled frame.  Single process: plain pinned tensors.
    Returns (frames, done_words, keepalive)."""
    import numpy as np
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(count)], np.zeros((count, 1), np.int64), None
    from multiprocessing import shared_memory, resource_tracker
    stride = (nbytes + 4095) // 4096 * 4096
    total = count * stride + 4096
    box = [None]
    if dist.get_rank() == 0:
        shm = shared_memory.SharedMemory(create=True, size=total)
        box[0] = shm.name
    dist.broadcast_object_list(box, src=0)
    if dist.get_rank() != 0:
        shm = shared_memory.SharedMemory(name=box[0])
        try:    # only the creating process owns (and unlinks) the segment
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:
            pass
    buf = np.ndarray((total,), np.uint8, buffer=shm.buf)
    rc = torch.cuda.cudart().cudaHostRegister(buf.ctypes.data, total, 0)
    if int(rc) != 0:
        raise RuntimeError("cudaHostRegister failed: %r" % (rc,))
    frames = [buf[i * stride:i * stride + nbytes] for i in range(count)]
    done = np.ndarray((count, world), np.int64, buffer=shm.buf, offset=count * stride)
    if dist.get_rank() == 0:
        done[:] = 0
    dist.barrier()

    class _Keep:
        def __init__(self, shm, ptr):
            self.shm, self.ptr = shm, ptr

        def close(self):
            try:
                torch.cuda.cudart().cudaHostUnregister(self.ptr)
            except Exception:
                pass
            try:
                self.shm.close()
                if dist.get_rank() == 0:
                    self.shm.unlink()
            except Exception:
                pass
    return frames, done, _Keep(shm, buf.ctypes.data)


class _RawCuda:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr, n, typestr="<i8"):
    """Zero-copy torch view of a raw device allocation owned by libpmb200 (e.g. the accumulators)."""
    return torch.as_tensor(_RawCuda(ptr, n, typestr), device="cuda")


def allreduce_accumulators(acc):
    """Sum the exact accumulators across ranks (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def gather_frame(frame, y0, y1):
    """All-gather the row bands of a [H, W, C] frame in place (every rank ends with the whole frame)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return frame
    band = frame[y0:y1].reshape(-1)
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(frame.view(-1), band)
    else:
        parts = [torch.empty_like(band) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, band.clone())
        frame.view(-1).copy_(torch.cat(parts))
    return frame


def allgather_records(pos_ptr, pow_ptr, count):
    """All-gather one record set (SoA float4 pos_meta / power_index, `count` rows on this rank) from every rank.
    Returns (pos4, power4, total): contiguous [total, 4] float32 device tensors in rank order -- the input of
    PhotonMapper.knn_build_points(..., records=True).  Single process: zero-copy views of the local buffers."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return device_tensor(pos_ptr, count * 4, "<f4").view(count, 4), device_tensor(pow_ptr, count * 4, "<f4").view(count, 4), count
    counts = torch.zeros(world, dtype=torch.int64, device="cuda")
    counts[dist.get_rank()] = count
    dist.all_reduce(counts)
    counts = [int(c) for c in counts.tolist()]
    cmax, total = max(counts), sum(counts)
    out = []
    for ptr in (pos_ptr, pow_ptr):
        local = torch.zeros((cmax, 4), dtype=torch.float32, device="cuda")
        if count:
            local[:count] = device_tensor(ptr, count * 4, "<f4").view(count, 4)
        gathered = torch.empty((world, cmax, 4), dtype=torch.float32, device="cuda")
        dist.all_gather_into_tensor(gathered.view(-1), local.view(-1))
        out.append(torch.cat([gathered[r, :counts[r]] for r in range(world)]) if min(counts) < cmax else gathered.view(-1, 4))
    return out[0], out[1], total
