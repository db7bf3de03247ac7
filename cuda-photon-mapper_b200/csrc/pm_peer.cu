// pm_peer.cu -- the path's one multi-GPU exchange, done by our own kernel over NVLink peer memory (sm_100a).
//
// each GPU traces 1/N of the photons into its own exact int64 accumulators (pm_layout.h); the photon
// map is built from their SUM.  The reference has no counterpart (single device, simplePBO.cpp:191-195).  Round 1 summed
// them with an NCCL all-reduce: 0.075 ms at 2 GPUs, 0.199 ms at 8 -- latency-bound, longer than the trace itself.  Here every
// rank PULLS the other ranks' accumulators straight out of their memory (cudaDeviceEnablePeerAccess inside one process,
// CUDA IPC handles between processes) inside one kernel that also does the synchronisation:
//
//   signal : block 0 stores this frame's sequence number into every peer's arrive[my rank] slot (st.release.sys).  The
//            kernel is stream-ordered after the trace and the count fold, so the accumulators are complete.
//   wait   : one thread per block polls the LOCAL arrive[] slots (ld.acquire.sys) until every peer has signalled.
//   reduce : out[e] = sum over ranks of acc_r[e], 16-byte volatile loads, in rank order (integer sums: any order gives the
//            same bits).  The acc_vox section (786 KB of the 1.2 MB, all zero unless a wall lies off the map boundary) is
//            only read from ranks whose vox_touched flag is set.
//
// The accumulators rotate through three buffers, so that a frame's trace never waits for the previous frame's exchange (the
// host side runs clear + trace on one stream, exchange + map build on a second and the render on a third).  A rank
// clears buffer b again three frames later, after its OWN reduce of frame f+1 has completed (an event wait in pm_clear_map,
// two frames old by then): that reduce saw every peer's signal f+1, which each peer sends only after finishing its reduce of
// frame f -- its reads of buffer b -- in stream order.  So no "consumed" handshake is needed.  A peer that never arrives trips
// a timeout (error word) instead of hanging the GPU.
#include <cstdlib>
#include "pm_kernels.cuh"

namespace pm {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ld_peer_2x64(const long long *p, long long &a, long long &b) {
  asm volatile("ld.volatile.global.v2.s64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long peer_clock_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// signal every peer on `channel`, then wait until every peer's signal for `seq` has arrived (thread 0 of the block)
__device__ __forceinline__ bool signal_and_wait(const PeerView &pv, int channel, uint32_t seq, bool do_signal) {
  if (do_signal && (int)threadIdx.x < pv.world && (int)threadIdx.x != pv.rank) {
    __threadfence_system();
    st_release_sys(pv.hdr[threadIdx.x]->arrive[channel] + pv.rank, seq);
  }
  __shared__ int ok;
  if (threadIdx.x == 0) {
    int good = 1;
    const unsigned long long t0 = peer_clock_ns();
    for (int p = 0; p < pv.world && good; p++) {
      if (p == pv.rank) continue;
      const uint32_t *slot = pv.hdr[pv.rank]->arrive[channel] + p;
      unsigned spins = 0;
      while ((int32_t)(ld_acquire_sys(slot) - seq) < 0) {   // wrap-safe "arrived < seq"
        if ((++spins & 1023u) == 0u && peer_clock_ns() - t0 > pv.timeout_ns) { good = 0; break; }
      }
    }
    if (!good) atomicExch(&pv.hdr[pv.rank]->error, 1u + (uint32_t)channel);
    ok = good;
  }
  __syncthreads();
  return ok != 0;
}

// seq_or_nowait: the frame's sequence number; kNoWait when the signal + wait ran as a kernel of its own in front of this one
// (PMB200_PEER_SPLIT, see launch_peer_reduce)
constexpr uint32_t kNoWait = 0xffffffffu;
__global__ void __launch_bounds__(256) peer_reduce_kernel(const __grid_constant__ PeerView pv, int buf, uint32_t seq_or_nowait,
                                                          long long *__restrict__ out) {
  if (seq_or_nowait != kNoWait) signal_and_wait(pv, 0, seq_or_nowait, blockIdx.x == 0);   // on a timeout the sums are garbage; the host sees the error word
  __shared__ unsigned vox_mask;
  if (threadIdx.x == 0) {
    unsigned m = 0;
    for (int p = 0; p < pv.world; p++) m |= (*(volatile const uint32_t *)(pv.acc[p] + (size_t)buf * kAccStride + kAccEntries) ? 1u : 0u) << p;
    vox_mask = m;
  }
  __syncthreads();
  const unsigned vm = vox_mask;
  const int pairs = kAccEntries / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += gridDim.x * blockDim.x) {
    const int e = 2 * i;
    const bool vox = e >= kAccHitEntries && e < kAccHitEntries + kAccVoxEntries;
    long long s0 = 0, s1 = 0;
#pragma unroll 8
    for (int p = 0; p < pv.world; p++) {
      if (vox && !((vm >> p) & 1u)) continue;
      long long a, b;
      ld_peer_2x64(pv.acc[p] + (size_t)buf * kAccStride + e, a, b);
      s0 += a; s1 += b;
    }
    out[e] = s0; out[e + 1] = s1;
  }
}

// stand-alone device-side barrier between the ranks (e.g. "every rank's band has landed in rank 0's frame buffer")
__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerView pv, int channel, uint32_t seq) {
  signal_and_wait(pv, channel, seq, true);   // on a timeout the host sees the error word (a reduce that follows sums garbage)
}

cudaError_t preload_peer_kernels() {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, peer_reduce_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, peer_barrier_kernel);
  return e;
}

cudaError_t launch_peer_reduce(const PeerView &pv, int buf, uint32_t seq, long long *out, int blocks, cudaStream_t st) {
  static_assert(kAccEntries % 2 == 0 && kAccHitEntries % 2 == 0 && kAccVoxEntries % 2 == 0, "sections must be 16-byte aligned");
  // PMB200_PEER_SPLIT=1: signal + wait as a one-warp kernel in front of the reduce, so that no SM is held by the reduce's spinning
  // blocks at the moment the next frame's persistent trace kernel wants its SMs.  One measurement at 8 GPUs (132 trace CTAs): end to
  // end 0.152 -> 0.148 ms, but the device-resident frame 0.125 -> 0.158 ms -- not understood, so off by default.
  static const bool split = getenv("PMB200_PEER_SPLIT") != nullptr && getenv("PMB200_PEER_SPLIT")[0] == '1';
  if (split && seq != kNoWait) {
    peer_barrier_kernel<<<1, 32, 0, st>>>(pv, 0, seq);
    peer_reduce_kernel<<<blocks, 256, 0, st>>>(pv, buf, kNoWait, out);
  } else {
    peer_reduce_kernel<<<blocks, 256, 0, st>>>(pv, buf, seq, out);
  }
  return cudaGetLastError();
}
cudaError_t launch_peer_barrier(const PeerView &pv, uint32_t seq, cudaStream_t st) {
  peer_barrier_kernel<<<1, 32, 0, st>>>(pv, 1, seq);
  return cudaGetLastError();
}

}  // namespace pm
