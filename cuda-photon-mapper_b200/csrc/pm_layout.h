// pm_layout.h -- HBM layout of the photon map and its derived tables (host + device).
#pragma once

#include <stdint.h>
#include "../../include/pmb200_types.h"

namespace pm {

// ---- exact accumulators (int64 fixed point; order-independent, so bit-reproducible for any GPU count) -----
// acc_hit[plane id 0..4][a][b][r,g,b,grey] : energy of wall hits whose clamped voxel lies on that wall's slab,
//                                      keyed by the two in-plane voxel coordinates (scale 2^24: every energy
//                                      the reference produces -- 10, 5/sqrt(b), -0.25 -- is exact).  A deposit with
//                                      r == g == b goes to the grey slot (one atomic instead of three)
// acc_vox[x][y][z][rgb]              : everything that is deposited straight into a voxel with unequal channels,
//                                      and the fully expanded splat of the rare off-slab hit (scale 2^36)
// acc_grey[x][y][z]                 : straight deposits with r == g == b -- every volume photon of the reference's
//                                      medium walk (scale 2^36).  Written by fold_volume_kernel only (one thread per voxel):
//                                      the walk itself counts into the replicated 32-bit vol_cnt below, so the exchanged
//                                      state stays 1.2 MB (it was 3 MB while the 64-bit atomics needed 8 replicas here).
constexpr int    kAccHitEntries = PM_MAX_PLANES * PM_GRID_N * PM_GRID_N * 4;   // 20 480
constexpr int    kAccVoxEntries = PM_GRID_VOXELS * 3;                          // 98 304
constexpr int    kGreyReplicas  = 1;
constexpr int    kAccGreyEntries = kGreyReplicas * PM_GRID_VOXELS;             // 32 768
constexpr int    kAccEntries    = kAccHitEntries + kAccVoxEntries + kAccGreyEntries;   // 151 552 int64 = 1 212 416 B
// Scratch for the medium walk: its deposits carry one of three energies (9, 8, 7 x 0.00005, a function of the step only),
// so volume_kernel COUNTS them -- 32-bit REDs, twice the L2 rate of 64-bit ones -- in vol_cnt[replica][step][voxel] and
// fold_volume_kernel adds count x quantum (exact integers) to acc_grey and clears the counts.  Not part of the accumulator
// state: always zero between launches.
constexpr int    kVolCntReplicas = 4;
constexpr int    kVolCntEntries  = kVolCntReplicas * 3 * PM_GRID_VOXELS;       // 393 216 u32 = 1.5 MB
constexpr double kHitScale      = 16777216.0;          // 2^24
constexpr double kVoxScale      = 68719476736.0;       // 2^36

// ---- multi-GPU exchange block (pm_peer.cu): ONE device allocation per context, visible to the other ranks (peer access
// inside a process, CUDA IPC between processes): a header of flags followed by kAccBuffers accumulator buffers (frames
// rotate through them, see pm_peer.cu).
constexpr int kMaxPeers = 16;
struct ExchangeHeader {
  uint32_t arrive[2][kMaxPeers];   // [channel][peer rank]: latest sequence number that peer has signalled (0 = accumulators, 1 = barrier)
  uint32_t error;                  // != 0: a wait timed out (1 + channel)
  uint32_t pad[64 - 2 * kMaxPeers - 1];
};
static_assert(sizeof(ExchangeHeader) == 256, "exchange header is 256 bytes");
// each accumulator buffer is followed by 256 bytes of per-buffer flags, cleared with it: word 0 of entry kAccEntries is
// "vox_touched" (something was deposited into the acc_vox section of this buffer)
constexpr int    kAccStride = kAccEntries + 32;
constexpr int    kAccBuffers = 3;
constexpr size_t kExchangeBytes = sizeof(ExchangeHeader) + kAccBuffers * sizeof(long long) * (size_t)kAccStride;

// ---- gather tables, rebuilt from the float photon map whenever it changes -------------------------------
// The reference's gathers depend only on the integer voxel of the query point, so their sums are tabulated
// once per map, in the reference's own summation order (=> bit-identical to summing per pixel):
//   vol_table [35][35][35] float4 : integrateVolumePhotons (PMK:831-870) for voxel coords -1..33 per axis
//   surf_table[5][37][37]  float4 : integrate (PMK:314-389) per wall id, in-plane voxel coords -2..34
constexpr int kVolLo = -1, kVolN = 35;
constexpr int kSurfLo = -2, kSurfN = 37;
constexpr int kVolTableEntries  = kVolN * kVolN * kVolN;                // 42 875
constexpr int kSurfTableEntries = PM_MAX_PLANES * kSurfN * kSurfN;      //  6 845

// ---- photon records (Mode B input): SoA float4 buffers -----------------------------------------------
// pos_meta[i]    = (x, y, z, bits(meta))     meta: seq[0:4) | kind[4] | (type+1)[5:7) | (id+1)[7:11)
// power_index[i] = (r, g, b, bits(photon index))
// dir[i]         = (dx, dy, dz, 0)
__host__ __device__ inline uint32_t pack_meta(int seq, int kind, int type, int id) {
  return (uint32_t)(seq & 15) | ((uint32_t)(kind & 1) << 4) | ((uint32_t)((type + 1) & 3) << 5) | ((uint32_t)((id + 1) & 15) << 7);
}
__host__ __device__ inline void unpack_meta(uint32_t m, int &seq, int &kind, int &type, int &id) {
  seq = (int)(m & 15u); kind = (int)((m >> 4) & 1u); type = (int)((m >> 5) & 3u) - 1; id = (int)((m >> 7) & 15u) - 1;
}

}  // namespace pm
