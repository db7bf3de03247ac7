// pm_map.cu -- photon-map finalisation (sm_100a): exact accumulators -> float voxel map -> gather tables.
//
// build_map_kernel applies, once per voxel, what the reference does once per photon (storePhoton's direct
// deposit PMK:1176-1177, splatEnergy's 6x6 wall stencil PMK:1074-1144 with weight 0.05/dist PMK:1068, and
// storeVolumePhoton PMK:1158): the splat is linear in the deposited energy, so
//     map[x] = vox[x] + sum_p [x on slab(p)] ( hit[p][x] + sum_{v != x, x in window(v)} 0.05 * hit[p][v] / |v - x| ).
// Sums are taken in double from exact int64 inputs, so the map is deterministic and independent of how the
// photons were sharded across GPUs.
//
// build_tables_kernel tabulates the reference's two gathers (integrate PMK:314-389 and
// integrateVolumePhotons PMK:831-870) for every integer voxel coordinate they can see, summing in the
// reference's loop order with the reference's FP32 operations, so a table lookup is bit-identical to the
// per-pixel loops it replaces (1400 voxel reads per pixel in media mode).
#include "pm_kernels.cuh"

namespace pm {

__device__ __forceinline__ void slab_of(int id, int x, int y, int z, int &on, int &a, int &b) {
  switch (id) {
    case 0: on = x == PM_GRID_N - 1; a = y; b = z; break;
    case 2: on = x == 0;             a = y; b = z; break;
    case 1: on = y == 0;             a = x; b = z; break;
    case 3: on = y == PM_GRID_N - 1; a = x; b = z; break;
    default: on = z == PM_GRID_N - 1; a = x; b = y; break;
  }
}

// splat weights 0.05f * (1.0f / dist) for the 6 x 6 offsets (da, db) in [-3, 2]^2, rounded as the reference rounds them
// (IEEE float sqrt and division), times 1 / kHitScale; [3][3] (da = db = 0) is the direct deposit, weight 1
struct StencilWeights { double w[6][6]; };
static StencilWeights make_stencil_weights() {
  StencilWeights s;
  for (int da = -3; da <= 2; da++)
    for (int db = -3; db <= 2; db++) {
      if (da == 0 && db == 0) { s.w[3][3] = 1.0 / kHitScale; continue; }
      volatile float dist = sqrtf((float)(da * da + db * db));
      volatile float inv = 1.0f / dist;
      s.w[da + 3][db + 3] = ((double)0.05f / kHitScale) * (double)inv;
    }
  return s;
}

// One voxel of the map.  kWalls: how many wall ids (0..kWalls-1) can have this voxel on their slab.
template <int kWalls>
__device__ __forceinline__ void build_voxel(const long long *__restrict__ acc, const StencilWeights &sw, float energy_scale,
                                            float *__restrict__ grid, int x, int y, int z) {
  const int v = (x * PM_GRID_N + y) * PM_GRID_N + z;
  const long long *vox = acc + kAccHitEntries + 3 * v;
  long long grey = 0;
#pragma unroll
  for (int r = 0; r < kGreyReplicas; r++) grey += acc[kAccHitEntries + kAccVoxEntries + r * PM_GRID_VOXELS + v];
  double s0 = (double)(vox[0] + grey) / kVoxScale, s1 = (double)(vox[1] + grey) / kVoxScale, s2 = (double)(vox[2] + grey) / kVoxScale;
#pragma unroll 1
  for (int id = 0; id < kWalls; id++) {
    int on, a, b;
    slab_of(id, x, y, z, on, a, b);
    if (!on) continue;
    const long long *hit = acc + id * PM_GRID_N * PM_GRID_N * 4;
    // sources (a', b') whose window [a'-3, a'+3) x [b'-3, b'+3) contains (a, b): a' in [a-2, a+3], likewise b'.  The loads of one
    // row of six sources are issued together (a tap-by-tap loop exposed the L2 latency 36 times: 21 us for this kernel in round 1);
    // the sums run in (a', b') order whatever the clipping, so the result does not depend on it
#pragma unroll 1
    for (int ap = a - 2; ap <= a + 3; ap++) {
      if ((unsigned)ap >= (unsigned)PM_GRID_N) continue;
      longlong2 rg[6], bg[6];
#pragma unroll
      for (int c = 0; c < 6; c++) {
        const int bp = b - 2 + c;
        const bool ok = (unsigned)bp < (unsigned)PM_GRID_N;
        const longlong2 *h = reinterpret_cast<const longlong2 *>(hit + (ap * PM_GRID_N + (ok ? bp : 0)) * 4);
        rg[c] = ok ? h[0] : make_longlong2(0, 0); bg[c] = ok ? h[1] : make_longlong2(0, 0);
      }
#pragma unroll
      for (int c = 0; c < 6; c++) {
        const long long h0 = rg[c].x + bg[c].y, h1 = rg[c].y + bg[c].y, h2 = bg[c].x + bg[c].y;
        if ((h0 | h1 | h2) == 0) continue;
        // source (ap, bp) reaches (a, b) at offset (a - ap, b - bp) in [-3, 2]^2; the weight only depends on |offset|
        const double wgt = sw.w[a - ap + 3][5 - c];
        s0 += (double)h0 * wgt; s1 += (double)h1 * wgt; s2 += (double)h2 * wgt;
      }
    }
  }
  double sc = (double)energy_scale;
  grid[3 * v + 0] = (float)(s0 * sc); grid[3 * v + 1] = (float)(s1 * sc); grid[3 * v + 2] = (float)(s2 * sc);
}

// Blocks [0, 256): one thread per voxel, a warp = the 32 z of one (x, y) column, so whether the column lies on the slab of a side
// wall (ids 0..3) is warp-uniform and the 36-tap stencil runs with every lane.  The back wall's slab (id 4, z = 31) is one lane
// per warp there (round 1: 5 active lanes per instruction), so those 1024 voxels are left to blocks [256, 264), whose warps run
// along y at fixed x.
constexpr int kBuildMapBlocks = PM_GRID_VOXELS / 128 + PM_GRID_N * PM_GRID_N / 128;
__global__ void __launch_bounds__(128) build_map_kernel(const long long *__restrict__ acc, const __grid_constant__ StencilWeights sw,
                                                        float energy_scale, float *__restrict__ grid) {
  if (blockIdx.x < PM_GRID_VOXELS / 128) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    int z = v % PM_GRID_N, y = (v / PM_GRID_N) % PM_GRID_N, x = v / (PM_GRID_N * PM_GRID_N);
    if (z == PM_GRID_N - 1) return;
    build_voxel<4>(acc, sw, energy_scale, grid, x, y, z);
  } else {
    int t = (blockIdx.x - PM_GRID_VOXELS / 128) * blockDim.x + threadIdx.x;
    build_voxel<5>(acc, sw, energy_scale, grid, t / PM_GRID_N, t % PM_GRID_N, PM_GRID_N - 1);
  }
}

__device__ __forceinline__ v3 ld3(const float *__restrict__ grid, int i, int j, int k) {
  const float *g = grid + 3 * ((i * PM_GRID_N + j) * PM_GRID_N + k);
  return V(g[0], g[1], g[2]);
}

// vol_table: one block per (wx, group of five wy) stages the <= 6 x 10 x 32 voxels its 175 entries can see in shared memory and
// every thread sums its own clipped window from there, in the reference's i, j, k order (round 1 read the 216 voxels of every entry
// through L1/L2: 30 us).  surf_table: one thread per entry straight from the map (<= 36 reads).
constexpr int kVolTileY = 5, kVolTilesY = kVolN / kVolTileY;                       // 7 groups of 5
constexpr int kVolBlocks = kVolN * kVolTilesY;                                     // 245
constexpr int kTableThreads = 192;                                                 // >= kVolTileY * kVolN = 175
constexpr int kSurfBlocksPerWall = (kSurfN * kSurfN + kTableThreads - 1) / kTableThreads;   // 8
constexpr int kSurfBlocks = PM_MAX_PLANES * kSurfBlocksPerWall;
constexpr int kTileX = 6, kTileYRows = kVolTileY + 5;
__global__ void __launch_bounds__(kTableThreads) build_tables_kernel(const float *__restrict__ grid, float4 *__restrict__ vol_table,
                                                                     float4 *__restrict__ surf_table) {
  __shared__ float tile[kTileX * kTileYRows * PM_GRID_N * 3];   // 23 040 B
  if (blockIdx.x < kVolBlocks) {   // integrateVolumePhotons: clipped [w-3, w+3) within [1, 31), loop order i, j, k
    const int bx = (int)blockIdx.x, tx = (int)threadIdx.x;
    const int wx = bx / kVolTilesY + kVolLo, wy0 = (bx % kVolTilesY) * kVolTileY + kVolLo;
    int x0, x1, y0, y1, d0, d1;
    window(wx, 3, 1, PM_GRID_N - 1, x0, x1);
    window(wy0, 3, 1, PM_GRID_N - 1, y0, d0);
    window(wy0 + kVolTileY - 1, 3, 1, PM_GRID_N - 1, d1, y1);
    const int ny = y1 - y0, row = PM_GRID_N * 3;                 // <= 10 rows of 96 floats, contiguous in the map for a fixed x
    const int per_x = ny * row, total = (x1 - x0) * per_x;
#pragma unroll 8
    for (int t = threadIdx.x; t < total; t += kTableThreads) {
      const int i = t / per_x, u = t - i * per_x;
      tile[i * kTileYRows * row + u] = grid[((x0 + i) * PM_GRID_N + y0) * row + u];
    }
    __syncthreads();
    const int wy = wy0 + tx / kVolN, wz = tx % kVolN + kVolLo;
    if (tx >= kVolTileY * kVolN) return;
    int mny, mxy, mnz, mxz;
    window(wy, 3, 1, PM_GRID_N - 1, mny, mxy);
    window(wz, 3, 1, PM_GRID_N - 1, mnz, mxz);
    v3 rgb = V(0.0f, 0.0f, 0.0f);
    const int nz = mxz - mnz;   // <= 6
    for (int i = x0; i < x1; i++)
      for (int j = mny; j < mxy; j++) {
        const float *r = tile + ((i - x0) * kTileYRows + (j - y0)) * row + 3 * mnz;
        float v[18];
#pragma unroll
        for (int q = 0; q < 18; q++) v[q] = q < 3 * nz ? r[q] : 0.0f;    // one batch of shared-memory loads, then the adds in k order
#pragma unroll
        for (int k = 0; k < 6; k++)
          if (k < nz) rgb = add(rgb, V(v[3 * k], v[3 * k + 1], v[3 * k + 2]));
      }
    vol_table[((wx - kVolLo) * kVolN + (wy - kVolLo)) * kVolN + (wz - kVolLo)] = make_float4(rgb.x, rgb.y, rgb.z, 0.0f);
    return;
  }
  // integrate + computeEnergy: energy += map * 0.0005f over the clipped 6x6 wall window, outer/inner loop = (a, b).  Eight blocks
  // per wall, each staging the wall's 32 x 32 slab of the map.
  const int sb = (int)blockIdx.x - kVolBlocks, id = sb / kSurfBlocksPerWall;
  for (int t = threadIdx.x; t < PM_GRID_N * PM_GRID_N; t += blockDim.x) {
    const int a = t / PM_GRID_N, b = t % PM_GRID_N;
    v3 g;
    switch (id) {
      case 0: g = ld3(grid, PM_GRID_N - 1, a, b); break;
      case 2: g = ld3(grid, 0, a, b); break;
      case 1: g = ld3(grid, a, 0, b); break;
      case 3: g = ld3(grid, a, PM_GRID_N - 1, b); break;
      default: g = ld3(grid, a, b, PM_GRID_N - 1); break;
    }
    tile[3 * t] = g.x; tile[3 * t + 1] = g.y; tile[3 * t + 2] = g.z;
  }
  __syncthreads();
  const int w = (sb % kSurfBlocksPerWall) * kTableThreads + (int)threadIdx.x;
  if (w >= kSurfN * kSurfN) return;
  const int wb = w % kSurfN + kSurfLo, wa = w / kSurfN + kSurfLo;
  int mna, mxa, mnb, mxb;
  window(wa, 3, 0, PM_GRID_N, mna, mxa);
  window(wb, 3, 0, PM_GRID_N, mnb, mxb);
  v3 e = V(0.0f, 0.0f, 0.0f);
  const int nb = mxb - mnb;   // <= 6
  for (int a = mna; a < mxa; a++) {
    const float *g = tile + 3 * (a * PM_GRID_N + mnb);
    float v[18];
#pragma unroll
    for (int q = 0; q < 18; q++) v[q] = q < 3 * nb ? g[q] : 0.0f;
#pragma unroll
    for (int b = 0; b < 6; b++)
      if (b < nb) e = add(e, mul(V(v[3 * b], v[3 * b + 1], v[3 * b + 2]), 0.0005f));
  }
  const int t = id * kSurfN * kSurfN + w;
  surf_table[t] = make_float4(e.x, e.y, e.z, 0.0f);
}

cudaError_t preload_map_kernels() {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, build_map_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, build_tables_kernel);
  return e;
}

cudaError_t launch_build_map(const long long *acc, float energy_scale, float *grid, cudaStream_t st) {
  static const StencilWeights sw = make_stencil_weights();
  build_map_kernel<<<kBuildMapBlocks, 128, 0, st>>>(acc, sw, energy_scale, grid);
  return cudaGetLastError();
}

cudaError_t launch_build_tables(const float *grid, float4 *vol_table, float4 *surf_table, cudaStream_t st) {
  build_tables_kernel<<<kVolBlocks + kSurfBlocks, kTableThreads, 0, st>>>(grid, vol_table, surf_table);
  return cudaGetLastError();
}

}  // namespace pm
