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

__global__ void __launch_bounds__(128) build_map_kernel(const long long *__restrict__ acc, float energy_scale, float *__restrict__ grid) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= PM_GRID_VOXELS) return;
  int z = v % PM_GRID_N, y = (v / PM_GRID_N) % PM_GRID_N, x = v / (PM_GRID_N * PM_GRID_N);
  const long long *vox = acc + kAccHitEntries + 3 * v;
  long long grey = 0;
#pragma unroll
  for (int r = 0; r < kGreyReplicas; r++) grey += acc[kAccHitEntries + kAccVoxEntries + r * PM_GRID_VOXELS + v];
  double s0 = (double)(vox[0] + grey) / kVoxScale, s1 = (double)(vox[1] + grey) / kVoxScale, s2 = (double)(vox[2] + grey) / kVoxScale;
  const double w05 = (double)0.05f / kHitScale;
  for (int id = 0; id < PM_MAX_PLANES; id++) {
    int on, a, b;
    slab_of(id, x, y, z, on, a, b);
    if (!on) continue;
    const long long *hit = acc + id * PM_GRID_N * PM_GRID_N * 4;
    // sources (a',b') whose window [a'-3, a'+3) x [b'-3, b'+3) contains (a,b): a' in [a-2, a+3]
    int a_lo = max(a - 2, 0), a_hi = min(a + 3, PM_GRID_N - 1);
    int b_lo = max(b - 2, 0), b_hi = min(b + 3, PM_GRID_N - 1);
    for (int ap = a_lo; ap <= a_hi; ap++)
      for (int bp = b_lo; bp <= b_hi; bp++) {
        const long long *h = hit + (ap * PM_GRID_N + bp) * 4;
        long long h0 = h[0] + h[3], h1 = h[1] + h[3], h2 = h[2] + h[3];
        if ((h0 | h1 | h2) == 0) continue;
        double wgt;
        if (ap == a && bp == b) wgt = 1.0 / kHitScale;   // the direct deposit
        else {
          int da = ap - a, db = bp - b;
          float dist = __fsqrt_rn((float)(da * da + db * db));
          wgt = w05 * (double)__fdiv_rn(1.0f, dist);       // 0.05f * e * (1.0f/dist), as the reference rounds it
        }
        s0 += (double)h0 * wgt; s1 += (double)h1 * wgt; s2 += (double)h2 * wgt;
      }
  }
  double sc = (double)energy_scale;
  grid[3 * v + 0] = (float)(s0 * sc); grid[3 * v + 1] = (float)(s1 * sc); grid[3 * v + 2] = (float)(s2 * sc);
}

__device__ __forceinline__ v3 ld3(const float *__restrict__ grid, int i, int j, int k) {
  const float *g = grid + 3 * ((i * PM_GRID_N + j) * PM_GRID_N + k);
  return V(g[0], g[1], g[2]);
}

__global__ void __launch_bounds__(128) build_tables_kernel(const float *__restrict__ grid, float4 *__restrict__ vol_table,
                                                           float4 *__restrict__ surf_table) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < kVolTableEntries) {   // integrateVolumePhotons: clipped [w-3, w+3) within [1, 31), loop order i, j, k
    int wz = t % kVolN + kVolLo, wy = (t / kVolN) % kVolN + kVolLo, wx = t / (kVolN * kVolN) + kVolLo;
    int mnx, mxx, mny, mxy, mnz, mxz;
    window(wx, 3, 1, PM_GRID_N - 1, mnx, mxx);
    window(wy, 3, 1, PM_GRID_N - 1, mny, mxy);
    window(wz, 3, 1, PM_GRID_N - 1, mnz, mxz);
    v3 rgb = V(0.0f, 0.0f, 0.0f);
    for (int i = mnx; i < mxx; i++) for (int j = mny; j < mxy; j++) for (int k = mnz; k < mxz; k++) rgb = add(rgb, ld3(grid, i, j, k));
    vol_table[t] = make_float4(rgb.x, rgb.y, rgb.z, 0.0f);
    return;
  }
  t -= kVolTableEntries;
  if (t >= kSurfTableEntries) return;
  // integrate + computeEnergy: energy += map * 0.0005f over the clipped 6x6 wall window, outer/inner loop = (a, b)
  int wb = t % kSurfN + kSurfLo, wa = (t / kSurfN) % kSurfN + kSurfLo, id = t / (kSurfN * kSurfN);
  int mna, mxa, mnb, mxb;
  window(wa, 3, 0, PM_GRID_N, mna, mxa);
  window(wb, 3, 0, PM_GRID_N, mnb, mxb);
  v3 e = V(0.0f, 0.0f, 0.0f);
  for (int a = mna; a < mxa; a++)
    for (int b = mnb; b < mxb; b++) {
      v3 g;
      switch (id) {
        case 0: g = ld3(grid, PM_GRID_N - 1, a, b); break;
        case 2: g = ld3(grid, 0, a, b); break;
        case 1: g = ld3(grid, a, 0, b); break;
        case 3: g = ld3(grid, a, PM_GRID_N - 1, b); break;
        default: g = ld3(grid, a, b, PM_GRID_N - 1); break;
      }
      e = add(e, mul(g, 0.0005f));
    }
  surf_table[t] = make_float4(e.x, e.y, e.z, 0.0f);
}

cudaError_t launch_build_map(const long long *acc, float energy_scale, float *grid, cudaStream_t st) {
  build_map_kernel<<<PM_GRID_VOXELS / 128, 128, 0, st>>>(acc, energy_scale, grid);
  return cudaGetLastError();
}

cudaError_t launch_build_tables(const float *grid, float4 *vol_table, float4 *surf_table, cudaStream_t st) {
  int total = kVolTableEntries + kSurfTableEntries;
  build_tables_kernel<<<(total + 127) / 128, 128, 0, st>>>(grid, vol_table, surf_table);
  return cudaGetLastError();
}

}  // namespace pm
