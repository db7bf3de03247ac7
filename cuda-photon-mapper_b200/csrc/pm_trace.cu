// pm_trace.cu -- stage 1: random-direction table + photon emission/tracing (sm_100a).
//
// Replaces init_random_numbers_kernel (PMK:1483-1498), init_photons_kernel (PMK:1503-1521) and
// emit_photons_kernel/emitPhotons (PMK:1464-1480, :1215-1375).  One thread per photon, identical FP32
// operation order to the sequential oracle (pm_math.cuh), but:
//   * the MWC generator is index-addressed (jump-ahead), so the table fill and the medium-scatter draws are
//     fully parallel yet bit-identical to the reference's serial stream;
//   * deposits go into exact int64 fixed-point accumulators keyed by wall voxel (pm_layout.h) instead of
//     ~65 racy float RMWs per photon: the 6x6 splat stencil is linear, so it is applied once per voxel in
//     pm_map.cu rather than once per photon;
//   * photon records (Mode B) are appended to SoA float4 buffers with warp-aggregated atomics.
#include "pm_kernels.cuh"

namespace pm {

// ------------------------------------------------------------------------------------------------------
// random table: thread i owns draws 3i..3i+2 of the stream that starts at (w0,z0); rand3 order x,y,z
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mwc_table_kernel(float *__restrict__ table, long long n, uint32_t w0, uint32_t z0,
                                                        const MwcJump *__restrict__ J) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Mwc s;
  uint32_t steps = (uint32_t)(3 * i);
  s.z = mwc_jump(J, 0, z0, steps);
  s.w = mwc_jump(J, 1, w0, steps);
  float x = rand_float(s, 1.0f), y = rand_float(s, 1.0f), z = rand_float(s, 1.0f);
  table[3 * i + 0] = x; table[3 * i + 1] = y; table[3 * i + 2] = z;
}

// ------------------------------------------------------------------------------------------------------
// deposits
// ------------------------------------------------------------------------------------------------------
struct Sink {
  unsigned long long *acc;       // kAccEntries, or nullptr (PM_TRACE_NO_MAP)
  float4 *rec_pos, *rec_pow, *rec_dir;
  unsigned long long *rec_count; // global append cursor
  long long rec_cap;
};

__device__ __forceinline__ void acc_add(unsigned long long *p, long long v) {
  if (v != 0) atomicAdd(p, (unsigned long long)v);
}

// warp-aggregated append: one atomic per warp per call site, lanes take consecutive slots
__device__ __forceinline__ void append_record(const Sink &sk, int seq, int kind, int type, int id, int index, v3 loc, v3 dir, v3 e) {
  unsigned mask = __activemask();
  int lane = threadIdx.x & 31;
  int leader = __ffs(mask) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(sk.rec_count, (unsigned long long)__popc(mask));
  base = __shfl_sync(mask, base, leader);
  long long slot = (long long)base + __popc(mask & ((1u << lane) - 1u));
  if (slot < sk.rec_cap) {
    sk.rec_pos[slot] = make_float4(loc.x, loc.y, loc.z, __uint_as_float(pack_meta(seq, kind, type, id)));
    sk.rec_pow[slot] = make_float4(e.x, e.y, e.z, __int_as_float(index));
    if (sk.rec_dir) sk.rec_dir[slot] = make_float4(dir.x, dir.y, dir.z, 0.0f);
  }
}

// storePhoton + splatEnergy + storeNeighborPhoton, PMK:1059-1144, :1164-1183 (type 0 = sphere: nothing is stored)
__device__ __forceinline__ void store_photon(const Sink &sk, int type, int id, v3 loc, v3 e) {
  if (!sk.acc || type == 0) return;
  int vx = clampi(voxel_x(loc.x)), vy = clampi(voxel_x(loc.y)), vz = clampi(voxel_z(loc.z));
  int a, b, on_slab;
  switch (id) {
    case 0: on_slab = vx == PM_GRID_N - 1; a = vy; b = vz; break;
    case 2: on_slab = vx == 0;             a = vy; b = vz; break;
    case 1: on_slab = vy == 0;             a = vx; b = vz; break;
    case 3: on_slab = vy == PM_GRID_N - 1; a = vx; b = vz; break;
    case 4: on_slab = vz == PM_GRID_N - 1; a = vx; b = vy; break;
    default: on_slab = -1; a = b = 0; break;
  }
  if (on_slab == 1) {   // the common case: keyed energy sum, stencil applied later
    unsigned long long *p = sk.acc + ((id * PM_GRID_N + a) * PM_GRID_N + b) * 3;
    acc_add(p + 0, __float2ll_rn(e.x * (float)kHitScale));
    acc_add(p + 1, __float2ll_rn(e.y * (float)kHitScale));
    acc_add(p + 2, __float2ll_rn(e.z * (float)kHitScale));
    return;
  }
  // rare: the clamped voxel is off the wall's slab (a wall that is not on the map boundary): expand per photon
  unsigned long long *vox = sk.acc + kAccHitEntries;
  {
    unsigned long long *p = vox + ((vx * PM_GRID_N + vy) * PM_GRID_N + vz) * 3;
    acc_add(p + 0, __double2ll_rn((double)e.x * kVoxScale));
    acc_add(p + 1, __double2ll_rn((double)e.y * kVoxScale));
    acc_add(p + 2, __double2ll_rn((double)e.z * kVoxScale));
  }
  if (on_slab < 0) return;
  int mn[3], mx[3];
  window(vx, 3, 0, PM_GRID_N, mn[0], mx[0]);
  window(vy, 3, 0, PM_GRID_N, mn[1], mx[1]);
  window(vz, 3, 0, PM_GRID_N, mn[2], mx[2]);
  int fixed_axis = (id == 0 || id == 2) ? 0 : ((id == 1 || id == 3) ? 1 : 2);
  int fixed_val = (id == 0 || id == 3 || id == 4) ? PM_GRID_N - 1 : 0;
  mn[fixed_axis] = fixed_val; mx[fixed_axis] = fixed_val + 1;
  v3 e05 = mul(e, 0.05f);
  for (int i = mn[0]; i < mx[0]; i++)
    for (int j = mn[1]; j < mx[1]; j++)
      for (int k = mn[2]; k < mx[2]; k++) {
        if (i == vx && j == vy && k == vz) continue;
        int dx = vx - i, dy = vy - j, dz = vz - k;
        float dist = __fsqrt_rn((float)(dx * dx + dy * dy + dz * dz));
        v3 t = divs(e05, dist);
        unsigned long long *p = vox + ((i * PM_GRID_N + j) * PM_GRID_N + k) * 3;
        acc_add(p + 0, __double2ll_rn((double)t.x * kVoxScale));
        acc_add(p + 1, __double2ll_rn((double)t.y * kVoxScale));
        acc_add(p + 2, __double2ll_rn((double)t.z * kVoxScale));
      }
}

// storeVolumePhoton, PMK:1147-1161
__device__ __forceinline__ void store_volume(const Sink &sk, v3 loc, v3 e) {
  if (!sk.acc) return;
  int vx = clampi(voxel_x(loc.x)), vy = clampi(voxel_x(loc.y)), vz = clampi(voxel_z(loc.z));
  unsigned long long *p = sk.acc + kAccHitEntries + ((vx * PM_GRID_N + vy) * PM_GRID_N + vz) * 3;
  acc_add(p + 0, __double2ll_rn((double)e.x * kVoxScale));
  acc_add(p + 1, __double2ll_rn((double)e.y * kVoxScale));
  acc_add(p + 2, __double2ll_rn((double)e.z * kVoxScale));
}

// getColor / filterColor, PMK:605-617
__device__ __forceinline__ v3 get_color(v3 in, int type, int idx) {
  v3 m = V(1.0f, 1.0f, 1.0f);
  if (type == 1 && idx == 0) m = V(0.0f, 1.0f, 0.0f);
  else if (type == 1 && idx == 2) m = V(1.0f, 0.0f, 0.0f);
  return V(fminf(m.x, in.x), fminf(m.y, in.y), fminf(m.z, in.z));
}

// ------------------------------------------------------------------------------------------------------
// emitPhotons, PMK:1215-1375 -- one thread per photon index in [first, last)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) trace_kernel(const __grid_constant__ DeviceScene sc, const float *__restrict__ table,
                                                    long long first, long long last, unsigned flags,
                                                    uint32_t w0, uint32_t z0, const MwcJump *__restrict__ J, Sink sk) {
  long long gi = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= last) return;
  const int index = (int)gi;
  const bool media = flags & PM_TRACE_MEDIA;
  const bool rec = (flags & PM_TRACE_RECORDS) != 0;
  int seq = 0;

  int bounces = 1;
  v3 rgb = V(10.0f, 10.0f, 10.0f);
  const v3 light = V(sc.light[0], sc.light[1], sc.light[2]);
  const v3 tdir = V(table[3 * gi], table[3 * gi + 1], table[3 * gi + 2]);
  v3 ray = normalize(tdir);
  const v3 original = ray;
  v3 prev = light, P = V(0.0f, 0.0f, 0.0f);
  Hit h; h.hit = 0; h.type = 0; h.idx = 0; h.dist = -1.0f;

  if (media) {   // PMK:1239-1272: 3 unit steps, deposit 5e-5*rgb, re-direct with 9 MWC draws (stream position 9*index)
    Mwc s;
    s.z = mwc_jump(J, 0, z0, 9u * (uint32_t)index);
    s.w = mwc_jump(J, 1, w0, 9u * (uint32_t)index);
#pragma unroll 1
    for (int i = 0; i < 3; i++) {
      rgb = subs(rgb, 1.0f);
      P = add(mul(ray, 1.0f), prev);
      v3 e = mul(rgb, 0.00005f);
      store_volume(sk, P, e);
      if (rec) append_record(sk, seq, 1, -1, -1, index, P, V(0.0f, 0.0f, 0.0f), e);
      seq++;
      v3 r;   // randomize(randomNumbers[i]), i = 0..2 (sic), PMK:1258
      r.x = rand_float(s, table[3 * i + 0]);
      r.y = rand_float(s, table[3 * i + 1]);
      r.z = rand_float(s, table[3 * i + 2]);
      ray = normalize(r);
      prev = P;
    }
    ray = original; prev = light;
  }
  if (index < 100) {   // CAUSTICS_PHOTONS, PMK:1274-1278: aimed at the glass sphere, jittered, not re-normalised
    ray = normalize(sub(V(sc.sph[0][0], sc.sph[0][1], sc.sph[0][2]), light));
    ray = add(ray, mul(normalize(tdir), 0.01f));
  }
  raytrace(sc, ray, prev, h);

  bool caustics = false, new_point = true;
#pragma unroll 1
  while (h.hit && bounces <= 5) {
    if (new_point) P = add(mul(ray, h.dist), prev);
    if (caustics) {
      rgb = mul(V(1.0f, 1.0f, 1.0f), 10.0f);
      store_photon(sk, h.type, h.idx, P, rgb);
      if (rec) append_record(sk, seq, 0, h.type, h.idx, index, P, ray, rgb);
      seq++;
    } else {
      rgb = mul(divs(mul(get_color(rgb, h.type, h.idx), 1.0f), __fsqrt_rn((float)bounces)), 5.0f);
      store_photon(sk, h.type, h.idx, P, rgb);
      if (rec) append_record(sk, seq, 0, h.type, h.idx, index, P, ray, rgb);
      seq++;
      {   // shadowPhoton, PMK:1185-1196: continue the same ray, deposit -0.25 at the next hit; dist/hit are not restored
        int t_type = h.type, t_idx = h.idx;
        v3 bumped = add(P, mul(ray, 0.00001f));
        raytrace(sc, ray, bumped, h);
        v3 sp = add(mul(ray, h.dist), bumped);
        v3 se = V(-0.25f, -0.25f, -0.25f);
        store_photon(sk, h.type, h.idx, sp, se);
        if (rec) append_record(sk, seq, 0, h.type, h.idx, index, sp, ray, se);
        seq++;
        h.type = t_type; h.idx = t_idx;
      }
    }
    prev = P;
    if (h.type == 0 && h.idx == 1) {          // mirror sphere
      follow_specular(sc, ray, prev, h, P, 1);
      caustics = false; new_point = false;
    } else if (h.type == 0 && h.idx == 0) {   // glass sphere
      follow_specular(sc, ray, prev, h, P, 0);
      caustics = true; new_point = false;
    } else {                                   // diffuse wall (prev == hit point: hazard H1)
      ray = reflect3(sc, ray, prev, h.type, h.idx, P);
      raytrace(sc, ray, P, h);
      caustics = false; new_point = true;
    }
    bounces++;
  }
}

// ------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------
cudaError_t launch_mwc_table(float *table, long long n, uint32_t w0, uint32_t z0, const MwcJump *J, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  unsigned blocks = (unsigned)((n + 255) / 256);
  mwc_table_kernel<<<blocks, 256, 0, st>>>(table, n, w0, z0, J);
  return cudaGetLastError();
}

cudaError_t launch_trace(const DeviceScene &sc, const float *table, long long first, long long last, unsigned flags,
                         uint32_t w0, uint32_t z0, const MwcJump *J, unsigned long long *acc, float4 *rec_pos,
                         float4 *rec_pow, float4 *rec_dir, unsigned long long *rec_count, long long rec_cap,
                         cudaStream_t st) {
  long long n = last - first;
  if (n <= 0) return cudaSuccess;
  Sink sk;
  sk.acc = (flags & PM_TRACE_NO_MAP) ? nullptr : acc;
  sk.rec_pos = rec_pos; sk.rec_pow = rec_pow; sk.rec_dir = rec_dir; sk.rec_count = rec_count; sk.rec_cap = rec_cap;
  unsigned blocks = (unsigned)((n + 255) / 256);
  trace_kernel<<<blocks, 256, 0, st>>>(sc, table, first, last, flags, w0, z0, J, sk);
  return cudaGetLastError();
}

}  // namespace pm
