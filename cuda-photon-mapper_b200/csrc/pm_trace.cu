// pm_trace.cu -- stage 1: random-direction table + photon emission/tracing (sm_100a).
//
// Replaces init_random_numbers_kernel (PMK:1483-1498), init_photons_kernel (PMK:1503-1521) and
// emit_photons_kernel/emitPhotons (PMK:1464-1480, :1215-1375).  Per photon the FP32 operation order is
// identical to the sequential oracle (pm_math.cuh), but the work is organised for the hardware:
//   * the MWC generator is index-addressed (jump-ahead), so the table fill and the medium-scatter draws are
//     fully parallel yet bit-identical to the reference's serial stream; a table row also carries 1/|row| (table_row);
//   * the medium walk (3 fixed steps, PMK:1239-1272) does not influence the surface walk.  trace_kernel is warp-specialised: warps
//     [0, vol_warps) of every 1024-thread CTA run the medium walk of the CTA's photon range, the other warps the surface walk
//     (PM_TRACE_SPLIT keeps the medium walk as its own launch, volume_kernel, for measurement).  When only deposit counts are wanted
//     (Mode A) the medium walk is a floating-point filter: approximate arithmetic, exact redo next to voxel boundaries
//     (volume_photon_fast; PM_TRACE_EXACT_MEDIUM switches it off);
//   * the surface walk is persistent, one CTA per SM.  Its general form is a state machine: every lane has ONE ray-scene intersection
//     site per iteration, and a lane that finishes its photon is refilled.  Mode A on a scene with the reference's object layout runs
//     in two phases instead (PM_TRACE_ONE_PHASE switches that off): 32 photons at the same point of their life go through the common
//     path in lock-step -- a ray that provably misses the spheres, wall hit, deposit, shadow ray, deposit, the bounce that dies on
//     normalize(0) --, survivors of the bounce are queued and come back as a block of their own, and only the photons whose ray may
//     hit a sphere are handed to the state machine (through a second queue);
//   * deposits go into exact int64 fixed-point accumulators keyed by wall voxel (pm_layout.h) instead of ~65 racy float RMWs per
//     photon; the wall accumulators are privatised per CTA in shared memory (184 320 B: 20 480 entries as 32-bit halves, because
//     shared memory has no native 64-bit add, plus 5 120 shadow-photon counters) and flushed once per CTA; medium deposits are
//     COUNTED in replicated 32-bit L2 counters and turned into energy by fold_volume_kernel;
//   * photon records (Mode B) are appended to SoA float4 buffers with warp-aggregated atomics.
#include "pm_kernels.cuh"

namespace pm {

// ------------------------------------------------------------------------------------------------------
// random table: thread i owns draws 3i..3i+2 of the stream that starts at (w0,z0); rand3 order x,y,z
// ------------------------------------------------------------------------------------------------------
// Rows [first, last) are generated, plus rows 0..2 (every photon's medium walk reads them, PMK:1258): a rank of a
// multi-GPU job only needs its own photon range.
// A table row carries w = 1 / sqrt(x*x + y*y + z*z), the factor of normalize(): emitPhotons normalises the same row in every frame
// (PMK:1229, :1242), so the two IEEE operations are done once when the row is written and the walks multiply (bit-identical).
__device__ __forceinline__ float4 table_row(float x, float y, float z) {
  return make_float4(x, y, z, rcp_rn(__fsqrt_rn(dot(V(x, y, z), V(x, y, z)))));
}
__device__ __forceinline__ v3 table_direction(float4 td) {
#ifdef PM_NO_TABLE_W
  return normalize(V(td.x, td.y, td.z));
#else
  return mul(V(td.x, td.y, td.z), td.w);
#endif
}

__global__ void __launch_bounds__(256) mwc_table_kernel(float4 *__restrict__ table, long long first, long long last, long long n, uint32_t w0,
                                                        uint32_t z0, const MwcJump *__restrict__ J) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long i = t < 3 ? t : first + (t - 3);
  if (t < 3 ? i >= n : (i >= last || i < 3)) return;   // rows 0..2 exist whatever the shard (a shard may end below row 3)
  Mwc s;
  uint32_t steps = (uint32_t)(3 * i);
  s.z = mwc_jump(J, 0, z0, steps);
  s.w = mwc_jump(J, 1, w0, steps);
  float x = rand_float(s, 1.0f), y = rand_float(s, 1.0f), z = rand_float(s, 1.0f);
  table[i] = table_row(x, y, z);
}

// Philox4x32-10 (Random123) table: row i = the reference's randFloat(1.0) mapping of the first three words of
// philox(counter = (i,0,0,0), key = seed) -- the counter-based alternative to the MWC stream for throughput runs
__global__ void __launch_bounds__(256) philox_table_kernel(float4 *__restrict__ table, long long n, uint32_t k0, uint32_t k1) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t c0 = (uint32_t)i, c1 = 0u, c2 = 0u, c3 = 0u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0, h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  auto unit = [](uint32_t u) { float rnd = div65535((float)((int)u)); rnd = rnd * 2.0f * 1.0f; return rnd - 1.0f; };
  table[i] = table_row(unit(c0), unit(c1), unit(c2));
}

// w of rows [0, n) from their (x, y, z): after the table was cleared (the reference's zero-initialised table: w = 1/0 = inf, so that
// the direction is 0 * inf = NaN like normalize(0)) or uploaded by the host
__global__ void __launch_bounds__(256) table_norm_kernel(float4 *__restrict__ table, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 t = table[i];
  table[i] = table_row(t.x, t.y, t.z);
}

// ------------------------------------------------------------------------------------------------------
// record sink
// ------------------------------------------------------------------------------------------------------
struct Sink {
  unsigned long long *acc;       // kAccEntries, or nullptr (PM_TRACE_NO_MAP)
  float4 *rec_pos, *rec_pow, *rec_dir;   // surface records: appended
  float4 *vrec_pos, *vrec_pow;           // volume records: slot = 3*(index-first)+step
  long long vrec_cap;
  unsigned long long *rec_count; // global append cursor (surface records)
  long long rec_cap;
  uint32_t *vol_cnt;             // kVolCntEntries deposit counters of the medium walk (volume_kernel only)
  uint32_t *vox_touched;         // set when anything lands in the acc_vox section (pm_layout.h ExchangeHeader), or nullptr
  unsigned long long *dbg;       // development aid (pm_trace_profile): kTraceDbgWords timestamps per CTA, or nullptr
  uint32_t *queue;               // per-warp queues of the two-phase surface walk (kQueueWords words per warp), or nullptr
};
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

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

// ------------------------------------------------------------------------------------------------------
// medium walk: PMK:1239-1272 + storeVolumePhoton PMK:1147-1161.  Grid-stride, no divergence, coalesced float4
// table rows.  Photon `index` owns draws [9*index, 9*index+9) of the MWC stream that starts at (w0,z0): a thread
// jumps to its first photon once (table look-ups) and then advances by the grid stride with one modular
// multiplication per lane (cz, cw = a^(9*stride) mod m, computed by the launcher).
// ------------------------------------------------------------------------------------------------------
// One photon of the medium walk: three unit steps from the light, a deposit at each, nine MWC draws (state s by value).
struct VolumeCtx {
  float4 t0, t1, t2;   // randomNumbers[i], i = 0..2 (sic, PMK:1258): the same three table rows scale every photon's draws
  uint32_t *cnt;       // this CTA's replica of the deposit counters
  v3 light;
  long long first;     // first photon of the launch (record slots are relative to it)
  bool rec;
  bool fast;           // volume_photon_fast may be used (see there)
  float g0, g1;        // its lower bounds on |r'|^2 for the draws scaled by t0 / t1: 4096 |t|^2
};
__device__ __forceinline__ VolumeCtx volume_ctx(const DeviceScene &sc, const float4 *__restrict__ table, long long first, unsigned flags,
                                                int replica, const Sink &sk) {
  VolumeCtx c;
  c.t0 = __ldg(table + 0); c.t1 = __ldg(table + 1); c.t2 = __ldg(table + 2);
  c.cnt = sk.vol_cnt + replica * 3 * PM_GRID_VOXELS;
  c.light = V(sc.light[0], sc.light[1], sc.light[2]);
  c.first = first;
  c.rec = (flags & PM_TRACE_RECORDS) != 0;
  // preconditions of volume_photon_fast's error bounds: finite scale factors that cannot overflow (table rows are 2 u / 65535 - 1 with a
  // 32-bit u, PMK:1039-1052: up to 65537 in magnitude), the light within 4.9 of the origin (so every deposit point has |coordinate| < 8),
  // and only counts are wanted
  bool ok = sk.acc != nullptr && !c.rec && !(flags & PM_TRACE_EXACT_MEDIUM);
  const float t[6] = {c.t0.x, c.t0.y, c.t0.z, c.t1.x, c.t1.y, c.t1.z};
#pragma unroll
  for (int i = 0; i < 6; i++) ok = ok && fabsf(t[i]) <= 1.0e9f;
  ok = ok && fabsf(c.light.x) < 4.9f && fabsf(c.light.y) < 4.9f && fabsf(c.light.z) < 4.9f;
  c.g0 = 4096.0f * (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  c.g1 = 4096.0f * (t[3] * t[3] + t[4] * t[4] + t[5] * t[5]);
  asm volatile("" : "+f"(c.g0), "+f"(c.g1));   // keep them in registers: the compiler would otherwise recompute both for every photon
  ok = ok && c.g0 > 0.0f && c.g1 > 0.0f;
#ifdef PM_NO_FAST_VOLUME
  ok = false;
#endif
  c.fast = ok;
  return c;
}
__device__ __forceinline__ void volume_photon(const VolumeCtx &c, const Sink &sk, float4 td, long long gi, Mwc s) {
  v3 rgb = V(10.0f, 10.0f, 10.0f);
  v3 ray = table_direction(td);
  v3 prev = c.light;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    rgb = subs(rgb, 1.0f);
    v3 P = add(mul(ray, 1.0f), prev);
    if (sk.acc) {   // e is a function of the step only: count the deposit, fold_volume_kernel turns counts into energy
      int vx = voxel_x_clamped(P.x), vy = voxel_x_clamped(P.y), vz = voxel_z_clamped(P.z);
      atomicAdd(c.cnt + i * PM_GRID_VOXELS + (vx * PM_GRID_N + vy) * PM_GRID_N + vz, 1u);
    }
    if (c.rec) {   // volume records have a fixed slot: deterministic order, coalesced, no atomics
      v3 e = mul(rgb, 0.00005f);
      long long slot = 3 * (gi - c.first) + i;
      if (slot < sk.vrec_cap) {
        sk.vrec_pos[slot] = make_float4(P.x, P.y, P.z, __uint_as_float(pack_meta(i, 1, -1, -1)));
        sk.vrec_pow[slot] = make_float4(e.x, e.y, e.z, __int_as_float((int)gi));
      }
    }
    const float4 tr = i == 0 ? c.t0 : (i == 1 ? c.t1 : c.t2);
    v3 r;
    r.x = rand_float(s, tr.x);
    r.y = rand_float(s, tr.y);
    r.z = rand_float(s, tr.z);
    ray = normalize(r);
    prev = P;
  }
}

// The same photon when only the deposit COUNTS are wanted (Mode A): a floating-point filter.  The voxel of a deposit point does not
// depend on the last bits of the point unless it lies next to a voxel boundary, so steps 1 and 2 are computed with approximate
// arithmetic (one multiply for x / 65535, fused multiply-adds, MUFU.RSQ for the normalisation: ~70 instead of ~100 instructions per
// step) and a photon whose approximate point comes within kVolEps of a boundary, in voxel units, in any coordinate is redone by the
// exact routine above.  Step 0 (light + table direction) is exact as it is.  Error budget, with |light| < 4.9 (VolumeCtx::fast) and
// |r'|^2 >= 4096 |m|^2 (checked per step; r = the random vector, r_j = 2 q_j m_j - m_j with |q_j| <= 32768.5 and m = the step's
// scale row, so |r| ~ 1e4 |m| and the check fails about once in 1e11 draws; ' = approximate):
//   * r'_j vs r_j: x / 65535 by one multiply is 3 * 2^-24 relative off the rounded quotient, the fused 2 q m - m saves roundings:
//     |r'_j - r_j| <= 2^-21 (|r_j| + |m_j|), so |r' - r| <= 2^-21 (1 + 1/64) |r|;
//   * direction: the exact chain (dot, sqrt, 1 / x, multiply) is within 4.5 * 2^-24 of r / |r|, the approximate one (rsqrt.approx:
//     2 ulp) within 6.5 * 2^-24 of r' / |r'|, and |r' / |r'| - r / |r|| <= 2 |r' - r| / |r|: per component <= 27.3 * 2^-24 = 1.63e-6;
//   * point: each step adds that plus two roundings of a coordinate below 8 (2 * 2^-24 * 8): 2.6e-6 after step 1, 5.2e-6 after step 2;
//   * voxel coordinate m = (32 p + 48) / 3 (or 16 p / 3): 10.67 * 5.2e-6 + the evaluation of m' itself (6e-6) = 6.2e-5 < kVolEps = 2^-13.
// If m' is farther than that from every integer, floor(m') is the exact voxel (pm_math.cuh: voxel = clamp(floor(m))); the one input
// range where the reference's double arithmetic deviates from that formula, p in [-2^-53, 0), sits 1e-15 from a boundary and falls
// back like the rest.  ~1.5e-3 of the photons take the fallback.  Checked end to end: the accumulators of a Mode A trace are compared
// bit for bit with those of the exact (records) trace at 16M photons (tests/test_gpu_fullsize.py).
#ifndef PM_VOL_EPS
#define PM_VOL_EPS 0x1p-13f
#endif
constexpr float kVolEps = PM_VOL_EPS;
__device__ __forceinline__ int fast_voxel(float p, float scale, float bias, bool &amb) {
  const float m = __fmaf_rn(p, scale, bias);
  const float f = floorf(m);
  const float d = m - f;
  amb = amb | (d < kVolEps) | (d > 1.0f - kVolEps);
  const int k = (int)f;
  return k < 0 ? 0 : (k < PM_GRID_N ? k : PM_GRID_N - 1);
}
// returns false when the photon has to be redone by volume_photon (nothing was deposited)
__device__ __forceinline__ bool volume_photon_fast(const VolumeCtx &c, float4 td, Mwc s) {
  bool amb = !(td.w > 0.0f && td.w < 3.0e38f);   // a zero / non-finite row: NaN directions, the exact routine knows what the reference does
  const v3 P0 = add(table_direction(td), c.light);
  int idx[3];
  idx[0] = (voxel_x_clamped(P0.x) * PM_GRID_N + voxel_x_clamped(P0.y)) * PM_GRID_N + voxel_z_clamped(P0.z);
  v3 P = P0;
#pragma unroll
  for (int i = 1; i < 3; i++) {
    const float4 tr = i == 1 ? c.t0 : c.t1;   // the draws after deposit i - 1 are scaled by row i - 1
    const float c65535 = 1.0f / 65535.0f;
    const float qx = (float)((int)mwc_next(s)) * c65535, qy = (float)((int)mwc_next(s)) * c65535, qz = (float)((int)mwc_next(s)) * c65535;
    const float rx = __fmaf_rn(qx, 2.0f * tr.x, -tr.x), ry = __fmaf_rn(qy, 2.0f * tr.y, -tr.y), rz = __fmaf_rn(qz, 2.0f * tr.z, -tr.z);
    const float dd = __fmaf_rn(rz, rz, __fmaf_rn(ry, ry, rx * rx));
    amb = amb | !(dd >= (i == 1 ? c.g0 : c.g1));
    float inv;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(dd));
    P = V(__fmaf_rn(rx, inv, P.x), __fmaf_rn(ry, inv, P.y), __fmaf_rn(rz, inv, P.z));
    const int vx = fast_voxel(P.x, 32.0f / 3.0f, 16.0f, amb), vy = fast_voxel(P.y, 32.0f / 3.0f, 16.0f, amb);
    const int vz = fast_voxel(P.z, 16.0f / 3.0f, 0.0f, amb);
    idx[i] = (vx * PM_GRID_N + vy) * PM_GRID_N + vz;
  }
  if (amb) return false;
#pragma unroll
  for (int i = 0; i < 3; i++) atomicAdd(c.cnt + i * PM_GRID_VOXELS + idx[i], 1u);
  return true;
}

// One thread's share of the medium walk, strided: photons gi, gi + stride, ... < last.  cw, cz = a^(9*stride) mod m.
__device__ __forceinline__ void volume_walk(const DeviceScene &sc, const float4 *__restrict__ table, long long first, long long gi,
                                            long long last, long long stride, unsigned flags, uint32_t w0, uint32_t z0,
                                            const MwcJump *__restrict__ J, uint32_t cw, uint32_t cz, int replica, const Sink &sk) {
  if (gi >= last) return;
  const VolumeCtx c = volume_ctx(sc, table, first, flags, replica, sk);
  Mwc base;
  base.z = mwc_jump(J, 0, z0, 9u * (uint32_t)gi);
  base.w = mwc_jump(J, 1, w0, 9u * (uint32_t)gi);
  float4 td_next = __ldg(table + gi);
  for (; gi < last; gi += stride) {
    const float4 td = td_next;
    if (gi + stride < last) td_next = __ldg(table + gi + stride);   // next row in flight under this photon's walk
    if (!(c.fast && volume_photon_fast(c, td, base))) volume_photon(c, sk, td, gi, base);
    base.z = mulmod(base.z, cz, mwc_modulus(0));
    base.w = mulmod(base.w, cw, mwc_modulus(1));
  }
}

// The medium walk inside the fused kernel.  The CTA's range is cut into S slices of `per` photons, one per surface warp; the V
// medium-walk warps visit 32-photon blocks in the order (block 0 of slices 0..S-1, block 1 of slices 0..S-1, ...), i.e. they
// sweep every slice at the same rate as its surface warp does, so the table rows are fetched from DRAM once and found in L2 by
// the other walk (a plain strided sweep of the CTA range read the whole table twice).  A lane's consecutive photons lie
// V slices apart, or V - S slices + one block when the slice index wraps: two jump multipliers per MWC lane, from the host.
struct SliceJump { uint32_t fz, fw, wz, ww; };   // forward (V*per photons) and wrap ((V-S)*per + 32 photons), z and w lanes
__device__ __forceinline__ void volume_walk_sliced(const DeviceScene &sc, const float4 *__restrict__ table, long long first,
                                                   long long cta_first, long long cta_last, long long per, int S, int V, unsigned flags,
                                                   uint32_t w0, uint32_t z0, const MwcJump *__restrict__ J, SliceJump jump, int replica,
                                                   const Sink &sk) {
  const int lane = threadIdx.x & 31;
  if (cta_first >= cta_last || per <= 0) return;
  // 32-bit offsets inside the CTA's range (it holds < 2^31 photons): the loop's bookkeeping is a handful of integer instructions
  const int per32 = (int)per, n_cta = (int)(cta_last - cta_first);
  const float4 *__restrict__ tab = table + cta_first;
  int s = threadIdx.x >> 5;            // V <= S: the warp's first block is block 0 of slice `warp`
  int b32 = 0;                         // first photon of the lane's block inside the slice
  int off = s * per32 + lane;          // the lane's photon, relative to cta_first
  const int step_f = V * per32, step_w = (V - S) * per32 + 32;
  const VolumeCtx c = volume_ctx(sc, table, first, flags, replica, sk);
  Mwc base;
  base.z = mwc_jump(J, 0, z0, 9u * (uint32_t)(cta_first + off));
  base.w = mwc_jump(J, 1, w0, 9u * (uint32_t)(cta_first + off));
  bool valid = lane < per32 && off < n_cta;
  float4 td_next = valid ? __ldg(tab + off) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  while (b32 < per32) {
    const float4 td = td_next;
    const int off_now = off;
    const bool valid_now = valid;
    // the lane's next block
    s += V;
    const bool wrap = s >= S;
    if (wrap) { s -= S; b32 += 32; }
    off += wrap ? step_w : step_f;
    valid = b32 + lane < per32 && off < n_cta;
    if (valid) td_next = __ldg(tab + off);   // in flight under this photon's walk
    if (valid_now && !(c.fast && volume_photon_fast(c, td, base))) volume_photon(c, sk, td, cta_first + off_now, base);
    base.z = mulmod(base.z, wrap ? jump.wz : jump.fz, mwc_modulus(0));
    base.w = mulmod(base.w, wrap ? jump.ww : jump.fw, mwc_modulus(1));
  }
}

// stand-alone medium walk (PM_TRACE_SPLIT): grid-stride over the whole range
__global__ void __launch_bounds__(256) volume_kernel(const __grid_constant__ DeviceScene sc, const float4 *__restrict__ table,
                                                     long long first, long long last, unsigned flags, uint32_t w0, uint32_t z0,
                                                     const MwcJump *__restrict__ J, uint32_t cw, uint32_t cz, Sink sk) {
  volume_walk(sc, table, first, first + (long long)blockIdx.x * blockDim.x + threadIdx.x, last, (long long)gridDim.x * blockDim.x, flags,
              w0, z0, J, cw, cz, blockIdx.x % kVolCntReplicas, sk);
}

// counts -> energy: acc_grey[0][v] += sum over replicas and steps of count * quantum(step); the counts are cleared.
// The quanta are computed with the same FP32 operations as the walk (rgb = 10 - 1 - ..., e = rgb * 0.00005f).
__global__ void __launch_bounds__(256) fold_volume_kernel(uint32_t *__restrict__ cnt, unsigned long long *__restrict__ acc) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= PM_GRID_VOXELS) return;
  long long sum = 0;
  float rgb = 10.0f;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    rgb = rgb - 1.0f;
    const long long q = __double2ll_rn((double)(rgb * 0.00005f) * kVoxScale);
    unsigned long long n = 0;
#pragma unroll
    for (int r = 0; r < kVolCntReplicas; r++) {
      uint32_t *p = cnt + (r * 3 + i) * PM_GRID_VOXELS + v;
      n += *p;
      *p = 0u;
    }
    sum += (long long)n * q;
  }
  if (sum) acc[kAccHitEntries + kAccVoxEntries + v] += (unsigned long long)sum;   // one thread per voxel, stream-ordered
}

// ------------------------------------------------------------------------------------------------------
// surface walk
// ------------------------------------------------------------------------------------------------------
// CTA-private accumulator for on-slab wall hits: kAccHitEntries x (lo, hi) 32-bit halves in shared memory.
struct SmemAcc {
  uint32_t *lo, *hi;
  uint32_t *shadow;   // per wall texel: number of shadow photons (each deposits exactly -0.25 grey), folded in at the flush
  // A surface deposit is one of 5 * {1, 0} / sqrt(bounces) or 10 (PMK:1296-1298; the shadow photons' -0.25 is counted): 0 <= e <= 10,
  // so e * 2^24 fits the low half and only the carry reaches the high half.
  __device__ __forceinline__ void add(int e, float energy) {
    const uint32_t v = __float2uint_rn(energy * (float)kHitScale);
    if (v == 0u) return;
    const uint32_t old = atomicAdd(lo + e, v);
    if ((uint32_t)(old + v) < old) atomicAdd(hi + e, 1u);   // carry out of the low half, counted exactly once
  }
};

// The clamped voxel is off the wall's slab (a wall that is not on the map boundary): expand the splat per photon.  Rare.
// (Measured: making this and the mirror/glass direction updates real calls costs 17% -- the call ABI spills the walk's state.)
static __device__ __forceinline__ void splat_offslab(const Sink &sk, int id, int on_slab, int vx, int vy, int vz, v3 e) {
  unsigned long long *vox = sk.acc + kAccHitEntries;
  if (sk.vox_touched) *sk.vox_touched = 1u;
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

// storePhoton + splatEnergy + storeNeighborPhoton, PMK:1059-1144, :1164-1183 (type 0 = sphere: nothing is stored)
__device__ __forceinline__ void store_photon(const Sink &sk, SmemAcc &sa, int type, int id, v3 loc, v3 e, bool shadow) {
  if (!sk.acc || type == 0) return;
  int vx = voxel_x_clamped(loc.x), vy = voxel_x_clamped(loc.y), vz = voxel_z_clamped(loc.z);
  // wall id -> slab axis / slab index / in-plane coordinates, as splatEnergy hard-codes them (selects, no branches)
  const int ax = (id == 0 || id == 2) ? 0 : ((id == 1 || id == 3) ? 1 : 2);
  const int slab = (id == 0 || id == 3 || id == 4) ? PM_GRID_N - 1 : 0;
  const int vfix = ax == 0 ? vx : (ax == 1 ? vy : vz);
  const int a = ax == 0 ? vy : vx, b = ax == 2 ? vy : vz;
  const int on_slab = (unsigned)id > 4u ? -1 : (vfix == slab ? 1 : 0);
  if (on_slab == 1) {   // the common case: keyed energy sum, the 6x6 stencil is applied once per voxel in pm_map.cu
    int en = ((id * PM_GRID_N + a) * PM_GRID_N + b) * 4;
    if (shadow) atomicAdd(sa.shadow + (en >> 2), 1u);   // half of all deposits: one 32-bit count instead of a 64-bit add with carry
    else if (e.x == e.y && e.y == e.z) sa.add(en + 3, e.x);
    else {
      sa.add(en + 0, e.x);
      sa.add(en + 1, e.y);
      sa.add(en + 2, e.z);
    }
    return;
  }
  splat_offslab(sk, id, on_slab, vx, vy, vz, e);   // rare: the clamped voxel is off the wall's slab
}

// store_photon for phase F of the two-phase walk: the hit is on wall `id` of the reference's layout, and the deposit is either a shadow
// photon or a bounce's energy, which is one value `v` (fixed point) in one channel `ch` of the wall texel (after a green wall: g, after
// a red one: r, white walls only: the grey plane; nothing, v = 0, after both).  Same voxel, same slab test, same accumulator entries as
// store_photon; e = the same energy as a vector, for the off-slab expansion.
__device__ __forceinline__ void store_photon_wall_std(const Sink &sk, SmemAcc &sa, int id, v3 loc, bool shadow, int ch, uint32_t v, v3 e) {
  if (!sk.acc) return;
  const int vx = voxel_x_clamped(loc.x), vy = voxel_x_clamped(loc.y), vz = voxel_z_clamped(loc.z);
  const int ax = std_axis(id);
  const int slab = ((0x19 >> id) & 1) ? PM_GRID_N - 1 : 0;   // ids 0, 3, 4 lie on the upper boundary of the map
  const int vfix = ax == 0 ? vx : (ax == 1 ? vy : vz);
  const int a = ax == 0 ? vy : vx, b = ax == 2 ? vy : vz;
  if (vfix == slab) {
    const int tex = (id * PM_GRID_N + a) * PM_GRID_N + b;
    if (shadow) atomicAdd(sa.shadow + tex, 1u);
    else if (v) {   // a photon that has lost all three channels still casts its shadow photons
      const uint32_t old = atomicAdd(sa.lo + tex * 4 + ch, v);
      if ((uint32_t)(old + v) < old) atomicAdd(sa.hi + tex * 4 + ch, 1u);
    }
    return;
  }
  splat_offslab(sk, id, 0, vx, vy, vz, e);
}

// getColor / filterColor, PMK:605-617
__device__ __forceinline__ v3 get_color(v3 in, int type, int idx) {
  v3 m = V(1.0f, 1.0f, 1.0f);
  if (type == 1 && idx == 0) m = V(0.0f, 1.0f, 0.0f);
  else if (type == 1 && idx == 2) m = V(1.0f, 0.0f, 0.0f);
  return V(fminf(m.x, in.x), fminf(m.y, in.y), fminf(m.z, in.z));
}

// planeNormal + reflect3 (PMK:196-209, :664-668) for an axis-aligned wall of the reference's layout, scalar: the normal is
// (0, .., wd / |wd|, .., 0) with wd = P[axis] - offset -- sqrt(RN(wd * wd)) == |wd| in binary floating point when the square neither
// under- nor overflows (2^-34 <= |wd| <= ~1e6 under std_walls_ok; checked by pm_selftest_fdiv) --, its two zero components contribute
// +-0 to the dot product (ray[axis] != 0: the ray has just hit this wall) and (+0) * k to the subtraction, which is kept because it
// turns a -0 ray component into +0 when k < 0.  Callers have ruled out wd * wd == 0 / NaN (hazard H1).
__device__ __forceinline__ v3 reflect_wall_std(const DeviceScene &sc, v3 ray, v3 P, int id, int wax) {
  const float wd = comp(P, wax) - sc.pl_off[id];
  const float na = wd * rcp_rn(fabsf(wd));
  const float ra = comp(ray, wax);
  const float k = 2.0f * (ra * na);
  const float z = 0.0f * k;
  v3 rr = V(ray.x - z, ray.y - z, ray.z - z);
  set_comp(rr, wax, ra - na * k);
  return normalize(rr);
}

// per-warp queues QL, QM of the two-phase walk: structure of arrays, word k of entry j at [k * kQueueCap + j]; word 0 = photon index
// (27 bits) | id of the wall last hit << 27 | fresh << 31, words 1-3 a ray, 4-6 a point, word 7 = colour mask | bounce number << 3
constexpr int kQueueCap = kTraceQueueCap, kQueueWords = kTraceQueueWords;

// lane states: what the NEXT intersection result means for this lane
enum : int { ST_IDLE = 0, ST_PRIMARY, ST_SHADOW, ST_CHAIN_R, ST_CHAIN_F1, ST_CHAIN_F2 };

constexpr int kSurfaceThreads = 1024;
constexpr int kRefillLanes = 8;   // measured: 4 is 13% slower, 12 and 16 are the same as 8; taking chunks off a CTA-wide cursor
                                  // instead of static per-warp slices is 20% slower; prefetching the table rows changes nothing.
                                  // Round 2 (the slowest warp of a CTA ends 5% after the mean at 16M photons, 11% at 2M): pooling only the
                                  // last 1/8..1/2 of every slice in 32..256-photon chunks, 32-bit cursors, no spills: still 4-27% slower
                                  // (0.985 -> 1.03..1.25 ms) -- every refill at a chunk boundary leaves lanes idle for an iteration
#ifndef PM_QREFILL
#define PM_QREFILL 8
#endif
constexpr int kQueueRefillLanes = PM_QREFILL;   // the same threshold when the lanes are refilled from the two-phase walk's queue
constexpr int kShadowEntries = kAccHitEntries / 4;                                      // one per wall texel
constexpr size_t kSurfaceSmem = sizeof(uint32_t) * (2 * kAccHitEntries + kShadowEntries);   // 184 320 B

// Warp-specialised: warps [0, vol_warps) of every CTA run the medium walk of the CTA's photon range (L2-atomic bound,
// nearly no issue slots), the other warps run the surface walk (issue bound, no L2 traffic) -- the two halves of
// emitPhotons overlap on the same SM instead of running back to back.  vol_warps = 0: surface walk only.
// kRec: the launch appends photon records (PM_TRACE_RECORDS); Mode A compiles that path out.  kStd: the scene has the reference's
// object layout (pm_math.cuh raytrace<kStd>)
template <bool kRec, bool kStd>
__global__ void __launch_bounds__(kSurfaceThreads, 1) trace_kernel(const __grid_constant__ DeviceScene sc,
                                                                   const float4 *__restrict__ table, long long first, long long last,
                                                                   unsigned flags, int vol_warps, uint32_t w0, uint32_t z0,
                                                                   const MwcJump *__restrict__ J, SliceJump jump, Sink sk) {
  extern __shared__ uint32_t smem_u32[];
  unsigned long long *dbg = sk.dbg ? sk.dbg + (size_t)blockIdx.x * kTraceDbgWords : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = global_ns();
  SmemAcc sa;
  sa.lo = smem_u32; sa.hi = smem_u32 + kAccHitEntries; sa.shadow = smem_u32 + 2 * kAccHitEntries;
  for (int i = threadIdx.x; i < 2 * kAccHitEntries + kShadowEntries; i += blockDim.x) smem_u32[i] = 0u;
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[1] = global_ns();

  const bool media = flags & PM_TRACE_MEDIA;
  const bool rec = kRec;
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  // the CTA owns a contiguous part of the photon range
  const long long per_cta = (last - first + gridDim.x - 1) / gridDim.x;
  const long long cta_first = first + (long long)blockIdx.x * per_cta < last ? first + (long long)blockIdx.x * per_cta : last;
  const long long cta_last = cta_first + per_cta < last ? cta_first + per_cta : last;
  const int warp = threadIdx.x >> 5;
  // each surface warp owns a contiguous slice of `per` photons of the CTA's range (the same `per` in every CTA, so that
  // the host can precompute the medium walk's jump multipliers); lanes are refilled from it
  const int surf_warps = (blockDim.x >> 5) - vol_warps;
  const long long per = (per_cta + surf_warps - 1) / surf_warps;
  if (warp < vol_warps)
    volume_walk_sliced(sc, table, first, cta_first, cta_last, per, surf_warps, vol_warps, kRec ? flags : (flags & ~PM_TRACE_RECORDS), w0,
                       z0, J, jump, blockIdx.x % kVolCntReplicas, sk);
  long long cur = cta_first + (long long)(warp - vol_warps) * per;
  long long end = cur + per < cta_last ? cur + per : cta_last;
  if (warp < vol_warps) { cur = 0; end = 0; }
  if (cur > end) cur = end;

  const v3 light = V(sc.light[0], sc.light[1], sc.light[2]);
  const v3 rgb0 = media ? V(7.0f, 7.0f, 7.0f) : V(10.0f, 10.0f, 10.0f);   // the medium walk leaves rgb = 10-1-1-1
  // ---- two-phase walk (Mode A, reference layout, scene conditions checked by the host: DeviceScene::fast_ok) ----
  // Phase F, lock-step: 32 photons at the same point of their life -- fresh from the light, or just past a wall bounce they survived --
  // go through the common path together: [reflection off the wall,] a ray that provably misses both spheres, wall hit, deposit, shadow
  // ray, deposit, and the bounce that dies on normalize(0) (hazard H1, ~88% of them).  Survivors of the bounce go to the warp's queue
  // QL and come back as a later block; a photon whose ray cannot be shown to miss the spheres, or that belongs to the caustic emitter,
  // goes to the queue QM.  Phase G, when QM is nearly full or nothing else is left: the general state machine below, its lanes
  // refilled from QM instead of the slice.  Per photon the operations are those of the machine, in the same order.
  constexpr bool kTwoPhase = kStd && !kRec;
  const bool two_phase = kTwoPhase && sc.fast_ok && sk.queue != nullptr && warp >= vol_warps;
  uint32_t *const ql_base = sk.queue + ((size_t)blockIdx.x * (kSurfaceThreads / 32) + warp) * kQueueWords;
  uint32_t *const qm_base = ql_base + 8 * kQueueCap;
  int ql = 0, qm = 0;   // entries in the queues (warp-uniform)
  // first-bounce energy of a channel the wall's colour lets through: min(1, rgb0) * 1 / sqrt(1) * 5 with the machine's operations
  const float e1 = fminf(1.0f, rgb0.x) * 1.0f * sc.inv_sqrt_bounce[1] * 5.0f;
  const uint32_t v1 = __float2uint_rn(e1 * (float)kHitScale);
  for (;;) {
    if (kTwoPhase && two_phase) {
      bool have_next = false;
      float4 td_next = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      // QL entry: the incoming ray and the hit point of the bounce just survived, the wall, the colour mask and bounce number that come
      // next; QM entry: the ray about to be traced and its origin with the mask / bounce number it has now, or a bare index (bit 31)
      auto push = [&](bool to_l, bool to_m, int index, int wall, bool fresh, v3 r, v3 pt, int mk, int bn) {
        const unsigned ml = __ballot_sync(0xffffffffu, to_l), mm = __ballot_sync(0xffffffffu, to_m);
        if (to_l | to_m) {
          uint32_t *const qb = to_l ? ql_base : qm_base;
          const int j = to_l ? ql + __popc(ml & lt_mask) : qm + __popc(mm & lt_mask);
          qb[j] = (uint32_t)index | ((uint32_t)wall << 27) | (fresh ? 0x80000000u : 0u);
          if (!fresh) {
            qb[1 * kQueueCap + j] = __float_as_uint(r.x); qb[2 * kQueueCap + j] = __float_as_uint(r.y); qb[3 * kQueueCap + j] = __float_as_uint(r.z);
            qb[4 * kQueueCap + j] = __float_as_uint(pt.x); qb[5 * kQueueCap + j] = __float_as_uint(pt.y); qb[6 * kQueueCap + j] = __float_as_uint(pt.z);
            qb[7 * kQueueCap + j] = (uint32_t)(mk | (bn << 3));
          }
        }
        ql += __popc(ml); qm += __popc(mm);
        __syncwarp();
      };
      for (;;) {
        if (qm > kQueueCap - 32) break;   // a block may add 32 entries to QM: the machine drains it first
        if (ql >= 32 || (cur >= end && ql > 0)) {
          // ---- a block of survivors: the rest of their wall bounce (PMK:1365-1369), then the next one ----
          const int n = ql < 32 ? ql : 32;
          const bool act = lane < n;
          int index = 0, b = 2, mask = 7, id_prev = 0;
          v3 ray = V(0.0f, 0.0f, 0.0f), org = ray;
          if (act) {
            const int j = ql - 1 - lane;
            const uint32_t w0q = ql_base[j], w7q = ql_base[7 * kQueueCap + j];
            index = (int)(w0q & 0x07ffffffu); id_prev = (int)(w0q >> 27) & 7;
            mask = (int)(w7q & 7u); b = (int)(w7q >> 3);
            const v3 rin = V(__uint_as_float(ql_base[1 * kQueueCap + j]), __uint_as_float(ql_base[2 * kQueueCap + j]), __uint_as_float(ql_base[3 * kQueueCap + j]));
            org = V(__uint_as_float(ql_base[4 * kQueueCap + j]), __uint_as_float(ql_base[5 * kQueueCap + j]), __uint_as_float(ql_base[6 * kQueueCap + j]));
            ray = reflect_wall_std(sc, rin, org, id_prev, std_axis(id_prev));
          }
          ql -= n;
          __syncwarp();   // the entries are read before this block's survivors overwrite them
          bool push_l = false, push_m = false;
          v3 P2 = org;
          int id = id_prev, mask2 = mask;
          if (act) {
            // A sphere is out when raySphere finds D <= 0, or B >= 0 with the origin outside it (sign = -1: the root -B - sqrt(D) is <= 0 and
            // checkDistance rejects it) -- FP32 logic on raySphere's own terms, no geometry; anything else is the machine's business
            const float A = dot(ray, ray);
            bool aimed = false;
#pragma unroll
            for (int i = 0; i < 2; i++) {
              const v3 sv = sub(V(sc.sph[i][0], sc.sph[i][1], sc.sph[i][2]), org);
              const float B = -2.0f * dot(sv, ray);
              const float C = dot(sv, sv) - sc.sph_r2[i];
              const float D = B * B - 4.0f * A * C;
              aimed = aimed | ((D > 0.0f) & !((B >= 0.0f) & !(C < -0.00001f)));   // (double)C < -0.00001 <=> C < (float)-0.00001 for a float C
            }
            if (aimed) push_m = true;
            else {
              float dist = 999999.9f;
              int best = -1;
              ray_walls_std(sc, ray, org, dist, best);
              if (best >= 0) {   // b <= 5 by construction
                id = best & 7;
                P2 = add(mul(ray, dist), org);
                // getColor + PMK:1298: a channel keeps min(1, previous >= 1) = 1, or is 0 from the first coloured wall on; times 1/sqrt(b), times 5
                mask2 = mask & (id == 0 ? 2 : (id == 2 ? 1 : 7));
                const float eb = 1.0f * 1.0f * sc.inv_sqrt_bounce[b] * 5.0f;
                {
                  const v3 e = V((mask2 & 1) ? eb : 0.0f, (mask2 & 2) ? eb : 0.0f, (mask2 & 4) ? eb : 0.0f);
                  store_photon_wall_std(sk, sa, id, P2, false, mask2 == 7 ? 3 : (mask2 == 2 ? 1 : 0), mask2 ? __float2uint_rn(eb * (float)kHitScale) : 0u, e);
                }
                const v3 o2 = add(P2, mul(ray, 0.00001f));
                float d2 = 999999.9f;
                int b2 = -1;
                const bool skippable = (comp(ray, std_axis(id)) * sc.wall_side[id] < 0.0f) & (fmaxf(fmaxf(fabsf(o2.x), fabsf(o2.y)), fabsf(o2.z)) <= 16.0f);
                const unsigned need = skippable ? sc.shadow_need[id] : 3u;
                if (need & 1u) ray_sphere(sc, 0, ray, o2, A, d2, b2);
                if (need & 2u) ray_sphere(sc, 1, ray, o2, A, d2, b2);
                ray_walls_std(sc, ray, o2, d2, b2);
                if (b2 < 0 || b2 >= 8)   // a miss keeps the primary hit's ids (stale, as in the reference); a sphere stores nothing
                  store_photon_wall_std(sk, sa, b2 >= 0 ? (b2 & 7) : id, add(mul(ray, d2), o2), true, 3, 0u, V(-0.25f, -0.25f, -0.25f));
                const float wd = comp(P2, std_axis(id)) - sc.pl_off[id];
                const float wdd = wd * wd;
                push_l = !(wdd == 0.0f || wdd != wdd) && b < 5;   // after the fifth bounce the loop condition ends the photon (PMK:1289)
              }
            }
          }
          push(push_l, push_m, index, push_l ? id : id_prev, false, ray, push_l ? P2 : org, push_l ? mask2 : mask, push_l ? b + 1 : b);
        } else if (cur < end) {
          // ---- a block of fresh photons: the same path from the light, with what is known about it ----
          if (!have_next) { if (cur + lane < end) td_next = __ldg(table + cur + lane); have_next = true; }
          const long long gi = cur + lane;
          const bool act = gi < end;
          cur = cur + 32 < end ? cur + 32 : end;
          const float4 td = td_next;
          if (cur + lane < end) td_next = __ldg(table + cur + lane);   // the next block's rows are in flight under this block
          bool push_l = false, push_m = false;
          v3 fr = V(0.0f, 0.0f, 0.0f), fP = fr;
          int fid = 0;
          if (act) {
            fr = table_direction(td);
            // raySphere's ray-independent terms come from the host (same operations); the light is outside both spheres (light_C >= 0,
            // part of fast_ok), so the rejection test is D <= 0 or B >= 0
            const float A = dot(fr, fr);
            bool aimed = (int)gi < 100;   // CAUSTICS_PHOTONS
#pragma unroll
            for (int i = 0; i < 2; i++) {
              const float B = -2.0f * dot(V(sc.light_s[i][0], sc.light_s[i][1], sc.light_s[i][2]), fr);
              const float D = B * B - 4.0f * A * sc.light_C[i];
              aimed = aimed | ((D > 0.0f) & !(B >= 0.0f));
            }
            if (aimed) push_m = true;
            else {
              float dist = 999999.9f;
              int best = -1;
              ray_walls_std(sc, fr, light, dist, best);
              if (best >= 0) {   // bounces == 1
                const int id = best & 7;
                fP = add(mul(fr, dist), light);
                {   // the first bounce's energy: e1 in the wall's colour channels (getColor: green wall 0, red wall 2), PMK:1296-1298
                  const v3 e = V(id == 0 ? 0.0f : e1, id == 2 ? 0.0f : e1, (id == 0 || id == 2) ? 0.0f : e1);
                  store_photon_wall_std(sk, sa, id, fP, false, id == 0 ? 1 : (id == 2 ? 0 : 3), v1, e);
                }
                // shadowPhoton: a ray that has just crossed the wall's plane from the light's side cannot come back to a sphere that lies
                // wholly on that side (DeviceScene::shadow_need, argued in pm_api.cu); the other spheres are tested as usual
                const v3 o2 = add(fP, mul(fr, 0.00001f));
                float d2 = 999999.9f;
                int b2 = -1;
                const unsigned need = fmaxf(fmaxf(fabsf(o2.x), fabsf(o2.y)), fabsf(o2.z)) <= 16.0f ? sc.shadow_need[id] : 3u;
                if (need & 1u) ray_sphere(sc, 0, fr, o2, A, d2, b2);
                if (need & 2u) ray_sphere(sc, 1, fr, o2, A, d2, b2);
                ray_walls_std(sc, fr, o2, d2, b2);
                if (b2 < 0 || b2 >= 8)
                  store_photon_wall_std(sk, sa, b2 >= 0 ? (b2 & 7) : id, add(mul(fr, d2), o2), true, 3, 0u, V(-0.25f, -0.25f, -0.25f));
                const float wd = comp(fP, std_axis(id)) - sc.pl_off[id];
                const float wdd = wd * wd;
                if (!(wdd == 0.0f || wdd != wdd)) { push_l = true; fid = id; }
              }
            }
          }
          push(push_l, push_m, (int)gi, fid, push_m, fr, fP, fid == 0 ? 2 : (fid == 2 ? 1 : 7), 2);
        } else break;   // no fresh photons, QL empty
      }
      if (qm == 0) break;   // nothing fresh, QL and QM empty
    }

    // ---- phase G / the only phase otherwise: the state machine ----
    int state = ST_IDLE, index = 0, seq = 0, bounces = 1, level = 1, t_type = 0, t_idx = 0;
    bool caustics = false, new_point = true, chain_glass = false;
    v3 rgb = V(0.0f, 0.0f, 0.0f), ray = rgb, prev = rgb, P = rgb, org = rgb;
    Hit h; h.hit = 0; h.type = 0; h.idx = 0; h.dist = -1.0f;
    for (;;) {
      // ---- refill idle lanes from the warp's slice (emitPhotons prologue, PMK:1229-1237, :1274-1280) or from its queue.  The
      //      prologue runs with only the idle lanes active, so it is deferred until a quarter of the warp is idle ----
      const unsigned idle = __ballot_sync(0xffffffffu, state == ST_IDLE);
      const long long avail = (kTwoPhase && two_phase) ? (long long)qm : end - cur;
      if (avail > 0 && (__popc(idle) >= ((kTwoPhase && two_phase) ? kQueueRefillLanes : kRefillLanes) || idle == 0xffffffffu)) {
        const int rank = __popc(idle & lt_mask);
        if (state == ST_IDLE && rank < avail) {
          long long cand = cur + rank;
          bool resumed = false;
          if (kTwoPhase && two_phase) {
            const int j = qm - 1 - rank;
            const uint32_t w0q = qm_base[j];
            cand = (long long)(w0q & 0x07ffffffu);
            if (!(w0q & 0x80000000u)) {
              // past a wall bounce, about to trace the reflected ray (top of the bounce loop, PMK:1289): the state phase F left it in
              resumed = true;
              const uint32_t w7q = qm_base[7 * kQueueCap + j];
              ray = V(__uint_as_float(qm_base[1 * kQueueCap + j]), __uint_as_float(qm_base[2 * kQueueCap + j]), __uint_as_float(qm_base[3 * kQueueCap + j]));
              P = V(__uint_as_float(qm_base[4 * kQueueCap + j]), __uint_as_float(qm_base[5 * kQueueCap + j]), __uint_as_float(qm_base[6 * kQueueCap + j]));
              index = (int)cand;
              bounces = (int)(w7q >> 3);
              const float ep = 1.0f * 1.0f * sc.inv_sqrt_bounce[(bounces - 1) & 7] * 5.0f;   // the colour the previous bounce left
              rgb = V((w7q & 1u) ? ep : 0.0f, (w7q & 2u) ? ep : 0.0f, (w7q & 4u) ? ep : 0.0f);
              prev = P; org = P;
              seq = (media ? 3 : 0) + 2 * (bounces - 1);
              caustics = false; new_point = true;
              h.type = 1; h.idx = (int)(w0q >> 27) & 7;
              state = ST_PRIMARY;
            }
          }
          if (!resumed) {
            index = (int)cand;
            ray = table_direction(__ldg(table + cand));
            rgb = rgb0;
            if (index < 100) {   // CAUSTICS_PHOTONS: aimed at the glass sphere, jittered, not re-normalised
              v3 aim = normalize(sub(V(sc.sph[0][0], sc.sph[0][1], sc.sph[0][2]), light));
              ray = add(aim, mul(ray, 0.01f));
            }
            prev = light; org = light;
            bounces = 1; seq = media ? 3 : 0;
            caustics = false; new_point = true;
            h.type = 0; h.idx = 0;
            state = ST_PRIMARY;
          }
        }
        const int took = __popc(idle) < avail ? __popc(idle) : (int)avail;
        if (kTwoPhase && two_phase) qm -= took; else cur += took;
      }
      if (__ballot_sync(0xffffffffu, state != ST_IDLE) == 0u) {
        if (((kTwoPhase && two_phase) ? (long long)qm : end - cur) <= 0) break;
        continue;
      }
      if (state == ST_IDLE) continue;

      // ---- the single intersection site ----
      raytrace<kStd>(sc, ray, org, h);

      // ---- mirror / glass chain: handleReflection/handleRefraction{,2,3,4}, PMK:673-827 ----
      if (state >= ST_CHAIN_R) {
        bool chain_done = false;
        if (state == ST_CHAIN_R) {
          if (!h.hit) chain_done = true;
          else {
            P = add(mul(ray, h.dist), P);
            if (!(h.type == 0 && h.idx == 0)) chain_done = true;
            else {
              ray = refract3(sc, ray, P, h.type, h.idx, P, 1.0f);
              P = add(mul(ray, 0.00001f), P);
              org = P; state = ST_CHAIN_F1;
            }
          }
        } else if (state == ST_CHAIN_F1) {
          P = add(mul(ray, h.dist), P);   // executed even on a miss
          if (!(h.hit && h.type == 0 && h.idx == 0)) chain_done = true;
          else {
            ray = refract3(sc, ray, P, h.type, h.idx, P, -1.0f);
            P = add(mul(ray, 0.00001f), P);
            org = P; state = ST_CHAIN_F2;
          }
        } else {   // ST_CHAIN_F2
          P = add(mul(ray, h.dist), P);
          if (level == 4 || !(h.type == 0 && h.idx == 1)) chain_done = true;   // not gated on h.hit: stale ids, as PMK:807
          else {
            level++;
            ray = reflect3<kStd>(sc, ray, prev, h.type, h.idx, P);
            org = P; state = ST_CHAIN_R;
          }
        }
        if (!chain_done) continue;
        caustics = chain_glass; new_point = false; bounces++;
        state = ST_PRIMARY;   // falls through: the while-condition is evaluated on the chain's last intersection
      }

      // ---- both remaining states deposit one photon at the intersection just found (single store site) ----
      v3 loc, e;
      if (state == ST_PRIMARY) {   // top of the bounce loop, PMK:1289-1301
        if (!(h.hit && bounces <= 5)) { state = ST_IDLE; continue; }
        if (new_point) P = add(mul(ray, h.dist), prev);
        if (caustics) rgb = mul(V(1.0f, 1.0f, 1.0f), 10.0f);
        else rgb = mul(mul(mul(get_color(rgb, h.type, h.idx), 1.0f), sc.inv_sqrt_bounce[bounces & 7]), 5.0f);   // 1 <= bounces <= 5 here
        loc = P; e = rgb;
      } else {                     // shadowPhoton, PMK:1185-1196: -0.25 at the next hit along the same ray
        loc = add(mul(ray, h.dist), org);
        e = V(-0.25f, -0.25f, -0.25f);
      }
      store_photon(sk, sa, h.type, h.idx, loc, e, state == ST_SHADOW);
      if (rec) append_record(sk, seq, 0, h.type, h.idx, index, loc, ray, e);
      seq++;
      if (state == ST_PRIMARY && !caustics) {
        t_type = h.type; t_idx = h.idx;
        org = add(P, mul(ray, 0.00001f));   // the shadow ray starts just beyond the hit
        state = ST_SHADOW;
        continue;
      }
      if (state == ST_SHADOW) { h.type = t_type; h.idx = t_idx; }   // dist/hit stay clobbered, as in the reference
      // PMK:1305-1370: where does the photon go next
      prev = P;
      if (h.type == 0 && h.idx == 1) {          // mirror sphere
        chain_glass = false; level = 1;
        ray = reflect3<kStd>(sc, ray, prev, h.type, h.idx, P);
        org = P; state = ST_CHAIN_R;
      } else if (h.type == 0 && h.idx == 0) {   // glass sphere
        chain_glass = true; level = 1;
        ray = refract3(sc, ray, P, h.type, h.idx, P, 1.0f);
        P = add(mul(ray, 0.00001f), P);
        org = P; state = ST_CHAIN_F1;
      } else {                                   // diffuse wall: reflect3(ray, prev, ...) with prev == the hit point
        // Hazard H1: the wall "normal" is normalize(e_axis * (prev.axis - offset)).  When that offset squares to 0
        // (the hit point lies exactly on the wall, ~88% of bounces) or is NaN, the normal, the reflected ray and hence
        // every intersection test of the next raytrace are NaN: no hit, the photon ends.  Skip straight to that outcome.
        const int wax = h.type == 1 ? (kStd ? std_axis(h.idx) : sc.pl_axis[h.idx]) : -1;
        const float wd = comp(prev, wax) - sc.pl_off[h.type == 1 ? h.idx : 0];
        const float wdd = wd * wd;
        if (h.type == 1 && wax >= 0 && wax <= 2 && (wdd == 0.0f || wdd != wdd)) { state = ST_IDLE; continue; }
        if (kStd) ray = reflect_wall_std(sc, ray, prev, h.idx, wax);
        else ray = reflect3<kStd>(sc, ray, prev, h.type, h.idx, P);
        org = P;
        caustics = false; new_point = true; bounces++;
        state = ST_PRIMARY;
      }
    }
    if (!(kTwoPhase && two_phase)) break;
  }

  // ---- flush the CTA-private accumulators ----
  if (dbg && (threadIdx.x & 31) == 0) dbg[8 + (threadIdx.x >> 5)] = global_ns();   // when this warp ran out of photons
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[2] = global_ns();
  if (sk.acc) {
    for (int e = threadIdx.x; e < kAccHitEntries; e += blockDim.x) {
      unsigned long long v = (unsigned long long)sa.lo[e] | ((unsigned long long)sa.hi[e] << 32);
      if ((e & 3) == 3) v += (unsigned long long)((long long)sa.shadow[e >> 2] * __float2ll_rn(-0.25f * (float)kHitScale));
      if (v) atomicAdd(sk.acc + e, v);
    }
  }
  if (dbg) {
    __syncthreads();
    if (threadIdx.x == 0) dbg[3] = global_ns();
  }
}

// ------------------------------------------------------------------------------------------------------
// self-test of fdiv_fastpath (pm_math.cuh) against the IEEE division, on the operand domain ray_walls_std argues about:
// numerator 0 or 2^-34 <= |a| <= 2^21, |b| < 4 (log-uniform over that exponent range, zeros and denormals included) or inf / NaN.
// (Each numerator also checks sqrt(RN(a * a)) == |a|, the identity behind the scalar wall normal of trace_kernel.)
// Claim checked per pair: if the IEEE quotient would be accepted by checkDistance (0 < q < 999999.9) the fast path returns the same
// bits; otherwise the fast path's value is rejected too.  out[0] = pairs violating the claim, out[1] = pairs with an accepted quotient, out[2..3] = operand bits of one violating pair.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void __launch_bounds__(256) selftest_fdiv_kernel(unsigned long long n, uint32_t seed, unsigned long long *out) {
  unsigned long long bad = 0, acc = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t h0 = mix32((uint32_t)i ^ seed), h1 = mix32(h0 + (uint32_t)(i >> 32) + 0x9e3779b9u), h2 = mix32(h1 ^ 0x85ebca6bu);
    // a: sign, exponent 2^-34 .. 2^21 (biased 93 .. 148), random mantissa; one in 64 is zero, one in 1024 NaN
    uint32_t abits = (h0 & 0x80000000u) | ((93u + (h0 >> 8) % 56u) << 23) | (h1 & 0x7fffffu);
    if ((h2 & 63u) == 0u) abits &= 0x80000000u;
    if ((h2 & 1023u) == 1u) abits = 0x7fc00000u;
    // b: half of the pairs put b where the quotient lands in (2^-12, 2^22) -- the accepted range and its edges --, the others anywhere
    uint32_t bbits;
    if (h2 & 0x40000000u) {
      const int ea = (int)((abits >> 23) & 255u), eb = ea - ((int)((h2 >> 10) % 34u) - 12);
      bbits = (h1 & 0x80000000u) | ((uint32_t)(eb < 0 ? 0 : eb) << 23) | (h2 >> 9 & 0x7fffffu ^ (h0 >> 3 & 0x7fffffu));
    } else {
      bbits = mix32(h2 + 0x632be59bu);
    }
    // b is a component of a (nearly) unit vector or garbage: |b| < 4, or inf / NaN
    if ((bbits >> 23 & 255u) > 128u && (bbits >> 23 & 255u) != 255u) bbits = (bbits & 0x807fffffu) | ((bbits >> 23 & 255u) % 129u) << 23;
    const float a = __uint_as_float(abits), b = __uint_as_float(bbits);
    const float q = __fdiv_rn(a, b), f = fdiv_fastpath(a, b);
    const bool q_ok = q > 0.0f && q < 999999.9f, f_ok = f > 0.0f && f < 999999.9f;
    acc += q_ok ? 1 : 0;
    if (a == a && __float_as_uint(__fsqrt_rn(a * a)) != (abits & 0x7fffffffu)) { bad++; out[2] = abits; out[3] = abits; }   // the wall normal's identity
    if (q_ok ? (__float_as_uint(q) != __float_as_uint(f)) : f_ok) { bad++; out[2] = abits; out[3] = bbits; }
  }
  if (bad) atomicAdd(out + 0, bad);
  if (acc) atomicAdd(out + 1, acc);
}
cudaError_t launch_selftest_fdiv(unsigned long long n, uint32_t seed, unsigned long long *out, int num_sms, cudaStream_t st) {
  selftest_fdiv_kernel<<<num_sms * 8, 256, 0, st>>>(n, seed, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------
cudaError_t preload_trace_kernels() {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, mwc_table_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, philox_table_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, table_norm_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, volume_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, fold_volume_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, trace_kernel<false, false>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, trace_kernel<true, false>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, trace_kernel<false, true>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, trace_kernel<true, true>);
  return e;
}

cudaError_t launch_mwc_table(float4 *table, long long first, long long last, long long n, uint32_t w0, uint32_t z0, const MwcJump *J,
                             cudaStream_t st) {
  long long threads = last - first + 3;
  if (n <= 0) return cudaSuccess;
  unsigned blocks = (unsigned)((threads + 255) / 256);
  mwc_table_kernel<<<blocks, 256, 0, st>>>(table, first, last, n, w0, z0, J);
  return cudaGetLastError();
}

cudaError_t launch_table_norm(float4 *table, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  table_norm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(table, n);
  return cudaGetLastError();
}

cudaError_t launch_philox_table(float4 *table, long long n, unsigned long long seed, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  philox_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(table, n, (uint32_t)seed, (uint32_t)(seed >> 32));
  return cudaGetLastError();
}

static Sink make_sink(unsigned flags, unsigned long long *acc, float4 *rec_pos, float4 *rec_pow, float4 *rec_dir,
                      unsigned long long *rec_count, long long rec_cap) {
  Sink sk;
  sk.vol_cnt = nullptr; sk.vrec_pos = sk.vrec_pow = nullptr; sk.vrec_cap = 0; sk.dbg = nullptr; sk.vox_touched = nullptr; sk.queue = nullptr;
  sk.acc = (flags & PM_TRACE_NO_MAP) ? nullptr : acc;
  sk.rec_pos = rec_pos; sk.rec_pow = rec_pow; sk.rec_dir = rec_dir; sk.rec_count = rec_count; sk.rec_cap = rec_cap;
  return sk;
}

static unsigned volume_blocks(long long n, int num_sms) {
  long long want = (n + 255) / 256, cap = (long long)num_sms * 8;
  return (unsigned)(want < cap ? want : cap);
}

static cudaError_t launch_fold(uint32_t *vol_cnt, unsigned long long *acc, cudaStream_t st) {
  fold_volume_kernel<<<PM_GRID_VOXELS / 256, 256, 0, st>>>(vol_cnt, acc);
  return cudaGetLastError();
}

// PM_TRACE_SPLIT: the medium walk as its own launch (+ the count fold)
int launch_trace_volume(const DeviceScene &sc, const float4 *table, long long first, long long last, unsigned flags, uint32_t w0,
                        uint32_t z0, const MwcJump *J, unsigned long long *acc, uint32_t *vol_cnt, float4 *rec_pos, float4 *rec_pow,
                        float4 *rec_dir, unsigned long long *rec_count, long long rec_cap, int num_sms, cudaStream_t st,
                        cudaError_t *err) {
  *err = cudaSuccess;
  long long n = last - first;
  if (n <= 0) return 0;
  Sink sk = make_sink(flags, acc, nullptr, nullptr, nullptr, rec_count, 0);
  sk.vol_cnt = vol_cnt; sk.vrec_pos = rec_pos; sk.vrec_pow = rec_pow; sk.vrec_cap = rec_cap;
  unsigned blocks = volume_blocks(n, num_sms);
  unsigned long long steps = 9ull * blocks * 256ull;
  volume_kernel<<<blocks, 256, 0, st>>>(sc, table, first, last, flags, w0, z0, J, host_powmod(18000u, steps, mwc_modulus(1)),
                                        host_powmod(36969u, steps, mwc_modulus(0)), sk);
  *err = cudaGetLastError();
  if (*err != cudaSuccess || !sk.acc) return 1;
  *err = launch_fold(vol_cnt, sk.acc, st);
  return 2;
}

// The trace launch.  vol_warps > 0: fused medium + surface walk (the volume records go to vrec_*); 0: surface walk only.
int launch_trace(const DeviceScene &sc, const float4 *table, long long first, long long last, unsigned flags, int vol_warps, uint32_t w0,
                 uint32_t z0, const MwcJump *J, unsigned long long *acc, uint32_t *vol_cnt, float4 *rec_pos, float4 *rec_pow,
                 float4 *rec_dir, float4 *vrec_pos, float4 *vrec_pow, long long vrec_cap, unsigned long long *rec_count, long long rec_cap,
                 int num_sms, cudaStream_t st, cudaError_t *err, unsigned long long *dbg, uint32_t *vox_touched, uint32_t *queue) {
  *err = cudaSuccess;
  long long n = last - first;
  if (n <= 0) return 0;
  Sink sk = make_sink(flags, acc, rec_pos, rec_pow, rec_dir, rec_count, rec_cap);
  sk.vol_cnt = vol_cnt; sk.vrec_pos = vrec_pos; sk.vrec_pow = vrec_pow; sk.vrec_cap = vrec_cap; sk.dbg = dbg; sk.vox_touched = vox_touched;
  sk.queue = (flags & PM_TRACE_ONE_PHASE) ? nullptr : queue;
  const bool rec = (flags & PM_TRACE_RECORDS) != 0;
  // the reference's object layout (only the layout: offsets, centres and radii stay scene data) selects the specialised instantiation
  bool std_scene = sc.n_spheres == 2 && sc.n_planes == 5;
  for (int i = 0; i < PM_MAX_PLANES; i++) std_scene = std_scene && sc.pl_axis[i] == std_axis(i);
  std_scene = std_scene && std_walls_ok(sc.pl_off);   // the branch-free wall test's range conditions (pm_math.cuh ray_walls_std)
#ifdef PM_NO_STD_SCENE
  std_scene = false;
#endif
  auto kernel = rec ? (std_scene ? trace_kernel<true, true> : trace_kernel<true, false>)
                    : (std_scene ? trace_kernel<false, true> : trace_kernel<false, false>);
  *err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSurfaceSmem);
  if (*err != cudaSuccess) return 0;
  const int cta_warps = kSurfaceThreads / 32;
  if (vol_warps < 0) vol_warps = 0;
  if (vol_warps > cta_warps / 2) vol_warps = cta_warps / 2;
  long long warps_needed = (n + 31) / 32;
  long long ctas = (warps_needed + (cta_warps - vol_warps) - 1) / (cta_warps - vol_warps);
  unsigned grid = (unsigned)(ctas < num_sms ? ctas : num_sms);
  // jump multipliers of the sliced medium walk (volume_walk_sliced): one MWC step forward is x * a, backward x * 2^16 (mod m)
  const long long S = cta_warps - vol_warps, per_cta = (n + grid - 1) / grid, per = (per_cta + S - 1) / S;
  const long long fwd = (long long)vol_warps * per, wrp = ((long long)vol_warps - S) * per + 32;
  auto mult = [](uint32_t a, long long photons, uint32_t m) {
    return photons >= 0 ? host_powmod(a, 9ull * (unsigned long long)photons, m) : host_powmod(65536u, 9ull * (unsigned long long)(-photons), m);
  };
  SliceJump jump;
  jump.fz = mult(36969u, fwd, mwc_modulus(0)); jump.fw = mult(18000u, fwd, mwc_modulus(1));
  jump.wz = mult(36969u, wrp, mwc_modulus(0)); jump.ww = mult(18000u, wrp, mwc_modulus(1));
  kernel<<<grid, kSurfaceThreads, kSurfaceSmem, st>>>(sc, table, first, last, flags, vol_warps, w0, z0, J, jump, sk);
  *err = cudaGetLastError();
  if (*err != cudaSuccess || !vol_warps || !sk.acc) return 1;
  *err = launch_fold(vol_cnt, sk.acc, st);
  return 2;
}

}  // namespace pm
