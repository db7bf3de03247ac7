// pm_api.cu -- host side of libpmb200.so: the context behind the C-ABI of include/pmb200.h.
//
// Mirrors the reference's host launch surface: the three extern "C" launchers of photonMappingKernel.cu
// (PMK:1523-1582) over a process-global default context, plus the handle-based extended API.  Host code is
// C++ calling the CUDA kernels of pm_trace.cu / pm_map.cu / pm_render.cu; no CPU compute path exists here.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "pm_context.h"

// ---- host-side MWC arithmetic (see pm_math.cuh MwcJump) ------------------------------------------------
static uint32_t h_mulmod(uint32_t a, uint32_t b, uint32_t m) { return host_mulmod(a, b, m); }
static uint32_t h_powmod(uint32_t a, unsigned long long e, uint32_t m) { return host_powmod(a, e, m); }
// one MWC lane is x -> x * a mod (a*2^16 - 1): (2^16)^-1 == a because a*2^16 == 1 mod m
static uint32_t h_lane_mult(int lane) { return lane == 0 ? 36969u : 18000u; }
static uint32_t h_mwc_advance(int lane, uint32_t x, unsigned long long steps) {
  uint32_t m = mwc_modulus(lane);
  return h_mulmod(x, h_powmod(h_lane_mult(lane), steps, m), m);
}
static bool mwc_state_ok(uint32_t w, uint32_t z) { return w != 0 && z != 0 && w < mwc_modulus(1) && z < mwc_modulus(0); }

static void fill_jump_tables(MwcJump *J) {
  for (int lane = 0; lane < 2; lane++) {
    uint32_t m = mwc_modulus(lane);
    uint32_t base = h_lane_mult(lane);
    for (int level = 0; level < 3; level++) {
      uint32_t cur = 1;
      for (int k = 0; k < 1024; k++) { J->pw[lane][level][k] = cur; cur = h_mulmod(cur, base, m); }
      base = cur;   // base^(1024)
    }
  }
}

// positionObjects, PMK:1380-1404, evaluated once on the host: cos(float)/sin(float) resolve to the float
// overloads, the other two arguments are double expressions.
static void position_objects(pm_scene *sc, float t) {
  if (!sc->animate) return;
  sc->spheres[0][0] = (float)(1.0 * (double)cosf(t));
  sc->spheres[0][1] = (float)(0.5 * sin(2.0 * (double)t));
  sc->spheres[0][2] = (float)((double)sinf(t) + 3.5);
  sc->spheres[1][0] = 0.0f;
  sc->spheres[1][1] = (float)(sin(0.5 * (double)t + 5.0) - 0.25);
  sc->spheres[1][2] = 3.5f;
}

// Terms of trace_kernel's two-phase surface walk (csrc/pm_trace.cu).
//   shadow_need[w], bit i clear: a shadow ray (PMK:1185-1196) behind wall w cannot be credited with sphere i, so the kernel skips that
//   raySphere -- provided the ray has crossed w's plane coming from the light's side (wall_side[w]; the kernel checks the sign of the ray
//   component) and starts within 16 of the origin (checked in the kernel as well).  Condition here: the sphere lies on the light's side
//   of the plane with a margin m = 0.05, i.e. side * (c[axis] - offset) - R >= m.  Argument: the shadow ray starts within 2e-5 of the
//   plane and moves away from it, so for t >= 0 it stays >= m - 2e-5 from every point of the sphere.  With s = c - o, t* = (s.r)/A:
//     t* >= 0: the line's closest approach is on the far side, its distance to the centre >= R + m - 2e-5, so the true discriminant is
//       <= -4A(2Rm + m^2) <= -0.03; the computed one differs by <= 3.3e-6 |s|^2 <= 0.006 (|s| <= 41.6), so D <= 0 and raySphere
//       returns without a candidate;
//     t* < 0 and the computed s.r <= 0: B >= 0 and C > 0 (the origin is outside the sphere by >= m: |s|^2 - R^2 >= 0.0075, error 3e-4),
//       so the root -B - sqrt(D) <= 0 is rejected by checkDistance whatever D is;
//     t* < 0 but the computed s.r > 0: then |s.r| <= 1.8e-7 |s| |r|, the closest approach lies within 1e-5 of the origin and the first
//       case applies.
//   The bounds used: light, sphere centres and wall offsets within [-8, 8], 0.05 <= R <= 4.
//   fast_ok: reference layout (2 spheres, 5 walls x,y,x,y,z, std_walls_ok), the bounds above, the light >= m from every wall plane and
//   outside both spheres with light_C >= 0 (the rejection test for a fresh photon's ray needs raySphere's sign = -1).
static void two_phase_terms(DeviceScene &d) {
  const float m = 0.05f;
  bool ok = d.n_spheres == 2 && d.n_planes == 5 && std_walls_ok(d.pl_off);
  for (int w = 0; w < PM_MAX_PLANES; w++) ok = ok && d.pl_axis[w] == std_axis(w) && fabsf(d.pl_off[w]) <= 8.0f;
  for (int j = 0; j < 3; j++) ok = ok && fabsf(d.light[j]) <= 8.0f;
  for (int i = 0; i < 2; i++) {
    // raySphere's ray-independent terms for rays that start at the light, with the kernel's own FP32 operations (one subtraction per
    // component; x*x + y*y + z*z summed left to right, then - r^2; volatile keeps the compiler from contracting or widening anything)
    volatile float s0 = d.sph[i][0] - d.light[0], s1 = d.sph[i][1] - d.light[1], s2 = d.sph[i][2] - d.light[2];
    volatile float p0 = s0 * s0, p1 = s1 * s1, p2 = s2 * s2;
    volatile float sum = p0 + p1;
    sum = sum + p2;
    d.light_s[i][0] = s0; d.light_s[i][1] = s1; d.light_s[i][2] = s2;
    d.light_C[i] = sum - d.sph_r2[i];
    ok = ok && d.light_C[i] >= 0.0f && d.sph[i][3] >= 0.05f && d.sph[i][3] <= 4.0f;
    for (int j = 0; j < 3; j++) ok = ok && fabsf(d.sph[i][j]) <= 8.0f;
  }
  for (int w = 0; w < PM_MAX_PLANES; w++) {
    d.shadow_need[w] = 3u;
    d.wall_side[w] = 1.0f;
    if (!ok) continue;
    const int a = d.pl_axis[w];
    const float gap = d.light[a] - d.pl_off[w];
    if (!(fabsf(gap) >= m)) { ok = false; continue; }
    d.wall_side[w] = gap > 0.0f ? 1.0f : -1.0f;
    for (int i = 0; i < 2; i++)
      if (d.wall_side[w] * (d.sph[i][a] - d.pl_off[w]) - d.sph[i][3] >= m) d.shadow_need[w] &= ~(1u << i);
  }
  d.fast_ok = ok ? 1 : 0;
}

static DeviceScene make_device_scene(const pm_scene &in, float t) {
  pm_scene s = in;
  position_objects(&s, t);
  DeviceScene d;
  memset(&d, 0, sizeof(d));
  d.n_spheres = std::min(std::max(s.n_spheres, 0), PM_MAX_SPHERES);
  d.n_planes = std::min(std::max(s.n_planes, 0), PM_MAX_PLANES);
  for (int i = 0; i < PM_MAX_SPHERES; i++) {
    for (int j = 0; j < 4; j++) d.sph[i][j] = s.spheres[i][j];
    d.sph_r2[i] = s.spheres[i][3] * s.spheres[i][3];   // pow(radius, 2.0f), PMK:120
  }
  for (int i = 0; i < PM_MAX_PLANES; i++) { d.pl_axis[i] = (int)s.planes[i][0]; d.pl_off[i] = s.planes[i][1]; }
  for (int b = 0; b < 8; b++) d.inv_sqrt_bounce[b] = 1.0f / sqrtf((float)b);   // IEEE sqrt and division, as __fsqrt_rn / __fdiv_rn
  for (int j = 0; j < 3; j++) d.light[j] = s.light[j];
  two_phase_terms(d);
  d.sz_img = (float)s.sz_img;
  d.cam_ox = s.cam_ox; d.cam_oy = s.cam_oy;
  return d;
}

extern "C" {

int pm_peer_disconnect(pm_context *c);

const char *pm_version(void) { return "pmb200 0.1 (sm_100a)"; }

void pm_scene_default(pm_scene *sc) {
  static const float sp[3][4] = {{1.0f, -1.0f, 1.0f, 0.4f}, {-0.6f, -1.0f, 4.5f, 0.4f}, {0.0f, 0.0f, 1.5f, 1.0f}};
  static const float pl[5][2] = {{0, 1.5f}, {1, -1.5f}, {0, -1.5f}, {1, 1.5f}, {2, 6.0f}};
  memset(sc, 0, sizeof(*sc));
  sc->n_spheres = 2; sc->n_planes = 5;
  memcpy(sc->spheres, sp, sizeof(sp)); memcpy(sc->planes, pl, sizeof(pl));
  sc->light[0] = 0.0f; sc->light[1] = 1.4f; sc->light[2] = 3.5f;
  sc->sz_img = 512; sc->cam_ox = 0.0f; sc->cam_oy = 0.0f; sc->animate = 1;
}

int pm_trace_plan(const pm_scene *in, float t, int32_t *two_phase, uint32_t *shadow_need5) {
  if (!in || !two_phase || !shadow_need5) return PM_ERR_ARG;
  const DeviceScene d = make_device_scene(*in, t);
  *two_phase = d.fast_ok;
  for (int w = 0; w < PM_MAX_PLANES; w++) shadow_need5[w] = d.shadow_need[w];
  return PM_OK;
}

int pm_position_objects(const pm_scene *in, float t, pm_scene *out) {
  if (!in || !out) return PM_ERR_ARG;
  *out = *in;
  position_objects(out, t);
  return PM_OK;
}

int pm_create(pm_context **out, int device) {
  if (!out) return PM_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return PM_ERR_NO_DEVICE;
  }
  pm_context *c = new pm_context();
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  c->device = device;
  pm_scene_default(&c->scene);
  c->dsc = make_device_scene(c->scene, 0.0f);
  auto fail = [&](cudaError_t e) { fprintf(stderr, "pmb200: pm_create: %s\n", cudaGetErrorString(e)); pm_destroy(c); return PM_ERR_CUDA; };
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e);
  if ((e = cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&c->d_xchg, kExchangeBytes)) != cudaSuccess) return fail(e);
  if ((e = cudaMemset(c->d_xchg, 0, kExchangeBytes)) != cudaSuccess) return fail(e);
  c->d_acc = (long long *)((ExchangeHeader *)c->d_xchg + 1);
  memset(&c->pv, 0, sizeof(c->pv));
  c->pv.world = 1; c->pv.rank = 0; c->pv.hdr[0] = (ExchangeHeader *)c->d_xchg; c->pv.acc[0] = c->d_acc;
  c->pv.timeout_ns = 5000000000ull;
  for (int b = 0; b < kAccBuffers; b++) {
    if ((e = cudaEventCreateWithFlags(&c->ev_reduced[b], cudaEventDisableTiming)) != cudaSuccess) return fail(e);
    if ((e = cudaEventCreateWithFlags(&c->ev_cleared[b], cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  }
  if ((e = cudaMalloc(&c->d_grid, sizeof(float) * PM_GRID_FLOATS)) != cudaSuccess) return fail(e);
  for (int b = 0; b < 2; b++) {
    if ((e = cudaMalloc(&c->d_vol_buf[b], sizeof(float4) * kVolTableEntries)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&c->d_surf_buf[b], sizeof(float4) * kSurfTableEntries)) != cudaSuccess) return fail(e);
    if ((e = cudaEventCreateWithFlags(&c->ev_tbl_read[b], cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  }
  if ((e = cudaEventCreateWithFlags(&c->ev_tbl_built, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  c->d_vol = c->d_vol_buf[0]; c->d_surf = c->d_surf_buf[0];
  if ((e = cudaMalloc(&c->d_jump, sizeof(MwcJump))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&c->d_rec_count, sizeof(unsigned long long))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&c->d_vol_cnt, sizeof(uint32_t) * kVolCntEntries)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&c->d_trace_queue, sizeof(uint32_t) * kTraceQueueWordsPerCta * (size_t)c->num_sms)) != cudaSuccess) return fail(e);
  if ((e = cudaMemset(c->d_vol_cnt, 0, sizeof(uint32_t) * kVolCntEntries)) != cudaSuccess) return fail(e);
  if ((e = cudaMemset(c->d_grid, 0, sizeof(float) * PM_GRID_FLOATS)) != cudaSuccess) return fail(e);
  if ((e = cudaMemset(c->d_rec_count, 0, sizeof(unsigned long long))) != cudaSuccess) return fail(e);
  {
    std::vector<MwcJump> J(1);
    fill_jump_tables(J.data());
    if ((e = cudaMemcpy(c->d_jump, J.data(), sizeof(MwcJump), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
  }
  *out = c;
  int rc = pm_set_photon_count(c, 10000);   // nrPhotons, PMK:29
  if (rc != PM_OK) { pm_destroy(c); *out = nullptr; return rc; }
  return PM_OK;
}

int pm_destroy(pm_context *c) {
  if (!c) return PM_ERR_ARG;
  cudaSetDevice(c->device);
  pm_peer_disconnect(c);
  cudaFree(c->d_table); cudaFree(c->d_xchg); cudaFree(c->d_acc_sum); cudaFree(c->d_vol_cnt); cudaFree(c->d_grid);
  for (int b = 0; b < 2; b++) {
    cudaFree(c->d_vol_buf[b]); cudaFree(c->d_surf_buf[b]);
    if (c->ev_tbl_read[b]) cudaEventDestroy(c->ev_tbl_read[b]);
  }
  if (c->ev_tbl_built) cudaEventDestroy(c->ev_tbl_built);
  cudaFree(c->d_jump); cudaFree(c->d_rec_pos); cudaFree(c->d_rec_pow); cudaFree(c->d_rec_dir); cudaFree(c->d_rec_count);
  cudaFree(c->d_fb_u8); cudaFree(c->d_fb_f32); cudaFree(c->d_vrec_pos); cudaFree(c->d_vrec_pow); cudaFree(c->d_trace_dbg); cudaFree(c->d_trace_queue);
  for (int k = 0; k < pm_context::kFrameRing; k++) {
    cudaFree(c->d_fb_async[k]);
    if (c->ev_rendered[k]) cudaEventDestroy(c->ev_rendered[k]);
    if (c->ev_copied[k]) cudaEventDestroy(c->ev_copied[k]);
  }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->render_stream) cudaStreamDestroy(c->render_stream);
  if (c->ev_traced) cudaEventDestroy(c->ev_traced);
  for (int b = 0; b < kAccBuffers; b++) {
    if (c->ev_reduced[b]) cudaEventDestroy(c->ev_reduced[b]);
    if (c->ev_cleared[b]) cudaEventDestroy(c->ev_cleared[b]);
  }
  knn_free(c->knn[0]); knn_free(c->knn[1]); cudaFree(c->d_work);
  for (auto &sp : c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  delete c;
  return PM_OK;
}

int pm_enable_timing(pm_context *c, bool on) {
  if (!c) return PM_ERR_ARG;
  c->timing = on; c->spans_used = 0;
  return PM_OK;
}
int pm_kernel_count(void) { return K_COUNT; }
const char *pm_kernel_name(int kind) { return kind >= 0 && kind < K_COUNT ? kKernelNames[kind] : ""; }
int pm_get_timings(pm_context *c, double *total_ms, int64_t *launches) {
  ARG(c, c && total_ms && launches, "null argument");
  CK(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < K_COUNT; k++) { total_ms[k] = 0.0; launches[k] = 0; }
  for (size_t i = 0; i < c->spans_used; i++) {
    float ms = 0.0f;
    CK(c, cudaEventElapsedTime(&ms, c->spans[i].a, c->spans[i].b));
    total_ms[c->spans[i].kind] += ms; launches[c->spans[i].kind]++;
  }
  c->spans_used = 0;
  return PM_OK;
}

const char *pm_last_error(const pm_context *c) { return c ? c->err.c_str() : "null context"; }
int64_t pm_launch_count(const pm_context *c) { return c ? c->launches : 0; }

int pm_set_stream(pm_context *c, void *stream) {
  if (!c) return PM_ERR_ARG;
  c->stream = (cudaStream_t)stream;
  return PM_OK;
}
int pm_sync(pm_context *c) {
  if (!c) return PM_ERR_ARG;
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));
  if (c->aux_stream) CK(c, cudaStreamSynchronize(c->aux_stream));   // the later stages of pipelined frames
  if (c->render_stream) CK(c, cudaStreamSynchronize(c->render_stream));
  return PM_OK;
}

int pm_set_scene(pm_context *c, const pm_scene *s) {
  ARG(c, c && s, "null argument");
  ARG(c, s->n_spheres >= 0 && s->n_spheres <= PM_MAX_SPHERES && s->n_planes >= 0 && s->n_planes <= PM_MAX_PLANES, "object counts out of range");
  ARG(c, s->sz_img > 0, "sz_img must be positive");
  c->scene = *s;
  return PM_OK;
}
int pm_get_scene(const pm_context *c, pm_scene *s) {
  if (!c || !s) return PM_ERR_ARG;
  *s = c->scene;
  return PM_OK;
}
int pm_set_energy_scale(pm_context *c, float scale) {
  if (!c) return PM_ERR_ARG;
  c->energy_scale = scale;
  return PM_OK;
}

int pm_set_photon_count(pm_context *c, int64_t n) {
  ARG(c, c != nullptr, "null context");
  ARG(c, n >= 3 && n <= (int64_t)100000000, "photon count must be in [3, 1e8]");   // 9*n medium draws must stay < 2^30
  CK(c, cudaSetDevice(c->device));
  if (n > c->table_cap) {
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_table); c->d_table = nullptr; c->table_cap = 0;
    CK(c, cudaMalloc(&c->d_table, sizeof(float4) * (size_t)n));
    c->table_cap = n;
  }
  // the reference's table is zero-initialised device memory until launch_init_random_numbers_kernel runs
  CK(c, cudaMemsetAsync(c->d_table, 0, sizeof(float4) * (size_t)n, c->stream));
  CK(c, launch_table_norm(c->d_table, n, c->stream));   // w = 1/|(0,0,0)| = inf: a zero row still normalises to NaN
  c->n_photons = n; c->first = 0; c->last = n;
  c->table_first = 0; c->table_last = n;   // zero rows are valid rows (the reference's uninitialised table)
  return PM_OK;
}
int pm_set_photon_range(pm_context *c, int64_t first, int64_t last) {
  ARG(c, c != nullptr, "null context");
  ARG(c, first >= 0 && first <= last && last <= c->n_photons, "photon range outside [0, photon count]");
  c->first = first; c->last = last;
  return PM_OK;
}

int pm_set_mwc_state(pm_context *c, uint32_t w, uint32_t z) {
  ARG(c, c != nullptr, "null context");
  ARG(c, mwc_state_ok(w, z), "MWC state must be non-zero and below a*65536-1");
  c->mwc_w = w; c->mwc_z = z;
  return PM_OK;
}
int pm_get_mwc_state(const pm_context *c, uint32_t *w, uint32_t *z) {
  if (!c || !w || !z) return PM_ERR_ARG;
  *w = c->mwc_w; *z = c->mwc_z;
  return PM_OK;
}

int pm_init_random_table(pm_context *c) {
  ARG(c, c != nullptr, "null context");
  CK(c, cudaSetDevice(c->device));
  {
    SpanGuard g(c, K_MWC_TABLE);
    // only the rows this context traces (its photon range) and rows 0..2 are generated; the stream still advances by 3*n
    CK(c, launch_mwc_table(c->d_table, c->first, c->last, c->n_photons, c->mwc_w, c->mwc_z, c->d_jump, c->stream));
    c->table_first = c->first; c->table_last = c->last;
  }
  c->launches++;
  unsigned long long draws = 3ull * (unsigned long long)c->n_photons;
  c->mwc_z = h_mwc_advance(0, c->mwc_z, draws);
  c->mwc_w = h_mwc_advance(1, c->mwc_w, draws);
  return PM_OK;
}
int pm_init_random_table_philox(pm_context *c, uint64_t seed) {
  ARG(c, c != nullptr, "null context");
  CK(c, cudaSetDevice(c->device));
  {
    SpanGuard g(c, K_MWC_TABLE);
    CK(c, launch_philox_table(c->d_table, c->n_photons, seed, c->stream));
  }
  c->table_first = 0; c->table_last = c->n_photons;
  c->launches++;
  return PM_OK;
}
int pm_set_random_table_host(pm_context *c, const float *xyz, int64_t n) {
  ARG(c, c && xyz, "null argument");
  ARG(c, n == c->n_photons, "table length must equal the photon count");
  CK(c, cudaSetDevice(c->device));
  std::vector<float4> rows((size_t)n);
  for (int64_t i = 0; i < n; i++) rows[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0f);
  CK(c, cudaMemcpyAsync(c->d_table, rows.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CK(c, launch_table_norm(c->d_table, n, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  c->table_first = 0; c->table_last = n;
  return PM_OK;
}
int pm_get_random_table_host(pm_context *c, float *xyz, int64_t n) {
  ARG(c, c && xyz, "null argument");
  ARG(c, n >= 0 && n <= c->n_photons, "table length exceeds the photon count");
  CK(c, cudaSetDevice(c->device));
  std::vector<float4> rows((size_t)n);
  CK(c, cudaMemcpyAsync(rows.data(), c->d_table, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  for (int64_t i = 0; i < n; i++) { xyz[3 * i] = rows[i].x; xyz[3 * i + 1] = rows[i].y; xyz[3 * i + 2] = rows[i].z; }
  return PM_OK;
}

int pm_clear_map(pm_context *c) {
  ARG(c, c != nullptr, "null context");
  CK(c, cudaSetDevice(c->device));
  // Frames rotate through the accumulator buffers: the pipelined frame calls build the map of frame f on a second stream while
  // this stream already clears and traces frame f+1, and with peers connected other ranks may still be reading frame f's buffer.
  c->frame_no++;
  c->cur = (int)(c->frame_no % kAccBuffers);
  c->d_acc = (long long *)((ExchangeHeader *)c->d_xchg + 1) + (size_t)c->cur * kAccStride;
  c->acc_summed = false;
  if (c->precleared[c->cur]) {   // pm_build_map of two frames ago has already cleared it (off this stream): wait for that only
    CK(c, cudaStreamWaitEvent(c->stream, c->ev_cleared[c->cur], 0));
    c->precleared[c->cur] = false;
  } else {
    // this buffer was last used three frames ago: our own reader of it (map build or reduce) has run ...
    CK(c, cudaStreamWaitEvent(c->stream, c->ev_reduced[c->cur], 0));
    // ... and every PEER has finished reading it once our reduce of the frame after that one has run (see pm_peer.cu): two frames old
    if (c->world > 1 && c->frame_no >= 2) CK(c, cudaStreamWaitEvent(c->stream, c->ev_reduced[(c->frame_no - 2) % kAccBuffers], 0));
    CK(c, cudaMemsetAsync(c->d_acc, 0, sizeof(long long) * kAccStride, c->stream));   // the buffer and its flag words
  }
  if (c->d_rec_count) CK(c, cudaMemsetAsync(c->d_rec_count, 0, sizeof(unsigned long long), c->stream));
  c->vrec_count = 0;
  c->tables_valid = false;
  return PM_OK;
}

// the record buffers a k-NN map points into (positions, powers) are about to be rewritten or freed
static void records_changed(pm_context *c, int which) {
  if (c->knn_from_records[which] && c->knn[which].n > 0) c->knn_stale[which] = true;
}
static int knn_usable(pm_context *c, int which) {
  if (c->knn_stale[which]) { c->err = "the photon records changed since pm_knn_build: rebuild the map"; return PM_ERR_STATE; }
  return PM_OK;
}

int pm_set_record_capacity(pm_context *c, int64_t cap) {
  ARG(c, c != nullptr && cap >= 0, "bad capacity");
  records_changed(c, PM_MAP_SURFACE);
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->d_rec_pos); cudaFree(c->d_rec_pow); cudaFree(c->d_rec_dir);
  c->d_rec_pos = c->d_rec_pow = c->d_rec_dir = nullptr; c->rec_cap = 0;
  if (cap > 0) {
    CK(c, cudaMalloc(&c->d_rec_pos, sizeof(float4) * (size_t)cap));
    CK(c, cudaMalloc(&c->d_rec_pow, sizeof(float4) * (size_t)cap));
    CK(c, cudaMalloc(&c->d_rec_dir, sizeof(float4) * (size_t)cap));
    c->rec_cap = cap;
  }
  CK(c, cudaMemsetAsync(c->d_rec_count, 0, sizeof(unsigned long long), c->stream));
  return PM_OK;
}

int pm_set_volume_warps(pm_context *c, int warps) {
  ARG(c, c != nullptr && warps >= 1 && warps <= 16, "volume warps must be 1..16");
  c->vol_warps = warps;
  return PM_OK;
}

int pm_set_trace_sms(pm_context *c, int sms) {
  ARG(c, c != nullptr && sms >= 0, "bad SM count");
  c->trace_sms = sms;
  return PM_OK;
}

int pm_trace(pm_context *c, float t, unsigned flags) {
  ARG(c, c != nullptr, "null context");
  if ((flags & PM_TRACE_RECORDS) && c->rec_cap == 0) { c->err = "PM_TRACE_RECORDS needs pm_set_record_capacity first"; return PM_ERR_STATE; }
  if (c->first < c->table_first || c->last > c->table_last) {
    c->err = "the photon range was widened after pm_init_random_table generated only the previous range: set the range first";
    return PM_ERR_STATE;
  }
  CK(c, cudaSetDevice(c->device));
  c->dsc = make_device_scene(c->scene, t);
  cudaError_t terr = cudaSuccess;
  if (flags & PM_TRACE_RECORDS) {   // the record buffers are rewritten: maps built over them no longer match
    records_changed(c, PM_MAP_SURFACE);
    if (flags & PM_TRACE_MEDIA) records_changed(c, PM_MAP_VOLUME);
  }
  if (flags & PM_TRACE_MEDIA) {
    if (flags & PM_TRACE_RECORDS) {   // the volume set has a fixed slot per (photon, step)
      int64_t need = 3 * (c->last - c->first);
      if (need > c->vrec_cap) {
        CK(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_vrec_pos); cudaFree(c->d_vrec_pow); c->d_vrec_pos = c->d_vrec_pow = nullptr; c->vrec_cap = 0;
        CK(c, cudaMalloc(&c->d_vrec_pos, sizeof(float4) * (size_t)need));
        CK(c, cudaMalloc(&c->d_vrec_pow, sizeof(float4) * (size_t)need));
        c->vrec_cap = need;
      }
      c->vrec_count = need;
    }
    if (flags & PM_TRACE_SPLIT) {
      SpanGuard g(c, K_VOLUME);
      c->launches += launch_trace_volume(c->dsc, c->d_table, c->first, c->last, flags, c->mwc_w, c->mwc_z, c->d_jump,
                                         (unsigned long long *)c->d_acc, c->d_vol_cnt, c->d_vrec_pos, c->d_vrec_pow, nullptr,
                                         c->d_rec_count, c->vrec_cap, c->num_sms, c->stream, &terr);
    }
  }
  CK(c, terr);
  {
    SpanGuard g(c, K_SURFACE);
    const int vol_warps = ((flags & PM_TRACE_MEDIA) && !(flags & PM_TRACE_SPLIT)) ? c->vol_warps : 0;
    c->launches += launch_trace(c->dsc, c->d_table, c->first, c->last, flags, vol_warps, c->mwc_w, c->mwc_z, c->d_jump,
                                (unsigned long long *)c->d_acc, c->d_vol_cnt, c->d_rec_pos, c->d_rec_pow, c->d_rec_dir, c->d_vrec_pos,
                                c->d_vrec_pow, c->vrec_cap, c->d_rec_count, c->rec_cap,
                                (c->trace_sms > 0 && c->trace_sms < c->num_sms) ? c->trace_sms : c->num_sms, c->stream, &terr, c->d_trace_dbg,
                                (uint32_t *)(c->d_acc + kAccEntries), c->d_trace_queue);
  }
  CK(c, terr);
  if (flags & PM_TRACE_MEDIA) {   // the medium scattering consumed 9 draws per photon of the WHOLE job
    unsigned long long draws = 9ull * (unsigned long long)c->n_photons;
    c->mwc_z = h_mwc_advance(0, c->mwc_z, draws);
    c->mwc_w = h_mwc_advance(1, c->mwc_w, draws);
  }
  c->tables_valid = false;
  return PM_OK;
}

int pm_trace_profile(pm_context *c, bool on) {
  ARG(c, c != nullptr, "null context");
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->d_trace_dbg); c->d_trace_dbg = nullptr;
  if (on) {
    CK(c, cudaMalloc(&c->d_trace_dbg, sizeof(unsigned long long) * kTraceDbgWords * (size_t)c->num_sms));
    CK(c, cudaMemset(c->d_trace_dbg, 0, sizeof(unsigned long long) * kTraceDbgWords * (size_t)c->num_sms));
  }
  return PM_OK;
}
int pm_get_trace_profile_host(pm_context *c, uint64_t *out, int64_t max_words, int64_t *words) {
  ARG(c, c && out && words, "null argument");
  if (!c->d_trace_dbg) { c->err = "pm_trace_profile is off"; return PM_ERR_STATE; }
  *words = (int64_t)kTraceDbgWords * c->num_sms;
  ARG(c, max_words >= *words, "host buffer too small");
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));
  CK(c, cudaMemcpy(out, c->d_trace_dbg, sizeof(unsigned long long) * (size_t)*words, cudaMemcpyDeviceToHost));
  return PM_OK;
}

int pm_selftest_fdiv(pm_context *c, uint64_t pairs, uint32_t seed, uint64_t *violations, uint64_t *accepted) {
  ARG(c, c && violations && accepted, "null argument");
  CK(c, cudaSetDevice(c->device));
  unsigned long long *d = nullptr, h[4] = {0, 0, 0, 0};
  CK(c, cudaMalloc(&d, sizeof(h)));
  cudaError_t e = cudaMemsetAsync(d, 0, sizeof(h), c->stream);
  if (e == cudaSuccess) e = launch_selftest_fdiv(pairs, seed, d, c->num_sms, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  CK(c, e);
  *violations = h[0]; *accepted = h[1];
  if (h[0]) { static char msg[96]; snprintf(msg, sizeof(msg), "fdiv_fastpath differs, e.g. a = 0x%08llx, b = 0x%08llx", h[2], h[3]); c->err = msg; }
  return PM_OK;
}

int pm_accumulators(pm_context *c, void **dev_ptr, size_t *n) {
  if (!c || !dev_ptr || !n) return PM_ERR_ARG;
  *dev_ptr = c->d_acc; *n = kAccEntries;
  return PM_OK;
}

int pm_get_accumulators_host(pm_context *c, int64_t *out) {
  ARG(c, c && out, "null argument");
  CK(c, cudaSetDevice(c->device));
  if (c->aux_stream) CK(c, cudaStreamSynchronize(c->aux_stream));
  if (c->render_stream) CK(c, cudaStreamSynchronize(c->render_stream));
  CK(c, cudaMemcpyAsync(out, c->acc_summed ? c->d_acc_sum : c->d_acc, sizeof(long long) * kAccEntries, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return PM_OK;
}

int pm_build_map(pm_context *c) {
  ARG(c, c != nullptr, "null context");
  CK(c, cudaSetDevice(c->device));
  const long long *src = c->d_acc;
  if (c->world > 1) {   // the path's one exchange: pull and sum every rank's accumulators over peer memory (pm_peer.cu)
    SpanGuard g(c, K_PEER_REDUCE);
    c->pv.hdr[c->rank] = (ExchangeHeader *)c->d_xchg;
    // ranks that share a device (tests) must leave SMs for each other's trace while they spin
    // ... and while the persistent trace kernel of the next frame holds all but a few SMs (pm_set_trace_sms) the spinning blocks
    // must not fill every free one either: the render of the previous frame runs beside this kernel (render_stream)
    int blocks = c->num_sms;
    if (c->trace_sms > 0 && c->trace_sms < c->num_sms) blocks = std::max(16, std::min(c->num_sms, 4 * (c->num_sms - c->trace_sms)));
    if (c->peer_on_same_device) blocks = 16;
    CK(c, launch_peer_reduce(c->pv, c->cur, ++c->seq[0], c->d_acc_sum, blocks, c->stream));
    CK(c, cudaEventRecord(c->ev_reduced[c->cur], c->stream));
    c->launches++;
    c->acc_summed = true;
    src = c->d_acc_sum;
  }
  {
    SpanGuard g(c, K_BUILD_MAP);
    CK(c, launch_build_map(src, c->energy_scale, c->d_grid, c->stream));
  }
  if (c->world == 1) CK(c, cudaEventRecord(c->ev_reduced[c->cur], c->stream));   // the last reader of this accumulator buffer
  // The PREVIOUS frame's buffer can be cleared for its next use (two frames from now) as soon as this frame's reader has run: without
  // peers its own reader ran earlier; with peers our reduce of this frame has seen every peer's signal, which each peer sends only
  // after its reads of the previous frame's buffers.  Doing it here keeps the 1.2 MB memset off the stream the next trace runs on.
  if (c->frame_no >= 1 && c->preclear_frame != c->frame_no) {
    const int prev = (int)((c->frame_no - 1) % kAccBuffers);
    CK(c, cudaStreamWaitEvent(c->stream, c->ev_reduced[prev], 0));
    CK(c, cudaMemsetAsync((long long *)((ExchangeHeader *)c->d_xchg + 1) + (size_t)prev * kAccStride, 0, sizeof(long long) * kAccStride, c->stream));
    CK(c, cudaEventRecord(c->ev_cleared[prev], c->stream));
    c->precleared[prev] = true;
    c->preclear_frame = c->frame_no;
  }
  {
    // into the table buffer the renders are NOT reading: the render of the previous frame may still be running (render_stream)
    const int nb = c->tbl ^ 1;
    CK(c, cudaStreamWaitEvent(c->stream, c->ev_tbl_read[nb], 0));   // the last render that read this buffer (two frames ago)
    {
      SpanGuard g(c, K_BUILD_TABLES);
      CK(c, launch_build_tables(c->d_grid, c->d_vol_buf[nb], c->d_surf_buf[nb], c->stream));
    }
    CK(c, cudaEventRecord(c->ev_tbl_built, c->stream));
    c->tbl = nb; c->d_vol = c->d_vol_buf[nb]; c->d_surf = c->d_surf_buf[nb];
  }
  c->launches += 2;
  c->tables_valid = true;
  return PM_OK;
}
int pm_get_map_host(pm_context *c, float *grid) {
  ARG(c, c && grid, "null argument");
  CK(c, cudaSetDevice(c->device));
  if (c->aux_stream) CK(c, cudaStreamSynchronize(c->aux_stream));
  if (c->render_stream) CK(c, cudaStreamSynchronize(c->render_stream));
  CK(c, cudaMemcpyAsync(grid, c->d_grid, sizeof(float) * PM_GRID_FLOATS, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return PM_OK;
}
int pm_set_map_host(pm_context *c, const float *grid) {
  ARG(c, c && grid, "null argument");
  CK(c, cudaSetDevice(c->device));
  if (c->aux_stream) CK(c, cudaStreamSynchronize(c->aux_stream));         // pipelined frames still building / rendering from the map
  if (c->render_stream) CK(c, cudaStreamSynchronize(c->render_stream));
  CK(c, cudaMemcpyAsync(c->d_grid, grid, sizeof(float) * PM_GRID_FLOATS, cudaMemcpyHostToDevice, c->stream));
  CK(c, launch_build_tables(c->d_grid, c->d_vol, c->d_surf, c->stream));
  CK(c, cudaEventRecord(c->ev_tbl_built, c->stream));
  c->launches++;
  CK(c, cudaStreamSynchronize(c->stream));
  c->tables_valid = true;
  return PM_OK;
}
int pm_map_device(pm_context *c, float **dev_grid) {
  if (!c || !dev_grid) return PM_ERR_ARG;
  *dev_grid = c->d_grid;
  return PM_OK;
}

static int surface_record_count(pm_context *c, int64_t *n) {
  unsigned long long cnt = 0;
  CK(c, cudaMemcpyAsync(&cnt, c->d_rec_count, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  *n = (int64_t)cnt;
  return PM_OK;
}
int pm_record_count(pm_context *c, int64_t *n) {
  ARG(c, c && n, "null argument");
  CK(c, cudaSetDevice(c->device));
  int64_t s = 0;
  int rc = surface_record_count(c, &s);
  if (rc != PM_OK) return rc;
  *n = s + c->vrec_count;
  return PM_OK;
}
int pm_record_buffers(pm_context *c, int which, float **pos, float **pow, float **dir, int64_t *count) {
  ARG(c, c && (which == PM_MAP_SURFACE || which == PM_MAP_VOLUME), "bad map id");
  if (which == PM_MAP_SURFACE) {
    if (pos) *pos = (float *)c->d_rec_pos;
    if (pow) *pow = (float *)c->d_rec_pow;
    if (dir) *dir = (float *)c->d_rec_dir;
    if (count) {
      int rc = surface_record_count(c, count);
      if (rc != PM_OK) return rc;
      if (*count > c->rec_cap) { c->err = "record buffers overflowed (raise pm_set_record_capacity)"; return PM_ERR_STATE; }
    }
  } else {
    if (pos) *pos = (float *)c->d_vrec_pos;
    if (pow) *pow = (float *)c->d_vrec_pow;
    if (dir) *dir = nullptr;
    if (count) *count = c->vrec_count;
  }
  return PM_OK;
}
int pm_get_records_host(pm_context *c, pm_record *out, int64_t max_records) {
  ARG(c, c && out, "null argument");
  CK(c, cudaSetDevice(c->device));
  int64_t ns = 0;
  int rc = surface_record_count(c, &ns);
  if (rc != PM_OK) return rc;
  if (ns > c->rec_cap) { c->err = "record buffers overflowed (raise pm_set_record_capacity)"; return PM_ERR_STATE; }
  const int64_t nv = c->vrec_count, n = ns + nv;
  ARG(c, n <= max_records, "host buffer too small");
  std::vector<float4> pos((size_t)n), pw((size_t)n), dir((size_t)n, make_float4(0.f, 0.f, 0.f, 0.f));
  if (ns) {
    CK(c, cudaMemcpy(pos.data(), c->d_rec_pos, sizeof(float4) * (size_t)ns, cudaMemcpyDeviceToHost));
    CK(c, cudaMemcpy(pw.data(), c->d_rec_pow, sizeof(float4) * (size_t)ns, cudaMemcpyDeviceToHost));
    CK(c, cudaMemcpy(dir.data(), c->d_rec_dir, sizeof(float4) * (size_t)ns, cudaMemcpyDeviceToHost));
  }
  if (nv) {
    CK(c, cudaMemcpy(pos.data() + ns, c->d_vrec_pos, sizeof(float4) * (size_t)nv, cudaMemcpyDeviceToHost));
    CK(c, cudaMemcpy(pw.data() + ns, c->d_vrec_pow, sizeof(float4) * (size_t)nv, cudaMemcpyDeviceToHost));
  }
  std::vector<unsigned long long> key((size_t)n);   // (photon index, call ordinal)
  std::vector<int64_t> order((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    uint32_t meta, idx;
    memcpy(&meta, &pos[i].w, 4); memcpy(&idx, &pw[i].w, 4);
    key[i] = ((unsigned long long)idx << 4) | (meta & 15u);
    order[i] = i;
  }
  std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key[a] < key[b]; });
  for (int64_t j = 0; j < n; j++) {
    int64_t i = order[j];
    uint32_t meta; int32_t idx;
    memcpy(&meta, &pos[i].w, 4); memcpy(&idx, &pw[i].w, 4);
    int seq, kind, type, id;
    unpack_meta(meta, seq, kind, type, id);
    pm_record &r = out[j];
    r.type = type; r.id = id; r.index = idx; r.kind = kind;
    r.loc[0] = pos[i].x; r.loc[1] = pos[i].y; r.loc[2] = pos[i].z;
    r.dir[0] = dir[i].x; r.dir[1] = dir[i].y; r.dir[2] = dir[i].z;
    r.energy[0] = pw[i].x; r.energy[1] = pw[i].y; r.energy[2] = pw[i].z;
  }
  return PM_OK;
}

// ---- Mode B ----------------------------------------------------------------------------------------------
int pm_knn_build_points(pm_context *c, int which, const float *pos4, const float *pow4, int64_t n, bool records) {
  ARG(c, c && (which == PM_MAP_SURFACE || which == PM_MAP_VOLUME), "bad map id");
  ARG(c, n >= 0 && (n == 0 || pos4), "null points");
  CK(c, cudaSetDevice(c->device));
  int launches = 0;
  cudaError_t e;
  {
    SpanGuard g(c, K_KNN_BUILD);
    e = knn_build(c->knn[which], (const float4 *)pos4, (const float4 *)pow4, n, (records && which == PM_MAP_SURFACE) ? 1 : 0, c->knn_curve,
                  c->stream, &launches);
  }
  c->knn_from_records[which] = false; c->knn_stale[which] = false;
  c->launches += launches;
  CK(c, e);
  return PM_OK;
}
int pm_knn_build(pm_context *c, int which) {
  ARG(c, c && (which == PM_MAP_SURFACE || which == PM_MAP_VOLUME), "bad map id");
  CK(c, cudaSetDevice(c->device));
  float *pos = nullptr, *pw = nullptr; int64_t n = 0;
  int rc = pm_record_buffers(c, which, &pos, &pw, nullptr, &n);
  if (rc != PM_OK) return rc;
  int launches = 0;
  // surface map: wall hits only (sphere hits carry no energy in the reference either, PMK:1176)
  cudaError_t e;
  {
    SpanGuard g(c, K_KNN_BUILD);
    e = knn_build(c->knn[which], (const float4 *)pos, (const float4 *)pw, n, which == PM_MAP_SURFACE ? 1 : 0, c->knn_curve, c->stream, &launches);
  }
  c->knn_from_records[which] = true; c->knn_stale[which] = false;
  c->launches += launches;
  CK(c, e);
  return PM_OK;
}
int pm_knn_set_curve(pm_context *c, int curve) {
  ARG(c, c && (curve == PM_CURVE_MORTON || curve == PM_CURVE_HILBERT), "bad curve id");
  c->knn_curve = curve;
  return PM_OK;
}
int pm_knn_set_batched(pm_context *c, bool on) {
  ARG(c, c != nullptr, "null context");
  c->knn_batched = on;
  return PM_OK;
}
int pm_knn_size(pm_context *c, int which, int64_t *n, int32_t *levels) {
  ARG(c, c && (which == PM_MAP_SURFACE || which == PM_MAP_VOLUME), "bad map id");
  if (n) *n = c->knn[which].n;
  if (levels) *levels = c->knn[which].levels;
  return PM_OK;
}
static int knn_query_common(pm_context *c, int which, const float *q4, int64_t nq, int k, float max_r2, int32_t *idx, float *d2,
                            int32_t *cnt, float *rgb4) {
  ARG(c, c && (which == PM_MAP_SURFACE || which == PM_MAP_VOLUME), "bad map id");
  ARG(c, k >= 1 && k <= 128, "k must be in [1, 128]");
  ARG(c, nq >= 0 && (nq == 0 || q4), "null queries");
  ARG(c, !(max_r2 < 0.0f), "negative radius");
  if (knn_usable(c, which) != PM_OK) return PM_ERR_STATE;
  CK(c, cudaSetDevice(c->device));
  if (rgb4 && !c->knn[which].power && c->knn[which].n > 0) { c->err = "map was built without powers"; return PM_ERR_STATE; }
  {
    SpanGuard g(c, K_KNN_QUERY);
    CK(c, knn_query(c->knn[which], (const float4 *)q4, nq, k, max_r2, idx, d2, cnt, which == PM_MAP_VOLUME, (float4 *)rgb4, c->num_sms, c->stream));
  }
  if (nq > 0) c->launches++;
  return PM_OK;
}
int pm_knn_query(pm_context *c, int which, const float *q4, int64_t nq, int k, float max_r2, int32_t *idx, float *d2, int32_t *cnt) {
  ARG(c, c && idx && d2 && cnt, "null output");
  return knn_query_common(c, which, q4, nq, k, max_r2, idx, d2, cnt, nullptr);
}
int pm_knn_radiance(pm_context *c, int which, const float *q4, int64_t nq, int k, float max_r2, float *rgb4) {
  ARG(c, c && rgb4, "null output");
  return knn_query_common(c, which, q4, nq, k, max_r2, nullptr, nullptr, nullptr, rgb4);
}
int pm_render_knn(pm_context *c, float t, bool media, int width, int height, int y0, int y1, int k, float max_r2, float w_surface,
                  float w_volume, pm_uchar4 *dev_rgba, float *dev_rgbf) {
  return pm_render_knn_rows(c, t, media, width, height, y0, y1, 1, k, max_r2, w_surface, w_volume, dev_rgba, dev_rgbf);
}
int pm_render_knn_rows(pm_context *c, float t, bool media, int width, int height, int y0, int y1, int y_step, int k, float max_r2,
                       float w_surface, float w_volume, pm_uchar4 *dev_rgba, float *dev_rgbf) {
  ARG(c, c != nullptr, "null context");
  ARG(c, width > 0 && height > 0 && y0 >= 0 && y0 <= y1 && y1 <= height && y_step >= 1, "bad frame geometry");
  ARG(c, k >= 1 && k <= 128, "k must be in [1, 128]");
  if ((c->knn[0].n > 0 && !c->knn[0].power) || (media && c->knn[1].n > 0 && !c->knn[1].power)) { c->err = "maps were built without powers"; return PM_ERR_STATE; }
  if (knn_usable(c, PM_MAP_SURFACE) != PM_OK || (media && knn_usable(c, PM_MAP_VOLUME) != PM_OK)) return PM_ERR_STATE;
  CK(c, cudaSetDevice(c->device));
  c->dsc = make_device_scene(c->scene, t);
  {
    SpanGuard g(c, K_KNN_RENDER);
    if (!c->d_work) CK(c, cudaMalloc(&c->d_work, sizeof(unsigned long long)));
    CK(c, knn_render(c->dsc, c->knn[0], c->knn[1], k, max_r2, w_surface, w_volume, width, height, y0, y1, y_step, media, c->d_work, (uchar4 *)dev_rgba,
                     (float4 *)dev_rgbf, c->num_sms, c->stream, c->knn_batched));
  }
  if (y1 > y0) c->launches++;
  return PM_OK;
}
static int ensure_framebuffers(pm_context *c, int64_t pixels);
int pm_render_knn_host(pm_context *c, float t, bool media, int width, int height, int k, float max_r2, float w_surface, float w_volume,
                       pm_uchar4 *host_rgba, float *host_rgbf) {
  ARG(c, c != nullptr, "null context");
  ARG(c, width > 0 && height > 0, "bad frame geometry");
  CK(c, cudaSetDevice(c->device));
  int64_t pixels = (int64_t)width * height;
  int rc = ensure_framebuffers(c, pixels);
  if (rc != PM_OK) return rc;
  rc = pm_render_knn(c, t, media, width, height, 0, height, k, max_r2, w_surface, w_volume, host_rgba ? (pm_uchar4 *)c->d_fb_u8 : nullptr,
                     host_rgbf ? (float *)c->d_fb_f32 : nullptr);
  if (rc != PM_OK) return rc;
  if (host_rgba) CK(c, cudaMemcpyAsync(host_rgba, c->d_fb_u8, sizeof(uchar4) * (size_t)pixels, cudaMemcpyDeviceToHost, c->stream));
  if (host_rgbf) CK(c, cudaMemcpyAsync(host_rgbf, c->d_fb_f32, sizeof(float4) * (size_t)pixels, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return PM_OK;
}
int pm_knn_radiance_cone(pm_context *c, const float *q4, int64_t nq, int k, float sq_radius, float exposure, float *rgb4) {
  ARG(c, c && rgb4, "null output");
  ARG(c, k >= 1 && k <= 128, "k must be in [1, 128]");
  ARG(c, nq >= 0 && (nq == 0 || q4), "null queries");
  ARG(c, sq_radius >= 0.0f && exposure != 0.0f, "bad radius / exposure");
  const KnnMap &m = c->knn[PM_MAP_SURFACE];
  if (knn_usable(c, PM_MAP_SURFACE) != PM_OK) return PM_ERR_STATE;
  if (m.n > 0 && (m.src_pos != c->d_rec_pos || !c->d_rec_dir)) { c->err = "the cone filter needs the surface map built from the record buffers (pm_knn_build)"; return PM_ERR_STATE; }
  CK(c, cudaSetDevice(c->device));
  // inward wall normals: surfaceNormal(1, id, p, gOrigin) = normalize(e_axis * (0 - offset)), photonMappingKernel - Copy.cu:193
  float normals[PM_MAX_PLANES][3];
  memset(normals, 0, sizeof(normals));
  for (int i = 0; i < PM_MAX_PLANES; i++) {
    int axis = (int)c->scene.planes[i][0];
    float off = c->scene.planes[i][1];
    if (axis >= 0 && axis <= 2 && off != 0.0f) normals[i][axis] = off > 0.0f ? -1.0f : 1.0f;
  }
  {
    SpanGuard g(c, K_KNN_QUERY);
    CK(c, knn_query(m, (const float4 *)q4, nq, k, sq_radius, nullptr, nullptr, nullptr, 0, (float4 *)rgb4, c->num_sms, c->stream, c->d_rec_pos,
                    c->d_rec_dir, &normals[0][0], exposure));
  }
  if (nq > 0) c->launches++;
  return PM_OK;
}
int pm_knn_sorted_host(pm_context *c, int which, uint32_t *keys, uint32_t *perm, int64_t n) {
  ARG(c, c && (which == PM_MAP_SURFACE || which == PM_MAP_VOLUME), "bad map id");
  const KnnMap &m = c->knn[which];
  ARG(c, n >= 0 && n <= m.n_sorted_pad, "more entries than were sorted");
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));
  if (keys && n) CK(c, cudaMemcpy(keys, m.keys[m.sorted], sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
  if (perm && n) CK(c, cudaMemcpy(perm, m.vals[m.sorted], sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
  return PM_OK;
}
int pm_knn_level_host(pm_context *c, int which, int level, int64_t *count, float *boxes6) {
  ARG(c, c && (which == PM_MAP_SURFACE || which == PM_MAP_VOLUME), "bad map id");
  const KnnMap &m = c->knn[which];
  ARG(c, level >= 0 && level < m.levels, "no such level");
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));
  if (count) *count = m.cnt[level];
  if (boxes6)
    for (int a = 0; a < 6; a++)   // device layout: six floats per entity; the caller gets six arrays
      CK(c, cudaMemcpy2D(boxes6 + (size_t)a * m.cnt[level], sizeof(float), m.boxes + m.off[level] + a, 6 * sizeof(float), sizeof(float),
                         (size_t)m.cnt[level], cudaMemcpyDeviceToHost));
  return PM_OK;
}

int pm_render(pm_context *c, float t, bool interp, bool media, int width, int height, int y0, int y1,
              pm_uchar4 *dev_rgba, float *dev_rgbf) {
  ARG(c, c != nullptr, "null context");
  ARG(c, width > 0 && height > 0 && y0 >= 0 && y0 <= y1 && y1 <= height, "bad frame geometry");
  if (!c->tables_valid) { c->err = "pm_render before pm_build_map / pm_set_map_host"; return PM_ERR_STATE; }
  CK(c, cudaSetDevice(c->device));
  c->dsc = make_device_scene(c->scene, t);
  CK(c, cudaStreamWaitEvent(c->stream, c->ev_tbl_built, 0));   // the tables may have been built on another stream (pipelined frames)
  {
    SpanGuard g(c, K_RENDER);
    CK(c, launch_render(c->dsc, c->d_vol, c->d_surf, width, height, y0, y1, interp, media, (uchar4 *)dev_rgba, (float4 *)dev_rgbf, c->stream));
  }
  CK(c, cudaEventRecord(c->ev_tbl_read[c->tbl], c->stream));
  if (y1 > y0) c->launches++;
  return PM_OK;
}

static int ensure_framebuffers(pm_context *c, int64_t pixels) {
  if (pixels <= c->fb_pixels) return PM_OK;
  CK(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->d_fb_u8); cudaFree(c->d_fb_f32);
  c->d_fb_u8 = nullptr; c->d_fb_f32 = nullptr; c->fb_pixels = 0;
  CK(c, cudaMalloc(&c->d_fb_u8, sizeof(uchar4) * (size_t)pixels));
  CK(c, cudaMalloc(&c->d_fb_f32, sizeof(float4) * (size_t)pixels));
  c->fb_pixels = pixels;
  return PM_OK;
}

// the rows the *_host frame calls cover: the whole frame, or this rank's band (pm_set_row_band)
static void frame_rows(const pm_context *c, int height, int *y0, int *y1) {
  *y0 = 0; *y1 = height;
  if (c->band_y0 >= 0) { *y0 = std::min(c->band_y0, height); *y1 = std::min(c->band_y1, height); }
}

static int ensure_async_buffers(pm_context *c, int64_t pixels) {
  if (!c->copy_stream) {
    CK(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < pm_context::kFrameRing; k++) {
      CK(c, cudaEventCreateWithFlags(&c->ev_rendered[k], cudaEventDisableTiming));
      CK(c, cudaEventCreateWithFlags(&c->ev_copied[k], cudaEventDisableTiming));
    }
  }
  if (pixels > c->async_pixels) {
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaStreamSynchronize(c->copy_stream));
    for (int k = 0; k < pm_context::kFrameRing; k++) {
      cudaFree(c->d_fb_async[k]); c->d_fb_async[k] = nullptr;
    }
    c->async_pixels = 0;
    for (int k = 0; k < pm_context::kFrameRing; k++) CK(c, cudaMalloc(&c->d_fb_async[k], sizeof(uchar4) * (size_t)pixels));
    c->async_pixels = pixels;
  }
  return PM_OK;
}

// Allocate the context-owned frame buffers of the *_host calls now.  Allocation synchronises with the device, so ranks that
// share a device must not do it between a peer's exchange kernel and their own (pm_group_* reserves before the first frame).
int pm_reserve_frame(pm_context *c, int width, int height) {
  ARG(c, c != nullptr && width > 0 && height > 0, "bad frame geometry");
  CK(c, cudaSetDevice(c->device));
  int rc = ensure_framebuffers(c, (int64_t)width * height);
  if (rc != PM_OK) return rc;
  return ensure_async_buffers(c, (int64_t)width * height);
}

int pm_render_host(pm_context *c, float t, bool interp, bool media, int width, int height, pm_uchar4 *host_rgba, float *host_rgbf) {
  ARG(c, c != nullptr, "null context");
  ARG(c, width > 0 && height > 0, "bad frame geometry");
  CK(c, cudaSetDevice(c->device));
  int64_t pixels = (int64_t)width * height;
  int rc = ensure_framebuffers(c, pixels);
  if (rc != PM_OK) return rc;
  int y0, y1;
  frame_rows(c, height, &y0, &y1);
  rc = pm_render(c, t, interp, media, width, height, y0, y1, host_rgba ? (pm_uchar4 *)c->d_fb_u8 : nullptr, host_rgbf ? (float *)c->d_fb_f32 : nullptr);
  if (rc != PM_OK) return rc;
  // Wait for the frame BEFORE handing the copies to the runtime: a copy into pageable host memory blocks inside the driver until
  // the stream reaches it, and while it does other host threads cannot launch -- ranks sharing this process would then never
  // reach the exchange this stream may be waiting in (cudaStreamSynchronize has no such side effect).
  CK(c, cudaStreamSynchronize(c->stream));
  const size_t off = (size_t)y0 * width, cnt = (size_t)(y1 - y0) * width;
  if (host_rgba && cnt) CK(c, cudaMemcpyAsync(host_rgba + off, c->d_fb_u8 + off, sizeof(uchar4) * cnt, cudaMemcpyDeviceToHost, c->stream));
  if (host_rgbf && cnt) CK(c, cudaMemcpyAsync(host_rgbf + 4 * off, c->d_fb_f32 + off, sizeof(float4) * cnt, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return PM_OK;
}

int pm_frame_host(pm_context *c, float t, bool emit, bool interp, bool media, int width, int height, pm_uchar4 *host_rgba, float *host_rgbf) {
  ARG(c, c != nullptr, "null context");
  int rc;
  if (emit) {
    if ((rc = pm_clear_map(c)) != PM_OK) return rc;
    if ((rc = pm_trace(c, t, media ? PM_TRACE_MEDIA : 0u)) != PM_OK) return rc;
    if ((rc = pm_build_map(c)) != PM_OK) return rc;
  }
  return pm_render_host(c, t, interp, media, width, height, host_rgba, host_rgbf);
}

// One frame of a three-stage pipeline: (emit: clear + trace) on the context's stream, then exchange + map build on the auxiliary
// stream, then the render -- and the rank barrier when asked -- on the render stream, so the next frame's trace does not wait for them
// and the next frame's exchange does not wait for this frame's render (the accumulators rotate through three buffers for exactly
// that, pm_peer.cu, and the gather tables through two, pm_build_map).  At 8 GPUs the part after the trace is a chain of small,
// latency-bound kernels with two cross-rank waits in it, longer than the trace itself.
static int frame_stages(pm_context *c, float t, bool emit, bool interp, bool media, int width, int height, int y0, int y1,
                        pm_uchar4 *dev_rgba, float *dev_rgbf, cudaEvent_t wait_before_render, bool barrier_after) {
  if (!c->aux_stream) {
    int least = 0, greatest = 0;   // the small kernels of the later stages take whatever SMs the trace kernel leaves, before anything else
    CK(c, cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CK(c, cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, greatest));
    CK(c, cudaStreamCreateWithPriority(&c->render_stream, cudaStreamNonBlocking, greatest));
    CK(c, cudaEventCreateWithFlags(&c->ev_traced, cudaEventDisableTiming));
  }
  int rc;
  if (emit) {
    if ((rc = pm_clear_map(c)) != PM_OK) return rc;
    if ((rc = pm_trace(c, t, media ? PM_TRACE_MEDIA : 0u)) != PM_OK) return rc;
  }
  CK(c, cudaEventRecord(c->ev_traced, c->stream));
  CK(c, cudaStreamWaitEvent(c->aux_stream, c->ev_traced, 0));
  cudaStream_t main_stream = c->stream;
  c->stream = c->aux_stream;   // pm_build_map / pm_render / pm_peer_barrier launch on the context's current stream
  rc = emit ? pm_build_map(c) : PM_OK;
  // third stage: the render waits for this frame's tables (pm_render waits for ev_tbl_built) -- and, without an emit, for the trace
  // stream's earlier work -- while aux_stream is free for the next frame's exchange and map build
  c->stream = c->render_stream;
  if (rc == PM_OK && !emit) rc = cudaStreamWaitEvent(c->render_stream, c->ev_traced, 0) == cudaSuccess ? PM_OK : PM_ERR_CUDA;
  if (rc == PM_OK && wait_before_render) rc = cudaStreamWaitEvent(c->render_stream, wait_before_render, 0) == cudaSuccess ? PM_OK : PM_ERR_CUDA;
  if (rc == PM_OK) rc = pm_render(c, t, interp, media, width, height, y0, y1, dev_rgba, dev_rgbf);
  if (rc == PM_OK && barrier_after) rc = pm_peer_barrier(c);
  c->stream = main_stream;
  return rc;
}

int pm_frame_device(pm_context *c, float t, bool emit, bool interp, bool media, int width, int height, pm_uchar4 *dev_rgba, float *dev_rgbf) {
  ARG(c, c != nullptr, "null context");
  ARG(c, width > 0 && height > 0, "bad frame geometry");
  CK(c, cudaSetDevice(c->device));
  int y0, y1;
  frame_rows(c, height, &y0, &y1);
  return frame_stages(c, t, emit, interp, media, width, height, y0, y1, dev_rgba, dev_rgbf, nullptr, c->world > 1);
}

int pm_frame_host_async(pm_context *c, float t, bool emit, bool interp, bool media, int width, int height, pm_uchar4 *host_rgba,
                        int64_t *ticket) {
  ARG(c, c != nullptr && host_rgba != nullptr && ticket != nullptr, "null argument");
  ARG(c, width > 0 && height > 0, "bad frame geometry");
  CK(c, cudaSetDevice(c->device));
  const int64_t pixels = (int64_t)width * height;
  {
    int rc0 = ensure_async_buffers(c, pixels);
    if (rc0 != PM_OK) return rc0;
  }
  const int64_t tk = c->next_ticket;
  const int k = (int)(tk % pm_context::kFrameRing);
  int y0, y1;
  frame_rows(c, height, &y0, &y1);
  // the render waits until the copy three frames ago has drained this frame buffer
  int rc = frame_stages(c, t, emit, interp, media, width, height, y0, y1, (pm_uchar4 *)c->d_fb_async[k], nullptr, c->ev_copied[k], false);
  if (rc != PM_OK) return rc;
  CK(c, cudaEventRecord(c->ev_rendered[k], c->render_stream));
  CK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_rendered[k], 0));
  const size_t off = (size_t)y0 * width, cnt = (size_t)(y1 - y0) * width;
  if (cnt) CK(c, cudaMemcpyAsync(host_rgba + off, c->d_fb_async[k] + off, sizeof(uchar4) * cnt, cudaMemcpyDeviceToHost, c->copy_stream));
  CK(c, cudaEventRecord(c->ev_copied[k], c->copy_stream));
  c->next_ticket = tk + 1;
  *ticket = tk;
  return PM_OK;
}

int pm_frame_wait(pm_context *c, int64_t ticket) {
  ARG(c, c != nullptr, "null context");
  ARG(c, ticket >= 0 && ticket < c->next_ticket && ticket + pm_context::kFrameRing >= c->next_ticket, "ticket is not one of the three most recent frames");
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaEventSynchronize(c->ev_copied[ticket % pm_context::kFrameRing]));
  return PM_OK;
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU: peers.  Every rank's exchange block is mapped into every other rank -- directly when the
// ranks are contexts of one process (peer access), through CUDA IPC handles when they are processes -- and pm_build_map
// then sums the accumulators with pm_peer.cu's kernel instead of a library collective.
// ---------------------------------------------------------------------------------------------------
static int peer_prepare(pm_context *c, int rank, int world) {
  ARG(c, c != nullptr, "null context");
  ARG(c, world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "rank / world out of range (at most 16 ranks)");
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));
  int rc = pm_peer_disconnect(c);
  if (rc != PM_OK) return rc;
  if (!c->d_acc_sum) CK(c, cudaMalloc(&c->d_acc_sum, sizeof(long long) * kAccEntries));
  // Everything a frame launches is launched once NOW, while this context is still alone (world == 1).  With CUDA's lazy module
  // loading the first launch of a kernel (ours, and the runtime's memset kernels) loads its code, which waits for work already
  // running on the device -- and a peer that shares the device may by then be spinning inside peer_reduce_kernel, waiting for
  // exactly this rank (measured: the same-device group timed out on its first frame).
  {
    CK(c, preload_trace_kernels()); CK(c, preload_map_kernels()); CK(c, preload_render_kernels()); CK(c, preload_peer_kernels());
    const int64_t first = c->first, last = c->last;
    const uint32_t w = c->mwc_w, z = c->mwc_z;
    c->first = std::min<int64_t>(first, c->n_photons); c->last = std::min<int64_t>(c->first + 64, c->n_photons);
    uchar4 *tmp = nullptr;
    CK(c, cudaMalloc(&tmp, sizeof(uchar4) * 64));
    const int64_t launches = c->launches;
    const bool timing = c->timing; c->timing = false;
    rc = pm_clear_map(c);
    if (rc == PM_OK) rc = pm_trace(c, 0.0f, 0u);
    if (rc == PM_OK) rc = pm_trace(c, 0.0f, PM_TRACE_MEDIA);
    if (rc == PM_OK) rc = pm_build_map(c);
    if (rc == PM_OK) rc = pm_render(c, 0.0f, false, true, 16, 4, 0, 4, (pm_uchar4 *)tmp, nullptr);
    if (rc == PM_OK && launch_peer_reduce(c->pv, 0, 0u, c->d_acc_sum, 1, c->stream) != cudaSuccess) rc = PM_ERR_CUDA;
    if (rc == PM_OK && launch_peer_barrier(c->pv, 0u, c->stream) != cudaSuccess) rc = PM_ERR_CUDA;
    if (rc == PM_OK) rc = pm_clear_map(c);
    cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    c->first = first; c->last = last; c->mwc_w = w; c->mwc_z = z; c->launches = launches; c->timing = timing;
    if (rc != PM_OK) return rc;
  }
  CK(c, cudaMemsetAsync(c->d_xchg, 0, kExchangeBytes, c->stream));   // sequence numbers restart at 0 on every rank
  CK(c, cudaStreamSynchronize(c->stream));
  c->seq[0] = c->seq[1] = 0;
  return PM_OK;
}
static void peer_finish(pm_context *c, int rank, int world) {
  c->rank = rank; c->world = world;
  c->pv.world = world; c->pv.rank = rank;
  c->pv.hdr[rank] = (ExchangeHeader *)c->d_xchg;
  for (int p = 0; p < world; p++) c->pv.acc[p] = (const long long *)(c->pv.hdr[p] + 1);
  c->cur = 0; c->d_acc = (long long *)((ExchangeHeader *)c->d_xchg + 1);
  c->frame_no = 0; c->preclear_frame = -1;
  for (int b = 0; b < kAccBuffers; b++) c->precleared[b] = false;
  c->acc_summed = false;
}

int pm_peer_export(pm_context *c, void *handle) {
  ARG(c, c && handle, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == PM_PEER_HANDLE_BYTES, "handle size");
  CK(c, cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CK(c, cudaIpcGetMemHandle(&h, c->d_xchg));
  memcpy(handle, &h, sizeof(h));
  return PM_OK;
}

int pm_peer_connect(pm_context *c, int rank, int world, const void *handles) {
  int rc = peer_prepare(c, rank, world);
  if (rc != PM_OK) return rc;
  ARG(c, handles != nullptr || world == 1, "null handles");
  for (int p = 0; p < world; p++) {
    if (p == rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (size_t)p * PM_PEER_HANDLE_BYTES, sizeof(h));
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      c->err = std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(p) + "): " + cudaGetErrorString(e);
      cudaGetLastError();
      pm_peer_disconnect(c);
      return PM_ERR_CUDA;
    }
    c->ipc_opened[p] = ptr;
    c->pv.hdr[p] = (ExchangeHeader *)ptr;
  }
  peer_finish(c, rank, world);
  return PM_OK;
}

int pm_peer_connect_local(pm_context *c, int rank, int world, pm_context *const *members) {
  int rc = peer_prepare(c, rank, world);
  if (rc != PM_OK) return rc;
  ARG(c, members != nullptr && members[rank] == c, "members[rank] must be this context");
  c->peer_on_same_device = false;
  for (int p = 0; p < world; p++) {
    if (p == rank) continue;
    ARG(c, members[p] != nullptr && members[p] != c, "bad member list");
    if (members[p]->device != c->device) {
      int can = 0;
      CK(c, cudaDeviceCanAccessPeer(&can, c->device, members[p]->device));
      if (!can) { c->err = "device " + std::to_string(c->device) + " cannot access device " + std::to_string(members[p]->device) + " (no P2P path)"; return PM_ERR_STATE; }
      cudaError_t e = cudaDeviceEnablePeerAccess(members[p]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { c->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return PM_ERR_CUDA; }
      cudaGetLastError();
    } else {
      c->peer_on_same_device = true;
    }
    c->pv.hdr[p] = (ExchangeHeader *)members[p]->d_xchg;
  }
  peer_finish(c, rank, world);
  return PM_OK;
}

int pm_peer_disconnect(pm_context *c) {
  if (!c) return PM_ERR_ARG;
  for (int p = 0; p < kMaxPeers; p++) {
    if (c->ipc_opened[p]) { cudaIpcCloseMemHandle(c->ipc_opened[p]); c->ipc_opened[p] = nullptr; }
    c->pv.hdr[p] = nullptr; c->pv.acc[p] = nullptr;
  }
  c->world = 1; c->rank = 0; c->peer_on_same_device = false;
  c->pv.world = 1; c->pv.rank = 0; c->pv.hdr[0] = (ExchangeHeader *)c->d_xchg;
  c->cur = 0;
  if (c->d_xchg) { c->d_acc = (long long *)((ExchangeHeader *)c->d_xchg + 1); c->pv.acc[0] = c->d_acc; }
  c->acc_summed = false;
  return PM_OK;
}

int pm_peer_info(const pm_context *c, int *rank, int *world) {
  if (!c) return PM_ERR_ARG;
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return PM_OK;
}

int pm_peer_barrier(pm_context *c) {
  ARG(c, c != nullptr, "null context");
  if (c->world == 1) return PM_OK;
  CK(c, cudaSetDevice(c->device));
  CK(c, launch_peer_barrier(c->pv, ++c->seq[1], c->stream));
  c->launches++;
  return PM_OK;
}

int pm_peer_status(pm_context *c) {
  ARG(c, c != nullptr, "null context");
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));
  if (c->aux_stream) CK(c, cudaStreamSynchronize(c->aux_stream));
  if (c->render_stream) CK(c, cudaStreamSynchronize(c->render_stream));
  uint32_t e = 0;
  CK(c, cudaMemcpy(&e, &((ExchangeHeader *)c->d_xchg)->error, sizeof(e), cudaMemcpyDeviceToHost));
  if (e) { c->err = e == 1 ? "a peer never signalled its accumulators (wait timed out)" : "a peer never reached the barrier (wait timed out)"; return PM_ERR_STATE; }
  return PM_OK;
}

int pm_peer_set_timeout(pm_context *c, double seconds) {
  ARG(c, c != nullptr && seconds > 0.0, "bad timeout");
  c->pv.timeout_ns = (unsigned long long)(seconds * 1e9);
  return PM_OK;
}

// device memory another rank can map: the frame buffer rank 0 lets the other ranks render their bands into
int pm_shared_alloc(pm_context *c, size_t bytes, void **dev_ptr, void *handle) {
  ARG(c, c && dev_ptr && bytes > 0, "bad argument");
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaMalloc(dev_ptr, bytes));
  CK(c, cudaMemset(*dev_ptr, 0, bytes));
  if (handle) {
    cudaIpcMemHandle_t h;
    CK(c, cudaIpcGetMemHandle(&h, *dev_ptr));
    memcpy(handle, &h, sizeof(h));
  }
  return PM_OK;
}
int pm_shared_open(pm_context *c, const void *handle, void **dev_ptr) {
  ARG(c, c && handle && dev_ptr, "null argument");
  CK(c, cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  CK(c, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return PM_OK;
}
int pm_shared_close(pm_context *c, void *dev_ptr, bool opened) {
  ARG(c, c && dev_ptr, "null argument");
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));
  if (opened) CK(c, cudaIpcCloseMemHandle(dev_ptr)); else CK(c, cudaFree(dev_ptr));
  return PM_OK;
}

int pm_set_row_band(pm_context *c, int y0, int y1) {
  ARG(c, c != nullptr && ((y0 == -1 && y1 == -1) || (y0 >= 0 && y0 <= y1)), "bad row band");
  c->band_y0 = y0; c->band_y1 = y1;
  return PM_OK;
}

// ---------------------------------------------------------------------------------------------------
// legacy drop-in symbols (PMK:1523-1582): default context, synchronous, abort on error like checkCUDAError
// ---------------------------------------------------------------------------------------------------
static pm_context *g_default = nullptr;
static std::once_flag g_default_once;

pm_context *pm_default_context(void) {
  std::call_once(g_default_once, [] {
    pm_context *c = nullptr;
    int rc = pm_create(&c, -1);
    if (rc != PM_OK) {
      fprintf(stderr, "Cuda error: %s: %s.\n", "pmb200 default context", rc == PM_ERR_NO_DEVICE ? "no CUDA device (there is no CPU fallback)" : "initialisation failed");
      exit(EXIT_FAILURE);
    }
    const char *env = getenv("PMB200_NR_PHOTONS");
    if (env && *env) {
      long long n = atoll(env);
      if (pm_set_photon_count(c, n) != PM_OK) {
        fprintf(stderr, "Cuda error: %s: %s.\n", "PMB200_NR_PHOTONS", pm_last_error(c));
        exit(EXIT_FAILURE);
      }
    }
    g_default = c;
  });
  return g_default;
}

static void legacy_check(pm_context *c, int rc, const char *msg) {   // checkCUDAError, PMK:49-55
  if (rc == PM_OK) {
    cudaError_t e = cudaStreamSynchronize(c->stream);   // cudaThreadSynchronize() after every launch
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) return;
    c->err = cudaGetErrorString(e);
  }
  fprintf(stderr, "Cuda error: %s: %s.\n", msg, c->err.c_str());
  exit(EXIT_FAILURE);
}

void launch_init_random_numbers_kernel(void) {
  pm_context *c = pm_default_context();
  legacy_check(c, pm_init_random_table(c), "init_random_numbers_kernel failed!");
}

void launch_emit_photons_kernel(pm_uchar4 *pos, unsigned int image_width, unsigned int image_height, float animTime,
                                bool interpolateFlag, bool participatingMediaFlag) {
  (void)pos; (void)image_width; (void)image_height; (void)interpolateFlag;   // unused by the reference too
  pm_context *c = pm_default_context();
  legacy_check(c, pm_clear_map(c), "init_photons_kernel failed!");
  int rc = pm_trace(c, animTime, participatingMediaFlag ? PM_TRACE_MEDIA : 0u);
  if (rc == PM_OK) rc = pm_build_map(c);
  legacy_check(c, rc, "emit_photons_kernel failed!");
}

void launch_photon_mapping_kernel(pm_uchar4 *pos, unsigned int image_width, unsigned int image_height, float animTime,
                                  bool interpolateFlag, bool participatingMediaFlag) {
  pm_context *c = pm_default_context();
  if (!c->tables_valid) {   // the reference renders whatever the (zeroed) grid holds
    legacy_check(c, pm_build_map(c), "photon_mapping_kernel failed!");
  }
  legacy_check(c, pm_render(c, animTime, interpolateFlag, participatingMediaFlag, (int)image_width, (int)image_height, 0,
                            (int)image_height, pos, nullptr), "photon_mapping_kernel failed!");
}

void launch_render_kernel(pm_uchar4 *pos, unsigned int image_width, unsigned int image_height, float time, const float *pixelData) {
  (void)time;   // unused by the reference's render_kernel too
  pm_context *c = pm_default_context();
  const long long pixels = (long long)image_width * image_height;
  int rc = PM_OK;
  float *d_rgb = nullptr;
  cudaError_t e = cudaSetDevice(c->device);
  if (e == cudaSuccess && pixels > 0) e = cudaMalloc(&d_rgb, sizeof(float) * 3 * (size_t)pixels);   // cudaMalloc + cudaMemcpy + cudaFree per call, as kernelPBO.cu:306-312
  if (e == cudaSuccess && pixels > 0) e = cudaMemcpyAsync(d_rgb, pixelData, sizeof(float) * 3 * (size_t)pixels, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) { e = launch_present_float3(d_rgb, pixels, (uchar4 *)pos, c->stream); c->launches += pixels > 0; }
  if (e != cudaSuccess) { c->err = cudaGetErrorString(e); rc = PM_ERR_CUDA; }
  legacy_check(c, rc, "kernel failed!");
  cudaFree(d_rgb);
}

void launch_kernel(pm_uchar4 *pos, unsigned int image_width, unsigned int image_height, float time, void *numPhotons, void *photons) {
  (void)pos; (void)image_width; (void)image_height; (void)time; (void)numPhotons; (void)photons;   // the reference's body is commented out
  pm_context *c = pm_default_context();
  legacy_check(c, PM_OK, "kernel failed!");
}

}  // extern "C"
