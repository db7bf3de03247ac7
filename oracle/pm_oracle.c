/* oracle/pm_oracle.c -- TEST INFRASTRUCTURE (CPU restatement; never linked into or called by the product).
 *
 * A plain-C restatement of the reference's photon-mapping hot path, written from the algorithm, not
 * copied: every routine cites the reference lines (PMK = /root/reference/photonMappingKernel.cu) whose
 * arithmetic it reproduces.  It is pinned bit-for-bit against the reference's own routines compiled
 * as host C++ (oracle/_ref/libpmref_host.so, tests/test_oracle_vs_ref.py) and against the golden
 * vectors minted from them (tests/golden/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * Arithmetic contract (the SAME contract the CUDA kernels implement, so that photon records and
 * framebuffers can be compared bit-for-bit):
 *   - FP32 unless a sub-expression is double in the reference because of an unsuffixed literal; those
 *     are kept in double here ("f64" comments).  No FMA contraction (-ffp-contract=off).
 *   - normalize(v) = v * (1.0f / sqrtf(dot(v,v))), dot summed left to right (oracle/shim/cutil_math.h).
 *   - float3 / float = float3 * (1.0f / s).
 *   - rand3 components are drawn x, y, z in that order (the nvcc device order).
 *   - (unsigned char) of a negative / NaN value is 0 (device cvt.rzi.u32.f64 semantics).
 *   - sphere positions come from the host C library (cosf/sinf/sin as the reference's overloads resolve).
 */
#include "pm_oracle.h"

#include <math.h>
#include <string.h>

/* ---- tunables: PMK:9-46, :83-95 --------------------------------------------------------------- */
#define SEARCH_R        3        /* MAX_SEARCH_RADIUS */
#define SPLAT_R         3        /* MAX_SPLAT_RADIUS */
#define VOLUME_R        3        /* VOLUME_INTEGRATION_RADIUS */
#define N_CAUSTICS      100      /* CAUSTICS_PHOTONS */
#define N_SCATTER       3        /* MEDIUM_SCATTERING_ITERATIONS */
#define N_MARCH         10       /* MARCHING_ITERATIONS */
#define MAX_BOUNCES     5        /* nrBounces */
#define FAR_DIST        999999.9 /* literal in raytrace, PMK:229 (double -> float on assignment) */

typedef struct { float x, y, z; } v3;
typedef struct { int hit, type, idx; float dist; } hit_t;

static inline v3 V(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 subs(v3 a, float s) { return V(a.x - s, a.y - s, a.z - s); }
static inline v3 mul(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 divs(v3 a, float s) { float inv = 1.0f / s; return mul(a, inv); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 normalize(v3 v) { float inv = 1.0f / sqrtf(dot(v, v)); return mul(v, inv); }
static inline float comp(v3 a, int axis) { return axis == 0 ? a.x : (axis == 1 ? a.y : a.z); }

/* ---- RNG: get_random / randFloat, PMK:1029-1052 ------------------------------------------------- */
uint32_t pmo_mwc_next(uint32_t *w, uint32_t *z) {
  *z = 36969u * (*z & 65535u) + (*z >> 16);
  *w = 18000u * (*w & 65535u) + (*w >> 16);
  return (*z << 16) + *w;
}
float pmo_rand_float(uint32_t *w, uint32_t *z, float max) {
  uint32_t u = pmo_mwc_next(w, z);
  float rnd = (float)((int32_t)u) / (float)65535;
  rnd = rnd * 2 * max;
  rnd = rnd - max;
  return rnd;
}
/* init_random_numbers_kernel, PMK:1483-1498, rand3 in device argument order (x first) */
void pmo_mwc_table(uint32_t *w, uint32_t *z, float *xyz, int n) {
  for (int i = 0; i < n; i++) {
    xyz[3 * i + 0] = pmo_rand_float(w, z, 1.0f);
    xyz[3 * i + 1] = pmo_rand_float(w, z, 1.0f);
    xyz[3 * i + 2] = pmo_rand_float(w, z, 1.0f);
  }
}
/* n serial steps of the generator (get_random, PMK:1029-1037), results discarded: where the stream stands after the draws of the
 * photons before a given one (9 per photon in the medium walk, PMK:1258-1262) -- the reference's one-thread loop, no jump-ahead */
void pmo_mwc_skip(uint32_t *w, uint32_t *z, long n) {
  for (long i = 0; i < n; i++) (void)pmo_mwc_next(w, z);
}
/* Philox4x32-10 (Salmon et al., SC'11; Random123) -- the counter-based alternative to the reference's MWC table for
 * throughput runs: row i of the table = randFloat-style mapping of the first three words of
 * philox(counter = (i, 0, 0, 0), key = (seed_lo, seed_hi)).  Not part of the reference; defined here, mirrored in
 * csrc/pm_trace.cu, pinned by the Random123 known-answer vectors in tests/test_oracle_golden.py. */
void pmo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static float unit_from_u32(uint32_t u) {   /* the reference's randFloat(1.0) mapping of a raw draw, PMK:1039-1052 */
  float rnd = (float)((int32_t)u) / (float)65535;
  rnd = rnd * 2 * 1.0f;
  return rnd - 1.0f;
}
void pmo_philox_table(uint64_t seed, float *xyz, int n) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int i = 0; i < n; i++) {
    uint32_t ctr[4] = {(uint32_t)i, 0u, 0u, 0u}, o[4];
    pmo_philox4x32_10(ctr, key, o);
    xyz[3 * i] = unit_from_u32(o[0]); xyz[3 * i + 1] = unit_from_u32(o[1]); xyz[3 * i + 2] = unit_from_u32(o[2]);
  }
}

/* randomize, PMK:1201-1211: three draws, statement order x,y,z, each scaled by the input component */
static v3 randomize(uint32_t *w, uint32_t *z, v3 r) {
  v3 m;
  m.x = pmo_rand_float(w, z, r.x);
  m.y = pmo_rand_float(w, z, r.y);
  m.z = pmo_rand_float(w, z, r.z);
  return m;
}

/* ---- scene animation: positionObjects, PMK:1380-1404 ------------------------------------------------
 * cos(float)/sin(float) resolve to the float overloads, sin(double expr) to double ("f64"). */
void pmo_position_objects(pm_scene *sc, float t) {
  if (!sc->animate) return;
  sc->spheres[0][0] = (float)(1.0 * (double)cosf(t));
  sc->spheres[0][1] = (float)(0.5 * sin(2.0 * (double)t));
  sc->spheres[0][2] = (float)((double)sinf(t) + 3.5);
  sc->spheres[1][0] = 0.0f;
  sc->spheres[1][1] = (float)(sin(0.5 * (double)t + 5.0) - 0.25);
  sc->spheres[1][2] = 3.5f;
}

void pmo_scene_default(pm_scene *sc) {
  static const float sp[3][4] = {{1.0f, -1.0f, 1.0f, 0.4f}, {-0.6f, -1.0f, 4.5f, 0.4f}, {0.0f, 0.0f, 1.5f, 1.0f}};
  static const float pl[5][2] = {{0, 1.5f}, {1, -1.5f}, {0, -1.5f}, {1, 1.5f}, {2, 6.0f}};
  memset(sc, 0, sizeof(*sc));
  sc->n_spheres = 2; sc->n_planes = 5;
  memcpy(sc->spheres, sp, sizeof(sp)); memcpy(sc->planes, pl, sizeof(pl));
  sc->light[0] = 0.0f; sc->light[1] = 1.4f; sc->light[2] = 3.5f;
  sc->sz_img = 512; sc->cam_ox = 0.0f; sc->cam_oy = 0.0f; sc->animate = 1;
}

/* ---- intersection: checkDistance / raySphere / rayPlane / raytrace, PMK:106-165, :223-241 ----------- */
static inline void closer(float d, int type, int idx, hit_t *h) {
  if (d < h->dist && d > 0.0f) { h->type = type; h->idx = idx; h->dist = d; h->hit = 1; }
}
static void ray_sphere(const pm_scene *sc, int idx, v3 r, v3 o, hit_t *h) {
  v3 s = sub(V(sc->spheres[idx][0], sc->spheres[idx][1], sc->spheres[idx][2]), o);
  float radius = sc->spheres[idx][3];
  float A = dot(r, r);
  float B = (float)(-2.0 * (double)dot(s, r));               /* f64 product, exact */
  float C = dot(s, s) - radius * radius;                     /* pow(radius, 2.0f) */
  float D = B * B - 4 * A * C;
  if (D > 0.0f) {
    float sign = ((double)C < -0.00001) ? 1.0f : -1.0f;      /* f64 compare against the double literal */
    float d = (-B + sign * sqrtf(D)) / (2 * A);
    closer(d, 0, idx, h);
  }
}
static void ray_plane(const pm_scene *sc, int idx, v3 r, v3 o, hit_t *h) {
  int axis = (int)sc->planes[idx][0];
  float off = sc->planes[idx][1];
  if (axis < 0 || axis > 2) return;
  float rc = comp(r, axis), oc = comp(o, axis);
  if (rc != 0.0f) closer((off - oc) / rc, 1, idx, h);
}
/* raytrace with ignoreMedium == true (the only live use): distance reset, type/idx left STALE on a miss */
static void raytrace(const pm_scene *sc, v3 ray, v3 org, hit_t *h) {
  h->hit = 0;
  h->dist = (float)FAR_DIST;
  for (int i = 0; i < sc->n_spheres; i++) ray_sphere(sc, i, ray, org, h);
  for (int i = 0; i < sc->n_planes; i++) ray_plane(sc, i, ray, org, h);
}

/* ---- normals, reflection, refraction: PMK:181-209, :620-668 --------------------------------------- */
static v3 surface_normal(const pm_scene *sc, int type, int idx, v3 P, v3 inside) {
  if (type == 0) return normalize(sub(P, V(sc->spheres[idx][0], sc->spheres[idx][1], sc->spheres[idx][2])));
  int axis = (int)sc->planes[idx][0];
  float off = sc->planes[idx][1];
  v3 N = V(0.0f, 0.0f, 0.0f);
  if (axis == 0) N.x = inside.x - off; else if (axis == 1) N.y = inside.y - off; else if (axis == 2) N.z = inside.z - off;
  return normalize(N);   /* 0/0 -> NaN when `inside` lies exactly on the plane: hazard H1, reproduced on purpose */
}
static v3 reflect3(const pm_scene *sc, v3 ray, v3 from, int type, int idx, v3 P) {
  v3 N = mul(surface_normal(sc, type, idx, P, from), 1.0f);
  return normalize(sub(ray, mul(N, 2 * dot(ray, N))));
}
static v3 refract3(const pm_scene *sc, v3 ray, v3 from, int type, int idx, v3 P, float factor) {
  v3 normal = mul(surface_normal(sc, type, idx, P, from), factor);
  float n = 1.0f / 1.3f;
  if (factor == -1.0f) n = 1.0f;                                           /* leaves the glass unbent, PMK:636 */
  float cosI = -dot(normal, ray);
  float cosT2 = (float)(1.0 - ((double)(n * n) * (1.0 - (double)(cosI * cosI))));  /* f64 */
  if (cosT2 > 0.0f) return add(mul(ray, n), mul(normal, n * cosI - sqrtf(cosT2)));
  return V(0.0f, 0.0f, 0.0f);
}

/* The reference hand-unrolls handleReflection/handleRefraction{,2,3,4} (PMK:673-827): a mirror->glass->mirror
 * chain of at most four levels; the fourth glass passage has no mirror continuation.  Restated as a loop. */
static void follow_specular(const pm_scene *sc, v3 *ray, v3 from, hit_t *h, v3 *P, int start_with_mirror) {
  int mirror = start_with_mirror;
  for (int level = 1;; level++) {
    if (mirror) {
      *ray = reflect3(sc, *ray, from, h->type, h->idx, *P);
      raytrace(sc, *ray, *P, h);
      if (!h->hit) return;
      *P = add(mul(*ray, h->dist), *P);
      if (!(h->type == 0 && h->idx == 0)) return;
    }
    *ray = refract3(sc, *ray, *P, h->type, h->idx, *P, 1.0f);         /* into the glass */
    *P = add(mul(*ray, 0.00001f), *P);
    raytrace(sc, *ray, *P, h);
    *P = add(mul(*ray, h->dist), *P);                                 /* executed even on a miss */
    if (!(h->hit && h->type == 0 && h->idx == 0)) return;
    *ray = refract3(sc, *ray, *P, h->type, h->idx, *P, -1.0f);        /* out of the glass */
    *P = add(mul(*ray, 0.00001f), *P);
    raytrace(sc, *ray, *P, h);
    *P = add(mul(*ray, h->dist), *P);
    if (level == 4) return;
    if (!(h->type == 0 && h->idx == 1)) return;                       /* NOT gated on h->hit: stale ids, as PMK:807 */
    mirror = 1;
  }
}

/* ---- voxel addressing: getVoxelCoordinates PMK:260-267 (f64 throughout, C truncation) ----------------- */
void pmo_voxel(const float p[3], int v[3]) {
  v[0] = (int)((((double)p[0] + 3.0 / 2.0) / 3.0) * 32);
  v[1] = (int)((((double)p[1] + 3.0 / 2.0) / 3.0) * 32);
  v[2] = (int)(((double)p[2] / 6.0) * 32);
}
static inline int clampi(int v) { v = v < PM_GRID_N ? v : PM_GRID_N - 1; return v < 0 ? 0 : v; }
static inline float *vox(float *grid, int i, int j, int k) { return grid + 3 * ((i * PM_GRID_N + j) * PM_GRID_N + k); }
static inline const float *cvox(const float *grid, int i, int j, int k) { return grid + 3 * ((i * PM_GRID_N + j) * PM_GRID_N + k); }

/* window [v-R, v+R) clipped to [lo, hi) exactly as the reference's if-chains do (PMK:318-340, :836-858, :1076-1098) */
static inline void window(int v, int R, int lo, int hi, int *mn, int *mx) {
  *mn = lo; if (v - R >= lo) *mn = v - R;
  *mx = hi; if (v + R <= hi) *mx = v + R;
}

/* ---- photon deposition: storePhoton / splatEnergy / storeNeighborPhoton / storeVolumePhoton, PMK:1059-1183 --- */
typedef struct { pm_record *rec; long cap, count; } rec_sink;

static void record(rec_sink *rs, int type, int id, int index, int kind, v3 loc, v3 dir, v3 e) {
  if (rs && rs->rec && rs->count < rs->cap) {
    pm_record *r = &rs->rec[rs->count];
    r->type = type; r->id = id; r->index = index; r->kind = kind;
    r->loc[0] = loc.x; r->loc[1] = loc.y; r->loc[2] = loc.z;
    r->dir[0] = dir.x; r->dir[1] = dir.y; r->dir[2] = dir.z;
    r->energy[0] = e.x; r->energy[1] = e.y; r->energy[2] = e.z;
  }
  if (rs) rs->count++;
}
/* Every deposit is the reference's `photons[..] += e` in FP32 (PMK:1068, :1158, :1177).  Test aid: when a shadow grid is set
 * (pmo_set_shadow_grid64) the SAME FP32 deposit values are also summed in double there -- what the deposits add up to without the
 * rounding (and, at millions of photons, the saturation) of the reference's float voxels. */
static double *g_shadow64 = NULL;
static const float *g_shadow_base = NULL;
void pmo_set_shadow_grid64(double *grid64) { g_shadow64 = grid64; }
static inline void deposit(float *g, v3 e) {
  g[0] += e.x; g[1] += e.y; g[2] += e.z;
  if (g_shadow64 && g_shadow_base) {
    double *d = g_shadow64 + (g - g_shadow_base);
    d[0] += (double)e.x; d[1] += (double)e.y; d[2] += (double)e.z;
  }
}

static void splat_neighbor(float *grid, v3 energy, const int v[3], int i, int j, int k) {
  if (v[0] != i || v[1] != j || v[2] != k) {
    int dx = v[0] - i, dy = v[1] - j, dz = v[2] - k;
    float dist = sqrtf((float)(dx * dx + dy * dy + dz * dz));      /* "sqDistance" is a plain distance, PMK:281 */
    deposit(vox(grid, i, j, k), divs(mul(energy, 0.05f), dist));   /* SPLAT_ENERGY_WEIGHT*energy/dist */
  }
}
static void store_photon(float *grid, rec_sink *rs, int type, int id, v3 loc, v3 dir, v3 energy, int index) {
  record(rs, type, id, index, 0, loc, dir, energy);
  if (!grid) return;
  float p[3] = {loc.x, loc.y, loc.z};
  int v[3];
  pmo_voxel(p, v);
  v[0] = clampi(v[0]); v[1] = clampi(v[1]); v[2] = clampi(v[2]);
  if (type != 0) deposit(vox(grid, v[0], v[1], v[2]), energy);
  if (type != 1) return;                                            /* splatEnergy: planes only */
  int mnx, mxx, mny, mxy, mnz, mxz;
  window(v[0], SPLAT_R, 0, PM_GRID_N, &mnx, &mxx);
  window(v[1], SPLAT_R, 0, PM_GRID_N, &mny, &mxy);
  window(v[2], SPLAT_R, 0, PM_GRID_N, &mnz, &mxz);
  if (id == 0 || id == 2) {                                         /* x walls: slab i = 31 / 0 */
    int i = (id == 0) ? PM_GRID_N - 1 : 0;
    for (int j = mny; j < mxy; j++) for (int k = mnz; k < mxz; k++) splat_neighbor(grid, energy, v, i, j, k);
  } else if (id == 1 || id == 3) {                                  /* y walls: slab j = 0 / 31 */
    int j = (id == 1) ? 0 : PM_GRID_N - 1;
    for (int i = mnx; i < mxx; i++) for (int k = mnz; k < mxz; k++) splat_neighbor(grid, energy, v, i, j, k);
  } else if (id == 4) {                                             /* back wall: slab k = 31 */
    int k = PM_GRID_N - 1;
    for (int i = mnx; i < mxx; i++) for (int j = mny; j < mxy; j++) splat_neighbor(grid, energy, v, i, j, k);
  }
}
static void store_volume_photon(float *grid, rec_sink *rs, v3 loc, v3 energy, int index) {
  record(rs, -1, -1, index, 1, loc, V(0.0f, 0.0f, 0.0f), energy);
  if (!grid) return;
  float p[3] = {loc.x, loc.y, loc.z};
  int v[3];
  pmo_voxel(p, v);
  deposit(vox(grid, clampi(v[0]), clampi(v[1]), clampi(v[2])), energy);
}

/* getColor / filterColor, PMK:605-617 */
static v3 get_color(v3 in, int type, int idx) {
  v3 m = V(1.0f, 1.0f, 1.0f);
  if (type == 1 && idx == 0) m = V(0.0f, 1.0f, 0.0f);
  else if (type == 1 && idx == 2) m = V(1.0f, 0.0f, 0.0f);
  return V(fminf(m.x, in.x), fminf(m.y, in.y), fminf(m.z, in.z));
}

/* shadowPhoton, PMK:1185-1196: restores point/type/idx, NOT dist/hit */
static void shadow_photon(const pm_scene *sc, float *grid, rec_sink *rs, v3 ray, hit_t *h, v3 P, int index) {
  int t_type = h->type, t_idx = h->idx;
  v3 bumped = add(P, mul(ray, 0.00001f));
  raytrace(sc, ray, bumped, h);
  v3 sp = add(mul(ray, h->dist), bumped);
  store_photon(grid, rs, h->type, h->idx, sp, ray, V(-0.25f, -0.25f, -0.25f), index);
  h->type = t_type; h->idx = t_idx;
}

/* ---- stage 1: emitPhotons, PMK:1215-1375 --------------------------------------------------------------- */
static void emit_one(const pm_scene *sc, const float *table, int index, int media, uint32_t *w, uint32_t *z,
                     float *grid, rec_sink *rs) {
  int bounces = 1;
  v3 rgb = V(10.0f, 10.0f, 10.0f);
  v3 light = V(sc->light[0], sc->light[1], sc->light[2]);
  v3 ray = normalize(V(table[3 * index], table[3 * index + 1], table[3 * index + 2]));
  v3 original = ray, prev = light, P = V(0.0f, 0.0f, 0.0f);
  hit_t h; h.hit = 0; h.type = 0; h.idx = 0; h.dist = -1.0f;

  if (media) {   /* three fixed unit steps, random re-direction, deposit 5e-5*rgb; does not affect the surface walk */
    for (int i = 0; i < N_SCATTER; i++) {
      rgb = subs(rgb, 1.0f);
      P = add(mul(ray, 1.0f), prev);
      store_volume_photon(grid, rs, P, mul(rgb, 0.00005f), index);
      ray = normalize(randomize(w, z, V(table[3 * i], table[3 * i + 1], table[3 * i + 2])));   /* table[i], i=0..2 (sic) */
      prev = P;
    }
    ray = original; prev = light;
  }
  if (index < N_CAUSTICS) {   /* aimed at the glass sphere, jittered, NOT re-normalised */
    ray = normalize(sub(V(sc->spheres[0][0], sc->spheres[0][1], sc->spheres[0][2]), light));
    ray = add(ray, mul(normalize(V(table[3 * index], table[3 * index + 1], table[3 * index + 2])), 0.01f));
  }
  raytrace(sc, ray, prev, &h);

  int caustics = 0, new_point = 1;
  while (h.hit && bounces <= MAX_BOUNCES) {
    if (new_point) P = add(mul(ray, h.dist), prev);
    if (caustics) {
      rgb = mul(V(1.0f, 1.0f, 1.0f), 10.0f);
      store_photon(grid, rs, h.type, h.idx, P, ray, rgb, index);
    } else {
      rgb = mul(divs(mul(get_color(rgb, h.type, h.idx), 1.0f), sqrtf((float)bounces)), 5.0f);
      store_photon(grid, rs, h.type, h.idx, P, ray, rgb, index);
      shadow_photon(sc, grid, rs, ray, &h, P, index);
    }
    prev = P;
    if (h.type == 0 && h.idx == 1) {          /* mirror sphere */
      follow_specular(sc, &ray, prev, &h, &P, 1);
      caustics = 0; new_point = 0;
    } else if (h.type == 0 && h.idx == 0) {   /* glass sphere */
      follow_specular(sc, &ray, prev, &h, &P, 0);
      caustics = 1; new_point = 0;
    } else {                                   /* diffuse wall: `prev` already equals the hit point -> hazard H1 */
      ray = reflect3(sc, ray, prev, h.type, h.idx, P);
      raytrace(sc, ray, P, &h);
      caustics = 0; new_point = 1;
    }
    bounces++;
  }
}

long pmo_emit(const pm_scene *scene, float t, const float *table, int n0, int n1, int media,
              uint32_t *w, uint32_t *z, float *grid, pm_record *rec, long max_rec) {
  pm_scene sc = *scene;
  pmo_position_objects(&sc, t);
  rec_sink rs; rs.rec = rec; rs.cap = max_rec; rs.count = 0;
  g_shadow_base = grid;
  for (int i = n0; i < n1; i++) emit_one(&sc, table, i, media, w, z, grid, &rs);
  return rs.count;
}

/* ---- stage 4 (reference flavour): integrate / computeEnergy / interpolateEnergy, PMK:286-600 ------------ */
static v3 integrate(const float *grid, v3 e, const int wp[3], int type, int id) {
  int mnx, mxx, mny, mxy, mnz, mxz;
  window(wp[0], SEARCH_R, 0, PM_GRID_N, &mnx, &mxx);
  window(wp[1], SEARCH_R, 0, PM_GRID_N, &mny, &mxy);
  window(wp[2], SEARCH_R, 0, PM_GRID_N, &mnz, &mxz);
  if (type != 1) return e;
#define ACC(i, j, k) do { const float *g = cvox(grid, i, j, k); e = add(e, mul(V(g[0], g[1], g[2]), 0.0005f)); } while (0)
  if (id == 0 || id == 2) {
    int i = (id == 0) ? PM_GRID_N - 1 : 0;
    for (int j = mny; j < mxy; j++) for (int k = mnz; k < mxz; k++) ACC(i, j, k);
  } else if (id == 1 || id == 3) {
    int j = (id == 1) ? 0 : PM_GRID_N - 1;
    for (int i = mnx; i < mxx; i++) for (int k = mnz; k < mxz; k++) ACC(i, j, k);
  } else if (id == 4) {
    int k = PM_GRID_N - 1;
    for (int i = mnx; i < mxx; i++) for (int j = mny; j < mxy; j++) ACC(i, j, k);
  }
#undef ACC
  return e;
}
static v3 integrate_at(const float *grid, v3 p, int type, int id) {
  float pp[3] = {p.x, p.y, p.z}; int wp[3];
  pmo_voxel(pp, wp);
  return integrate(grid, V(0.0f, 0.0f, 0.0f), wp, type, id);
}
/* centerPoint PMK:392-400 via getWorldCoordinates PMK:269-274 (f64 scale and shift) */
static v3 center_point(v3 p) {
  float pp[3] = {p.x, p.y, p.z}; int wp[3];
  pmo_voxel(pp, wp);
  v3 c;
  c.x = (float)((double)((float)wp[0] / 32.0f) * 3.0 - 3.0 / 2.0);
  c.y = (float)((double)((float)wp[1] / 32.0f) * 3.0 - 3.0 / 2.0);
  c.z = (float)((double)((float)wp[2] / 32.0f) * 6.0);
  return add(c, V((float)(3.0f / (32 * 2.0)), (float)(3.0f / (32 * 2.0)), (float)(6.0f / (32 * 2.0))));
}
static inline float alpha_of(v3 p, v3 a, v3 b, int axis) {
  return (comp(p, axis) - comp(a, axis)) / (comp(b, axis) - comp(a, axis));
}
static inline v3 lerp3(v3 a, v3 b, float alfa) {   /* (1.0 - alfa) is f64, then narrowed to float: PMK:424-426 */
  return add(mul(a, (float)(1.0 - (double)alfa)), mul(b, alfa));
}
static inline void set_comp(v3 *a, int axis, float v) { if (axis == 0) a->x = v; else if (axis == 1) a->y = v; else a->z = v; }

static v3 interpolate_energy(const float *grid, v3 p, int type, int id) {
  if (type != 1) return V(0.0f, 0.0f, 0.0f);
  int a1, a2;                 /* firstAxis, secondAxis of bilinearInterpolate */
  if (id == 0 || id == 2) { a1 = 2; a2 = 1; }
  else if (id == 1 || id == 3) { a1 = 0; a2 = 2; }
  else if (id == 4) { a1 = 0; a2 = 1; }
  else return V(0.0f, 0.0f, 0.0f);
  static const float dim[3] = {3.0f / 32.0f, 3.0f / 32.0f, 6.0f / 32.0f};   /* getVoxelDim */
  v3 p1 = center_point(p), p2 = V(0, 0, 0), p3 = V(0, 0, 0), p4 = V(0, 0, 0);
  float s1 = comp(p, a1) > comp(p1, a1) ? dim[a1] : -dim[a1];
  set_comp(&p2, a1, comp(p1, a1) + s1); set_comp(&p3, a1, comp(p1, a1) + s1); set_comp(&p4, a1, comp(p1, a1));
  float s2 = comp(p, a2) > comp(p1, a2) ? dim[a2] : -dim[a2];
  set_comp(&p2, a2, comp(p1, a2)); set_comp(&p3, a2, comp(p1, a2) + s2); set_comp(&p4, a2, comp(p1, a2) + s2);
  v3 c1 = integrate_at(grid, p1, type, id), c2 = integrate_at(grid, p2, type, id);
  v3 c3 = integrate_at(grid, p3, type, id), c4 = integrate_at(grid, p4, type, id);
  float alfa = alpha_of(p, p1, p2, a1);
  v3 p12 = lerp3(p1, p2, alfa), c12 = lerp3(c1, c2, alfa);
  alfa = alpha_of(p, p3, p4, a1);
  v3 p34 = lerp3(p3, p4, alfa), c34 = lerp3(c3, c4, alfa);
  alfa = alpha_of(p, p12, p34, a2);
  return lerp3(c12, c34, alfa);
}
static v3 gather(const float *grid, v3 p, int type, int id, int interp) {
  return interp ? interpolate_energy(grid, p, type, id) : integrate_at(grid, p, type, id);
}

/* ---- stage 5 (reference flavour): integrateVolumePhotons, PMK:831-870 ----------------------------------- */
static v3 integrate_volume(const float *grid, v3 p) {
  float pp[3] = {p.x, p.y, p.z}; int wp[3];
  pmo_voxel(pp, wp);
  int mnx, mxx, mny, mxy, mnz, mxz;
  window(wp[0], VOLUME_R, 1, PM_GRID_N - 1, &mnx, &mxx);
  window(wp[1], VOLUME_R, 1, PM_GRID_N - 1, &mny, &mxy);
  window(wp[2], VOLUME_R, 1, PM_GRID_N - 1, &mnz, &mxz);
  v3 rgb = V(0.0f, 0.0f, 0.0f);
  for (int i = mnx; i < mxx; i++) for (int j = mny; j < mxy; j++) for (int k = mnz; k < mxz; k++) {
    const float *g = cvox(grid, i, j, k);
    rgb = add(rgb, V(g[0], g[1], g[2]));
  }
  return rgb;
}

/* ---- stages 3+4+5: computePixelColor PMK:926-1017 and the quantisation of photon_mapping_kernel PMK:1451-1459 -- */
static v3 pixel_color(const pm_scene *sc, const float *grid, float x, float y, int interp, int media) {
  v3 rgb = V(0.0f, 0.0f, 0.0f), origin = V(0.0f, 0.0f, 0.0f);
  v3 ray = V((float)((double)(x / (float)sc->sz_img) - 0.5), (float)(-((double)(y / (float)sc->sz_img) - 0.5)), 1.0f);
  hit_t h; h.hit = 0; h.type = 0; h.idx = 0; h.dist = -1.0f;
  if (media) {   /* 10 steps of 0.6 along the UNNORMALISED eye ray, raw box sums */
    v3 prev = origin;
    h.dist = 0.6f;
    for (int i = 0; i < N_MARCH; i++) { prev = add(mul(ray, h.dist), prev); rgb = add(rgb, integrate_volume(grid, prev)); }
  }
  raytrace(sc, ray, origin, &h);
  if (h.hit) {
    v3 P = mul(ray, h.dist);
    if (h.type == 0 && h.idx == 1) follow_specular(sc, &ray, origin, &h, &P, 1);
    else if (h.type == 0 && h.idx == 0) follow_specular(sc, &ray, origin, &h, &P, 0);
    if (h.hit) {
      v3 c = gather(grid, P, h.type, h.idx, interp);
      rgb = media ? add(rgb, mul(c, 0.15f)) : add(rgb, c);
    }
  }
  return rgb;
}
static inline uint8_t quantise(float v) {   /* device semantics: NaN and negatives -> 0 */
  double d = (double)v * 255.0;
  d = d > 255.0 ? 255.0 : d;
  return d > 0.0 ? (uint8_t)d : 0;
}
void pmo_render(const pm_scene *scene, float t, const float *grid, int width, int height, int y0, int y1,
                int interp, int media, float *rgb, uint8_t *rgba) {
  pm_scene sc = *scene;
  pmo_position_objects(&sc, t);
  for (int y = y0; y < y1; y++)
    for (int x = 0; x < width; x++) {
      v3 c = pixel_color(&sc, grid, (float)x + sc.cam_ox, (float)y + sc.cam_oy, interp, media);
      size_t i = (size_t)y * width + x;
      if (rgb) { rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z; }
      if (rgba) { rgba[4 * i] = quantise(c.x); rgba[4 * i + 1] = quantise(c.y); rgba[4 * i + 2] = quantise(c.z); rgba[4 * i + 3] = 0; }
    }
}

/* Eye-ray geometry for the Mode B renderer's oracle: per pixel the final hit after the mirror/glass chain (hit flag,
 * type, id, point) and the ten ray-march sample points of PMK:937-965.  out_hit: 8 floats per pixel
 * (hit, type, id, px, py, pz, 0, 0); out_march: 30 floats per pixel. */
void pmo_eye_geometry(const pm_scene *scene, float t, int width, int height, int y0, int y1, float *out_hit, float *out_march) {
  pm_scene sc = *scene;
  pmo_position_objects(&sc, t);
  for (int y = y0; y < y1; y++)
    for (int x = 0; x < width; x++) {
      float fx = (float)x + sc.cam_ox, fy = (float)y + sc.cam_oy;
      v3 origin = V(0.0f, 0.0f, 0.0f), P = V(0.0f, 0.0f, 0.0f);
      v3 ray = V((float)((double)(fx / (float)sc.sz_img) - 0.5), (float)(-((double)(fy / (float)sc.sz_img) - 0.5)), 1.0f);
      size_t i = (size_t)y * width + x;
      v3 prev = origin;
      for (int k = 0; k < N_MARCH; k++) {
        prev = add(mul(ray, 0.6f), prev);
        if (out_march) { out_march[30 * i + 3 * k] = prev.x; out_march[30 * i + 3 * k + 1] = prev.y; out_march[30 * i + 3 * k + 2] = prev.z; }
      }
      hit_t h; h.hit = 0; h.type = 0; h.idx = 0; h.dist = -1.0f;
      raytrace(&sc, ray, origin, &h);
      if (h.hit) {
        P = mul(ray, h.dist);
        if (h.type == 0 && h.idx == 1) follow_specular(&sc, &ray, origin, &h, &P, 1);
        else if (h.type == 0 && h.idx == 0) follow_specular(&sc, &ray, origin, &h, &P, 0);
      }
      float *o = out_hit + 8 * i;
      o[0] = (float)h.hit; o[1] = (float)h.type; o[2] = (float)h.idx; o[3] = P.x; o[4] = P.y; o[5] = P.z; o[6] = 0.0f; o[7] = 0.0f;
    }
}

/* ---- single-routine probes for unit-level parity ------------------------------------------------------- */
int pmo_raytrace(const pm_scene *sc, const float ray[3], const float org[3], float *dist, int *type, int *idx) {
  hit_t h; h.hit = 0; h.type = -1; h.idx = -1; h.dist = -1.0f;
  raytrace(sc, V(ray[0], ray[1], ray[2]), V(org[0], org[1], org[2]), &h);
  *dist = h.dist; *type = h.type; *idx = h.idx;
  return h.hit;
}
void pmo_integrate_volume(const float *grid, const float p[3], float rgb[3]) {
  v3 c = integrate_volume(grid, V(p[0], p[1], p[2])); rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}
void pmo_gather(const float *grid, const float p[3], int type, int id, int interp, float rgb[3]) {
  v3 c = gather(grid, V(p[0], p[1], p[2]), type, id, interp); rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}
