/* oracle/knn_oracle.c -- TEST INFRASTRUCTURE: CPU oracle for the Mode B photon-map pipeline (Morton keys, stable
 * key sort, brute-force k-nearest-photon search, radiance estimate).
 *
 * The reference has NO counterpart for these stages (its photon map is a dense voxel grid), so there is
 * nothing of the reference's to pin them against: PARITY UNPINNED BY THE REFERENCE.  The oracle is instead the
 * definition itself, written the slow obvious way (last row):
 *   - Morton key: 10 bits per axis over the map's world box (PMK:21-23), bit-interleaved x (lowest), y, z;
 *   - sort oracle: stable sort on the key;
 *   - k-NN: over ALL photons, d2 = (dx*dx + dy*dy) + dz*dz in FP32 without contraction, order by (d2, index)
 *     ascending, keep those with d2 <= max_r2, take k.  Index sets must match bit-exactly.
 *   - radiance estimate: sum of the k photon powers divided by pi*r_k^2 (surface) or 4/3*pi*r_k^3 (volume), r_k the
 *     distance of the k-th photon found; summed in (d2, index) order in double.
 * The only point-based estimator the reference's author ever wrote is the fixed-radius cone filter of the legacy file
 * (photonMappingKernel - Copy.cu:191-208); it is kept as an optional weight below for relating Mode B back to it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* expand 10 bits to every third bit */
static inline uint32_t spread10(uint32_t v) {
  v &= 1023u;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
static inline uint32_t quant10(float p, float lo, float inv_extent) {
  float f = (p - lo) * inv_extent * 1024.0f;
  if (!(f > 0.0f)) return 0u;        /* negatives and NaN -> cell 0 */
  if (f >= 1023.0f) return 1023u;
  return (uint32_t)f;
}
/* Sort keys.  Cell = 10 bits per axis over a box, in three tiers (tier number in key bits 30-31):
 *   tier 0  the photon map's world box (x,y in [-1.5,1.5], z in [0,6], PMK:21-23);
 *   tier 1  photons outside it but inside the 4x larger concentric box (x,y in [-6,6], z in [-9,15]): the reference's
 *           shadow photons land on the infinite wall planes outside the box (PMK:1185-1196) and its volume photons up to
 *           three units from the light, i.e. above the ceiling;
 *   tier 2  everything farther, over a coarse 192-unit box.
 * Clamping outside photons into the border cells of tier 0 (first version) made them share leaves with the in-box photons
 * next to the border, giving those leaves boxes that reach far outside the scene; keying ALL of them over the coarse box
 * (second version) put the millions of volume photons above the ceiling into a handful of cells, i.e. into leaves of
 * random points whose boxes all overlap: queries there scanned the whole population (58 ms for one query). */
static inline int inside_axis(float p, float lo, float inv_extent) {
  float f = (p - lo) * inv_extent * 1024.0f;
  return f >= -0.5f && f <= 1024.5f;
}
static inline void cell_of(const float p[3], uint32_t X[3], uint32_t *flag) {
  if (inside_axis(p[0], -1.5f, 1.0f / 3.0f) && inside_axis(p[1], -1.5f, 1.0f / 3.0f) && inside_axis(p[2], 0.0f, 1.0f / 6.0f)) {
    X[0] = quant10(p[0], -1.5f, 1.0f / 3.0f); X[1] = quant10(p[1], -1.5f, 1.0f / 3.0f); X[2] = quant10(p[2], 0.0f, 1.0f / 6.0f);
    *flag = 0u;
  } else if (inside_axis(p[0], -6.0f, 1.0f / 12.0f) && inside_axis(p[1], -6.0f, 1.0f / 12.0f) && inside_axis(p[2], -9.0f, 1.0f / 24.0f)) {
    X[0] = quant10(p[0], -6.0f, 1.0f / 12.0f); X[1] = quant10(p[1], -6.0f, 1.0f / 12.0f); X[2] = quant10(p[2], -9.0f, 1.0f / 24.0f);
    *flag = 1u << 30;
  } else {
    X[0] = quant10(p[0], -96.0f, 1.0f / 192.0f); X[1] = quant10(p[1], -96.0f, 1.0f / 192.0f); X[2] = quant10(p[2], -93.0f, 1.0f / 192.0f);
    *flag = 2u << 30;
  }
}
uint32_t pmo_morton30(const float p[3]) {
  uint32_t X[3], flag;
  cell_of(p, X, &flag);
  return flag | spread10(X[0]) | (spread10(X[1]) << 1) | (spread10(X[2]) << 2);
}
/* The same cell along the 3-D Hilbert curve (Skilling's transpose algorithm, AIP Conf. Proc. 707, 2004): the sort key the
 * product uses by default.  Runs of consecutive points of a Z-curve straddle octant boundaries, so the bounding boxes of
 * fixed-size runs (the tree's leaves and nodes) are loose; Hilbert runs are always connected. */
uint32_t pmo_hilbert30(const float p[3]) {
  uint32_t X[3], flag;
  cell_of(p, X, &flag);
  const uint32_t M = 1u << 9;
  for (uint32_t Q = M; Q > 1; Q >>= 1) {
    uint32_t P = Q - 1;
    for (int i = 0; i < 3; i++) {
      if (X[i] & Q) X[0] ^= P;
      else { uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
    }
  }
  X[1] ^= X[0]; X[2] ^= X[1];
  uint32_t t = 0;
  for (uint32_t Q = M; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
  X[0] ^= t; X[1] ^= t; X[2] ^= t;
  return flag | (spread10(X[0]) << 2) | (spread10(X[1]) << 1) | spread10(X[2]);   /* X[0] holds the most significant bit of each triple */
}
void pmo_morton30_many(const float *pos4, long n, uint32_t *keys) {
  for (long i = 0; i < n; i++) keys[i] = pmo_morton30(pos4 + 4 * i);
}
void pmo_hilbert30_many(const float *pos4, long n, uint32_t *keys) {
  for (long i = 0; i < n; i++) keys[i] = pmo_hilbert30(pos4 + 4 * i);
}

typedef struct { uint32_t key; uint32_t idx; } kv_t;
static int kv_cmp(const void *a, const void *b) {
  const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);   /* index as tie-break == stable */
}
/* stable sort of (key, index 0..n-1) pairs: perm[i] = original index of the i-th smallest */
void pmo_stable_sort_perm(const uint32_t *keys, long n, uint32_t *perm) {
  kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)n);
  for (long i = 0; i < n; i++) { kv[i].key = keys[i]; kv[i].idx = (uint32_t)i; }
  qsort(kv, (size_t)n, sizeof(kv_t), kv_cmp);
  for (long i = 0; i < n; i++) perm[i] = kv[i].idx;
  free(kv);
}

static inline float dist2(const float *a, const float *b) {
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return (dx * dx + dy * dy) + dz * dz;
}
static inline uint64_t knn_key(float d2, uint32_t idx) {
  uint32_t bits;
  memcpy(&bits, &d2, 4);
  return ((uint64_t)bits << 32) | idx;   /* d2 >= 0: its bit pattern orders like its value */
}

/* Brute force.  pos4: n photons as float4 (xyz + ignored w); queries: nq float4.  For each query writes up to k indices
 * (ascending (d2, index)) to out_idx[q*k..], their d2 to out_d2, and the number found to out_cnt[q]; unused slots get
 * index -1 / d2 = +inf.  A photon qualifies if d2 <= max_r2 (pass INFINITY for pure k-NN); NaN distances never qualify. */
void pmo_knn_bruteforce(const float *pos4, long n, const float *queries4, long nq, int k, float max_r2,
                        int32_t *out_idx, float *out_d2, int32_t *out_cnt) {
#pragma omp parallel
  {
    uint64_t *best = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(dynamic, 16)
    for (long q = 0; q < nq; q++) {
      int cnt = 0;
      const float *qp = queries4 + 4 * q;
      for (long i = 0; i < n; i++) {
        float d2 = dist2(pos4 + 4 * i, qp);
        if (!(d2 <= max_r2)) continue;
        uint64_t key = knn_key(d2, (uint32_t)i);
        if (cnt == k && key >= best[k - 1]) continue;
        int j = cnt < k ? cnt : k - 1;           /* insertion into the sorted prefix */
        while (j > 0 && best[j - 1] > key) { best[j] = best[j - 1]; j--; }
        best[j] = key;
        if (cnt < k) cnt++;
      }
      for (int j = 0; j < k; j++) {
        if (j < cnt) {
          uint32_t bits = (uint32_t)(best[j] >> 32);
          float d2;
          memcpy(&d2, &bits, 4);
          out_idx[q * k + j] = (int32_t)(uint32_t)(best[j] & 0xffffffffu);
          out_d2[q * k + j] = d2;
        } else { out_idx[q * k + j] = -1; out_d2[q * k + j] = INFINITY; }
      }
      out_cnt[q] = cnt;
    }
    free(best);
  }
}

/* Radiance estimate from a k-NN result: sum of powers (power4: rgb + ignored w) over the found photons in (d2,index)
 * order, in double, divided by the disc area pi*r_k^2 (volume == 0) or the ball volume 4/3*pi*r_k^3 (volume != 0),
 * r_k^2 = the largest d2 found; cnt == 0 or r_k == 0 gives 0. */
void pmo_knn_estimate(const float *power4, const int32_t *idx, const float *d2, const int32_t *cnt, long nq, int k,
                      int volume, float *out_rgb3) {
  const double PI = 3.14159265358979323846;
  for (long q = 0; q < nq; q++) {
    double s[3] = {0, 0, 0};
    int c = cnt[q];
    for (int j = 0; j < c; j++) {
      const float *p = power4 + 4 * (long)idx[q * k + j];
      s[0] += p[0]; s[1] += p[1]; s[2] += p[2];
    }
    double r2 = c > 0 ? (double)d2[q * k + c - 1] : 0.0, den = volume ? (4.0 / 3.0) * PI * r2 * sqrt(r2) : PI * r2;
    for (int ch = 0; ch < 3; ch++) out_rgb3[3 * q + ch] = den > 0.0 ? (float)(s[ch] / den) : 0.0f;
  }
}

/* Legacy cone-filter estimate ("photonMappingKernel - Copy.cu":191-208) over a k-NN result restricted to sq_radius:
 * for the found photons that hit wall `wall[q]` (meta type 1, same id): power * max(0, -N.dir) * (1 - sqrt(d2)) / exposure,
 * FP32 weights as the reference, accumulated in double.  meta layout: seq[0:4) | kind[4] | (type+1)[5:7) | (id+1)[7:11). */
void pmo_knn_cone_estimate(const float *pos_meta4, const float *dir4, const float *power4, const int32_t *idx, const float *d2,
                           const int32_t *cnt, const int32_t *wall, const float *normals15, float exposure, long nq, int k,
                           float *out_rgbn4) {
  for (long q = 0; q < nq; q++) {
    double s[3] = {0, 0, 0};
    int used = 0;
    int w = wall[q];
    for (int j = 0; j < cnt[q] && w >= 0 && w < 5; j++) {
      long o = idx[q * k + j];
      uint32_t meta;
      memcpy(&meta, pos_meta4 + 4 * o + 3, 4);
      int type = (int)((meta >> 5) & 3u) - 1, id = (int)((meta >> 7) & 15u) - 1;
      if (type != 1 || id != w) continue;
      const float *n = normals15 + 3 * w, *d = dir4 + 4 * o, *p = power4 + 4 * o;
      float weight = fmaxf(0.0f, -((n[0] * d[0] + n[1] * d[1]) + n[2] * d[2])) * ((1.0f - sqrtf(d2[q * k + j])) / exposure);
      s[0] += (double)(p[0] * weight); s[1] += (double)(p[1] * weight); s[2] += (double)(p[2] * weight);
      used++;
    }
    out_rgbn4[4 * q] = (float)s[0]; out_rgbn4[4 * q + 1] = (float)s[1]; out_rgbn4[4 * q + 2] = (float)s[2]; out_rgbn4[4 * q + 3] = (float)used;
  }
}
