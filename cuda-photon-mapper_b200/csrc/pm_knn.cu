// pm_knn.cu -- Mode B photon map (sm_100a): Morton keys -> hand-written LSD radix sort -> implicit 32-wide LBVH ->
// warp-cooperative k-nearest-photon search.
//
// The reference has no counterpart (its photon map is a dense voxel grid); the definition of every stage is
// the CPU oracle oracle/knn_oracle.c, and k-NN index sets are required to match it bit for bit.
//
// Layout / algorithm (all sizes for n photons):
//   keys[n']        30-bit Morton code of the position clamped to the map's world box (10 bits/axis, x lowest), n' = n
//                   rounded up to the sort tile; records that are not wall hits and the padding get key 0xFFFFFFFF
//   radix sort      4 passes x 8-bit digits over (key, index) pairs: per-tile histogram -> exclusive scan of the
//                   digit-major histogram table -> stable scatter (ranks from warp ballots + per-warp digit counters, tile staged in shared memory)
//   spos[n]         float4 (x, y, z, bits(original record index)) in sorted order: a leaf = 32 consecutive rows = one
//                   coalesced 512-byte load by one warp
//   tree            implicit and 32-wide: level 0 = leaves, level l+1 groups 32 entities of level l; only axis-aligned
//                   boxes are stored, six floats per entity, so the 32 lanes of a warp test the 32 children of a node with
//                   three 8-byte loads each (six separate arrays cost twice the instructions: address arithmetic).  The two top levels (<= 1056 boxes, 25 KB) are staged into shared
//                   memory once per CTA with a TMA bulk copy (cp.async.bulk + mbarrier).
//   query           one warp per query.  Depth-first, nearest child first (warp min-reduction over the lanes' box
//                   distances), pruned by the current k-th distance; the lowest levels are explicit nested loops (a node's 32
//                   children are evaluated once, qualifying children visited from the distances the lanes hold).  In the renderer
//                   the neighbouring pixel's k-th distance bounds the search through the triangle inequality.  Candidates of a leaf that beat the current k-th
//                   key are appended to a per-warp shared-memory buffer; every 32 candidates the buffer is sorted (bitonic,
//                   shuffles) and merged into the sorted top-K list that lives in registers (K/32 keys per lane).
//                   Keys are (bits(d2) << 32 | original index): the k smallest keys are exactly the oracle's answer.
//                   Pruning is conservative in FP32: a box's distance is computed with the same association as a point's,
//                   so by monotonicity of rounding it never exceeds the distance of any point inside it.
#include <cuda/std/limits>

#include "pm_kernels.cuh"

namespace pm {

// box e of a level: six floats (lx, ly, lz, hx, hy, hz) at b + 6 e, read as three 8-byte loads (global or shared memory)
__device__ __forceinline__ void load_box(const float *__restrict__ b, long long e, bool global, float &lx, float &ly, float &lz, float &hx,
                                         float &hy, float &hz) {
  const float2 *q = (const float2 *)b + 3 * e;
  const float2 a = global ? __ldg(q) : q[0], c = global ? __ldg(q + 1) : q[1], d = global ? __ldg(q + 2) : q[2];
  lx = a.x; ly = a.y; lz = c.x; hx = c.y; hy = d.x; hz = d.y;
}

// ---------------------------------------------------------------------------------------------------------
// Morton keys
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
  v &= 1023u;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__device__ __forceinline__ uint32_t quant10(float p, float lo, float inv_extent) {
  float f = (p - lo) * inv_extent * 1024.0f;
  if (!(f > 0.0f)) return 0u;
  if (f >= 1023.0f) return 1023u;
  return (uint32_t)f;
}
// Cell of a position: 10 bits per axis over one of three concentric boxes, tier number in key bits 30-31 (world box,
// 4x box, 192-unit box) -- identical to oracle/knn_oracle.c cell_of, which explains why out-of-box photons are neither
// clamped into the border cells nor all lumped into one coarse tier.
__device__ __forceinline__ bool inside_axis(float p, float lo, float inv_extent) {
  float f = (p - lo) * inv_extent * 1024.0f;
  return f >= -0.5f && f <= 1024.5f;
}
__device__ __forceinline__ uint32_t cell_of(float x, float y, float z, uint32_t &X0, uint32_t &X1, uint32_t &X2) {
  if (inside_axis(x, -1.5f, 1.0f / 3.0f) && inside_axis(y, -1.5f, 1.0f / 3.0f) && inside_axis(z, 0.0f, 1.0f / 6.0f)) {
    X0 = quant10(x, -1.5f, 1.0f / 3.0f); X1 = quant10(y, -1.5f, 1.0f / 3.0f); X2 = quant10(z, 0.0f, 1.0f / 6.0f);
    return 0u;
  }
  if (inside_axis(x, -6.0f, 1.0f / 12.0f) && inside_axis(y, -6.0f, 1.0f / 12.0f) && inside_axis(z, -9.0f, 1.0f / 24.0f)) {
    X0 = quant10(x, -6.0f, 1.0f / 12.0f); X1 = quant10(y, -6.0f, 1.0f / 12.0f); X2 = quant10(z, -9.0f, 1.0f / 24.0f);
    return 1u << 30;
  }
  X0 = quant10(x, -96.0f, 1.0f / 192.0f); X1 = quant10(y, -96.0f, 1.0f / 192.0f); X2 = quant10(z, -93.0f, 1.0f / 192.0f);
  return 2u << 30;
}
__device__ __forceinline__ uint32_t morton30(float x, float y, float z) {
  uint32_t X0, X1, X2, flag = cell_of(x, y, z, X0, X1, X2);
  return flag | spread10(X0) | (spread10(X1) << 1) | (spread10(X2) << 2);
}

// the same cell along the 3-D Hilbert curve (Skilling's transpose algorithm); identical integer code in oracle/knn_oracle.c
__device__ __forceinline__ uint32_t hilbert30(float x, float y, float z) {
  uint32_t X0, X1, X2, flag = cell_of(x, y, z, X0, X1, X2);
#pragma unroll
  for (uint32_t Q = 1u << 9; Q > 1; Q >>= 1) {
    const uint32_t P = Q - 1;
    if (X0 & Q) X0 ^= P;                                   // i = 0: the swap with itself is the identity
    if (X1 & Q) X0 ^= P; else { uint32_t t = (X0 ^ X1) & P; X0 ^= t; X1 ^= t; }
    if (X2 & Q) X0 ^= P; else { uint32_t t = (X0 ^ X2) & P; X0 ^= t; X2 ^= t; }
  }
  X1 ^= X0; X2 ^= X1;
  uint32_t t = 0;
#pragma unroll
  for (uint32_t Q = 1u << 9; Q > 1; Q >>= 1) if (X2 & Q) t ^= Q - 1;
  X0 ^= t; X1 ^= t; X2 ^= t;
  return flag | (spread10(X0) << 2) | (spread10(X1) << 1) | spread10(X2);
}

// filter: 0 = every row is a point; 1 = keep only wall hits (meta type == 1) of a record buffer
__global__ void __launch_bounds__(256) morton_kernel(const float4 *__restrict__ pos, long long n, long long n_pad, int filter, int curve,
                                                     uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                                     unsigned long long *__restrict__ n_valid) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  uint32_t key = 0xFFFFFFFFu;
  if (i < n) {
    float4 p = pos[i];
    bool keep = true;
    if (filter) { int seq, kind, type, id; unpack_meta(__float_as_uint(p.w), seq, kind, type, id); keep = (type == 1); }
    if (keep) key = curve ? hilbert30(p.x, p.y, p.z) : morton30(p.x, p.y, p.z);
  }
  keys[i] = key; vals[i] = (uint32_t)i;
  int kept = __syncthreads_count(key != 0xFFFFFFFFu);   // n_pad is a multiple of the block size: no thread has returned
  if (threadIdx.x == 0 && kept) atomicAdd(n_valid, (unsigned long long)kept);
}

// ---------------------------------------------------------------------------------------------------------
// radix sort: (key, value) pairs, 8-bit digits, stable
// ---------------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256, kSortRows = 16, kSortTile = kSortThreads * kSortRows;   // 4096 pairs per CTA

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint32_t *__restrict__ keys, int shift, uint32_t nb,
                                                                  uint32_t *__restrict__ ghist) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t *k = keys + (size_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int r = 0; r < kSortRows; r++) atomicAdd(&sh[(k[r * kSortThreads + threadIdx.x] >> shift) & 255u], 1u);
  __syncthreads();
  ghist[(size_t)threadIdx.x * nb + blockIdx.x] = sh[threadIdx.x];   // digit-major: a flat exclusive scan gives the offsets
}

// in-place exclusive scan of m counters (m = 256 * tiles: 1 M entries at 16 M keys) in three coalesced steps:
// per-chunk sums -> scan of the chunk sums by one CTA -> per-chunk scan with the chunk's offset
constexpr int kScanChunk = 2048, kScanThreads = 256;   // 8 counters per thread

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total, uint32_t *sh /* >= 32 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) sh[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < nw ? sh[lane] : 0u, si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += t; }
    sh[lane] = si - s;                 // exclusive offsets of the warps
    if (lane == 31 && total) *total = si;
  }
  __syncthreads();
  uint32_t r = inc - v + sh[w];
  __syncthreads();
  return r;
}
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(const uint32_t *__restrict__ a, size_t m, uint32_t *__restrict__ sums) {
  __shared__ uint32_t red[kScanThreads / 32];
  size_t base = (size_t)blockIdx.x * kScanChunk;
  uint32_t v = 0;
#pragma unroll
  for (int i = 0; i < kScanChunk / kScanThreads; i++) { size_t j = base + i * kScanThreads + threadIdx.x; v += j < m ? a[j] : 0u; }
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < kScanThreads / 32; i++) t += red[i]; sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) scan_top_kernel(uint32_t *__restrict__ sums, int n) {   // n <= a few thousand
  __shared__ uint32_t sh[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    int j = base + threadIdx.x;
    uint32_t v = j < n ? sums[j] : 0u, total = 0;
    uint32_t ex = block_exclusive_scan(v, &total, sh);
    uint32_t c = carry;
    if (j < n) sums[j] = ex + c;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + ex + v;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint32_t *__restrict__ a, size_t m, const uint32_t *__restrict__ sums) {
  __shared__ uint32_t sh[32];
  size_t base = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * (kScanChunk / kScanThreads);
  uint32_t v[kScanChunk / kScanThreads], s = 0;
#pragma unroll
  for (int i = 0; i < kScanChunk / kScanThreads; i++) { v[i] = base + i < m ? a[base + i] : 0u; s += v[i]; }
  uint32_t run = block_exclusive_scan(s, nullptr, sh) + sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanChunk / kScanThreads; i++) { if (base + i < m) a[base + i] = run; run += v[i]; }
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                                     int shift, uint32_t nb, const uint32_t *__restrict__ gbase) {
  __shared__ uint32_t wh[kSortThreads / 32][256];   // per-warp digit counters, then per-warp offsets inside the staged tile
  __shared__ uint2 stage[kSortTile];                // the tile, sorted by digit (stable), before it is written out
  __shared__ uint32_t gofs[256], scan_sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads) (&wh[0][0])[i] = 0;
  __syncthreads();
  // tile order (== stable order): warp w, row r, lane  ->  (w * rows + r) * 32 + lane
  const size_t base = (size_t)blockIdx.x * kSortTile + (size_t)w * kSortRows * 32 + lane;
  uint32_t key[kSortRows], val[kSortRows], rank[kSortRows];
#pragma unroll
  for (int r = 0; r < kSortRows; r++) { key[r] = keys_in[base + r * 32]; val[r] = vals_in[base + r * 32]; }
#pragma unroll
  for (int r = 0; r < kSortRows; r++) {
    uint32_t d = (key[r] >> shift) & 255u;
    unsigned peers = 0xffffffffu;   // lanes with the same digit, from eight ballots (MATCH.ANY is far slower than VOTE here)
#pragma unroll
    for (int b = 0; b < 8; b++) { unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1u); peers &= ((d >> b) & 1u) ? bal : ~bal; }
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader) { old = wh[w][d]; wh[w][d] = old + __popc(peers); }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  {   // thread d: where digit d starts inside the tile (scan over digits) and inside each warp's share of it (scan over warps)
    uint32_t tot = 0;
#pragma unroll
    for (int i = 0; i < kSortThreads / 32; i++) tot += wh[i][threadIdx.x];
    uint32_t start = block_exclusive_scan(tot, nullptr, scan_sh);
    gofs[threadIdx.x] = gbase[(size_t)threadIdx.x * nb + blockIdx.x] - start;   // global position = gofs[d] + position in the staged tile
    uint32_t run = start;
#pragma unroll
    for (int i = 0; i < kSortThreads / 32; i++) { uint32_t t = wh[i][threadIdx.x]; wh[i][threadIdx.x] = run; run += t; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRows; r++) {
    uint32_t d = (key[r] >> shift) & 255u;
    stage[wh[w][d] + rank[r]] = make_uint2(key[r], val[r]);
  }
  __syncthreads();
  // write-out: consecutive threads take consecutive staged elements, i.e. consecutive addresses within a digit run, instead
  // of 32 scattered 4-byte stores per warp instruction (the unstaged version was bound by L2 sector writes)
#pragma unroll 4
  for (int i = threadIdx.x; i < kSortTile; i += kSortThreads) {
    uint2 kv = stage[i];
    uint32_t p = gofs[(kv.x >> shift) & 255u] + (uint32_t)i;
    keys_out[p] = kv.x; vals_out[p] = kv.y;
  }
}

// sorted rows: (x, y, z, bits(original index)), and -- a warp's 32 rows being exactly one leaf -- the leaf boxes
// (six floats per leaf, p_out leaves; leaves >= n_out, the padding, stay empty)
__device__ __forceinline__ float warp_min(float v);
__device__ __forceinline__ float warp_max(float v);
__global__ void __launch_bounds__(256) permute_kernel(const float4 *__restrict__ pos, const uint32_t *__restrict__ vals, long long n,
                                                      float4 *__restrict__ spos, long long n_out, long long p_out, float *__restrict__ boxes) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long e = i >> 5;
  if (e >= p_out) return;
  const float inf = cuda::std::numeric_limits<float>::infinity();
  float lx = inf, ly = inf, lz = inf, hx = -inf, hy = -inf, hz = -inf;
  if (i < n) {
    uint32_t v = vals[i];
    float4 p = pos[v];
    spos[i] = make_float4(p.x, p.y, p.z, __uint_as_float(v));
    lx = hx = p.x; ly = hy = p.y; lz = hz = p.z;
  } else {   // rows of the padding leaves and the tail of the last leaf: NaN, so that a search needs no bounds test (no distance passes)
    const float nan = __int_as_float(0x7fc00000);
    spos[i] = make_float4(nan, nan, nan, 0.0f);
  }
  lx = warp_min(lx); ly = warp_min(ly); lz = warp_min(lz); hx = warp_max(hx); hy = warp_max(hy); hz = warp_max(hz);
  if ((threadIdx.x & 31) == 0) {
    float2 *q = (float2 *)boxes + 3 * e;
    q[0] = make_float2(lx, ly); q[1] = make_float2(lz, hx); q[2] = make_float2(hy, hz);
  }
}

// ---------------------------------------------------------------------------------------------------------
// boxes: one warp per parent, lane = child
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__global__ void __launch_bounds__(256) node_box_kernel(const float *__restrict__ in, long long n_in, long long p_in, long long n_out,
                                                       long long p_out, float *__restrict__ out) {
  long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (e >= p_out) return;
  const float inf = cuda::std::numeric_limits<float>::infinity();
  float lx = inf, ly = inf, lz = inf, hx = -inf, hy = -inf, hz = -inf;
  long long c = e * 32 + lane;
  if (e < n_out && c < n_in) load_box(in, c, true, lx, ly, lz, hx, hy, hz);
  lx = warp_min(lx); ly = warp_min(ly); lz = warp_min(lz); hx = warp_max(hx); hy = warp_max(hy); hz = warp_max(hz);
  if (lane == 0) {
    float2 *q = (float2 *)out + 3 * e;
    q[0] = make_float2(lx, ly); q[1] = make_float2(lz, hx); q[2] = make_float2(hy, hz);
  }
}

// ---------------------------------------------------------------------------------------------------------
// k-NN query
// ---------------------------------------------------------------------------------------------------------
typedef unsigned long long u64;
constexpr u64 kMaxKey = ~0ull;
#ifdef PM_KNN_STATS
__device__ unsigned long long g_knn_stats[8];   // 0 leaves, 1 nodes, 2 candidates passed, 3 merges, 4 queries
#define KSTAT(i, v) do { if (lane == 0) atomicAdd(&g_knn_stats[i], (unsigned long long)(v)); } while (0)
#else
#define KSTAT(i, v) do { } while (0)
#endif

__device__ __forceinline__ void cmpx(u64 &a, u64 other, bool keep_min) { a = (keep_min == (other < a)) ? other : a; }

// bitonic sort of one key per lane, ascending by lane
__device__ __forceinline__ u64 warp_sort32(u64 v, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
    for (int stride = size >> 1; stride; stride >>= 1) {
      u64 o = __shfl_xor_sync(0xffffffffu, v, stride);
      bool asc = (lane & size) == 0, lower = (lane & stride) == 0;
      cmpx(v, o, asc == lower);
    }
  return v;
}

// top-K list: K = 32*KL keys, global position i lives in slot i/32 of lane i%32, ascending
template <int KL>
struct TopK {
  u64 s[KL];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < KL; i++) s[i] = kMaxKey;
  }
  // merge 32 candidates (one per lane, any order, kMaxKey = none) keeping the K smallest
  __device__ __forceinline__ void merge(u64 cand, int lane) {
    cand = warp_sort32(cand, lane);
    u64 rev = __shfl_sync(0xffffffffu, cand, 31 - lane);
    s[KL - 1] = rev < s[KL - 1] ? rev : s[KL - 1];   // lower half of (list, padded candidates): a bitonic sequence
#pragma unroll
    for (int st = KL / 2; st; st >>= 1)              // strides >= 32: between slots of the same lane
#pragma unroll
      for (int i = 0; i < KL; i++)
        if ((i & st) == 0) { u64 a = s[i], b = s[i | st]; s[i] = a < b ? a : b; s[i | st] = a < b ? b : a; }
#pragma unroll
    for (int stride = 16; stride; stride >>= 1)      // strides < 32: between lanes of the same slot
#pragma unroll
      for (int i = 0; i < KL; i++) { u64 o = __shfl_xor_sync(0xffffffffu, s[i], stride); cmpx(s[i], o, (lane & stride) == 0); }
  }
  __device__ __forceinline__ u64 at(int pos) const {   // warp-uniform pos
    u64 v = s[0];
#pragma unroll
    for (int i = 1; i < KL; i++) v = (pos >> 5) == i ? s[i] : v;
    return __shfl_sync(0xffffffffu, v, pos & 31);
  }
};

struct TreeView {
  const float4 *spos;
  long long n;            // points
  int levels;             // number of box levels (level 0 = leaves); 0 when n == 0
  long long cnt[8];       // entities per level
  long long pad[8];       // padded (multiple of 32) array length per level
  const float *box[8];    // pad[l] entities of six floats (lx, ly, lz, hx, hy, hz) each
  int staged_from;        // levels >= staged_from are read from shared memory
  long long staged_floats;
  int staged_off[8];      // float offset of a staged level inside the shared-memory copy
};

constexpr int kQueryThreads = 256;
constexpr int kMaxStagedFloats = 6 * (1024 + 32);

__device__ __forceinline__ float box_dist2(float lx, float ly, float lz, float hx, float hy, float hz, float qx, float qy, float qz) {
  float dx = fmaxf(fmaxf(lx - qx, 0.0f), qx - hx), dy = fmaxf(fmaxf(ly - qy, 0.0f), qy - hy), dz = fmaxf(fmaxf(lz - qz, 0.0f), qz - hz);
  return (dx * dx + dy * dy) + dz * dz;
}

// The search itself, for one warp and one query point; returns with `top` holding the k smallest keys.
template <int KL>
__device__ __forceinline__ void knn_search(const TreeView &tv, const float *__restrict__ sbox, u64 *__restrict__ pend, float qx, float qy,
                                           float qz, int k, float max_r2, int lane, TopK<KL> &top) {
  top.init();
  float thr_d2 = max_r2;        // prune boxes farther than this; candidates need d2 <= max_r2 and key < thr_key
  u64 thr_key = kMaxKey;
  int npend = 0;
  if (tv.n <= 0) return;
  // a NaN query is at distance NaN from every photon and finds nothing; it must not walk the tree either, where fmaxf drops the NaN
  // and every box -- the empty padding boxes included -- comes out at distance 0
  if (!(qx == qx && qy == qy && qz == qz)) return;
  KSTAT(4, 1);

  auto leaf = [&](long long e) {
    u64 key = kMaxKey;
    {   // rows past the last photon are NaN (permute_kernel): their distance passes no test
      float4 p = __ldg(tv.spos + (e * 32 + lane));
      float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
      float d2 = (dx * dx + dy * dy) + dz * dz;
      if (d2 <= max_r2) key = ((u64)__float_as_uint(d2) << 32) | __float_as_uint(p.w);
    }
    bool pass = key < thr_key;
    unsigned m = __ballot_sync(0xffffffffu, pass);
    KSTAT(0, 1); KSTAT(2, __popc(m));
    if (!m) return;
    if (pass) pend[npend + __popc(m & ((1u << lane) - 1u))] = key;
    npend += __popc(m);
    __syncwarp();
    if (npend >= 32) {
      KSTAT(3, 1);
      top.merge(pend[lane], lane);
      u64 rest = lane + 32 < npend ? pend[lane + 32] : kMaxKey;
      __syncwarp();
      pend[lane] = rest;
      npend -= 32;
      __syncwarp();
      thr_key = top.at(k - 1);
      if (thr_key != kMaxKey) thr_d2 = fminf(max_r2, __uint_as_float((uint32_t)(thr_key >> 32)));
    }
  };

  // distance of this lane's child (entity e of level cl) from the query: infinity when the child does not exist
  auto child_dist = [&](int cl, long long e) -> float {   // e < pad[cl]: the padding entities are empty boxes, infinitely far away
    float lx, ly, lz, hx, hy, hz;
    if (cl >= tv.staged_from) load_box(sbox + tv.staged_off[cl], e, false, lx, ly, lz, hx, hy, hz);
    else load_box(tv.box[cl], e, true, lx, ly, lz, hx, hy, hz);
    return box_dist2(lx, ly, lz, hx, hy, hz, qx, qy, qz);
  };
  // The two lowest levels are explicit loops: a node's 32 children are evaluated ONCE, every child that still qualifies is visited
  // nearest first from the distances the lanes already hold (a visit can only lower thr_d2, so a lane re-tests its own distance
  // instead of the node being evaluated again after every child).  Measured: the leaf loop alone took the 16M-photon k = 50 gather
  // from 171 to 136 ms.
  auto visit_leaves = [&](long long j) {   // j: a node of level 1, its children are leaves
    const float d = child_dist(0, j * 32 + lane);
    KSTAT(1, 1);
    const uint32_t bits = __float_as_uint(d);
    bool ok = d <= thr_d2 && bits < 0x7f800000u;
    uint32_t mn = __reduce_min_sync(0xffffffffu, ok ? bits : 0xffffffffu);
    while (mn != 0xffffffffu) {
      const int c = __ffs(__ballot_sync(0xffffffffu, ok && bits == mn)) - 1;
      if (lane == c) ok = false;
      leaf(j * 32 + c);
      ok = ok && d <= thr_d2;
      mn = __reduce_min_sync(0xffffffffu, ok ? bits : 0xffffffffu);
    }
  };
  auto visit_level1 = [&](long long j) {   // j: a node of level 2, its children are nodes of level 1
    const float d = child_dist(1, j * 32 + lane);
    KSTAT(1, 1);
    const uint32_t bits = __float_as_uint(d);
    bool ok = d <= thr_d2 && bits < 0x7f800000u;
    uint32_t mn = __reduce_min_sync(0xffffffffu, ok ? bits : 0xffffffffu);
    while (mn != 0xffffffffu) {
      const int c = __ffs(__ballot_sync(0xffffffffu, ok && bits == mn)) - 1;
      if (lane == c) ok = false;
      visit_leaves(j * 32 + c);
      ok = ok && d <= thr_d2;
      mn = __reduce_min_sync(0xffffffffu, ok ? bits : 0xffffffffu);
    }
  };

  // the levels above: per-level traversal state is warp-uniform; level l's (node, visited mask) is parked in lane l's registers
  long long my_node = 0; unsigned my_mask = 0;
  int level = tv.levels;            // "virtual" level above the top: its single node 0 has the top-level entities as children
  if (level == 1) visit_leaves(0);
  else if (level == 2) visit_level1(0);
  else for (;;) {
    // children of node `j` at `level` are entities 32*j .. 32*j+31 of level-1 (level >= 3 here)
    long long j = __shfl_sync(0xffffffffu, my_node, level);
    unsigned visited = __shfl_sync(0xffffffffu, my_mask, level);
    int cl = level - 1;
    float d = cuda::std::numeric_limits<float>::infinity();
    if (!((visited >> lane) & 1u)) d = child_dist(cl, j * 32 + lane);
    KSTAT(1, 1);
    uint32_t bits = __float_as_uint(d);
    bool ok = d <= thr_d2 && bits < 0x7f800000u;
    uint32_t mn = __reduce_min_sync(0xffffffffu, ok ? bits : 0xffffffffu);
    if (mn == 0xffffffffu) {          // nothing (left) to visit under this node: pop
      level++;
      if (level > tv.levels) break;
      continue;
    }
    int c = __ffs(__ballot_sync(0xffffffffu, ok && bits == mn)) - 1;
    if (lane == level) my_mask |= 1u << c;
    long long child = j * 32 + c;
    if (cl == 2) visit_level1(child);   // the whole subtree at once; the next pass over this node picks its next child
    else { level = cl; if (lane == level) { my_node = child; my_mask = 0; } }
  }
  if (npend > 0) {
    top.merge(lane < npend ? pend[lane] : kMaxKey, lane);
    __syncwarp();
  }
}

// stage the top box levels into shared memory with one TMA bulk copy per CTA
__device__ __forceinline__ void stage_top_levels(const TreeView &tv, float *sbox, unsigned long long *bar) {
  if (tv.staged_floats <= 0) return;
  const uint32_t bytes = (uint32_t)(tv.staged_floats * sizeof(float));
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar), dst_a = (uint32_t)__cvta_generic_to_shared(sbox);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
                 "l"(tv.box[tv.staged_from]), "r"(bytes), "r"(bar_a)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar_a) : "memory");
  }
}

// radiance estimate from a finished search: (sum of the found powers / (pi r_k^2) or (4/3 pi r_k^3), r_k^2), on every lane
template <int KL>
__device__ __forceinline__ float4 knn_radiance(const TopK<KL> &top, int k, const float4 *__restrict__ power, int volume, int lane) {
  float r = 0.0f, g = 0.0f, b = 0.0f, rk2 = 0.0f;
#pragma unroll
  for (int i = 0; i < KL; i++) {
    int pos = i * 32 + lane;
    if (pos < k && top.s[i] != kMaxKey) {
      float4 pw = __ldg(power + (uint32_t)(top.s[i] & 0xffffffffu));
      r += pw.x; g += pw.y; b += pw.z;
      rk2 = fmaxf(rk2, __uint_as_float((uint32_t)(top.s[i] >> 32)));
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    r += __shfl_xor_sync(0xffffffffu, r, o); g += __shfl_xor_sync(0xffffffffu, g, o); b += __shfl_xor_sync(0xffffffffu, b, o);
    rk2 = fmaxf(rk2, __shfl_xor_sync(0xffffffffu, rk2, o));
  }
  const float PI = 3.14159265358979323846f;
  float den = volume ? (4.0f / 3.0f) * PI * rk2 * __fsqrt_rn(rk2) : PI * rk2;
  float inv = den > 0.0f ? __fdiv_rn(1.0f, den) : 0.0f;
  return make_float4(r * inv, g * inv, b * inv, rk2);
}

// The only point-based estimator the reference's author wrote: the fixed-radius cone filter of the legacy file
// (photonMappingKernel - Copy.cu:191-208): over the photons that hit the SAME object within sqRadius,
//   energy += power * max(0, -dot(N, dir)) * (1 - sqrt(d2)) / exposure.
// Here it is evaluated over the (at most k) nearest photons within the radius that the search returned.
struct ConeArgs {
  const float4 *pos_meta, *dir;   // per ORIGINAL record index
  float normal[PM_MAX_PLANES][3]; // inward wall normals (surfaceNormal(type 1, id, p, gOrigin))
  float exposure;
  int enabled;
};
template <int KL>
__device__ __forceinline__ float4 knn_cone(const TopK<KL> &top, int k, const float4 *__restrict__ power, const ConeArgs &ca, int wall, int lane) {
  float r = 0.0f, g = 0.0f, b = 0.0f, n = 0.0f;
  const float nx = ca.normal[wall][0], ny = ca.normal[wall][1], nz = ca.normal[wall][2];
#pragma unroll
  for (int i = 0; i < KL; i++) {
    int pos = i * 32 + lane;
    if (pos < k && top.s[i] != kMaxKey) {
      uint32_t o = (uint32_t)(top.s[i] & 0xffffffffu);
      int seq, kind, type, id;
      unpack_meta(__float_as_uint(__ldg(&ca.pos_meta[o].w)), seq, kind, type, id);
      if (type == 1 && id == wall) {
        float4 d = __ldg(ca.dir + o), pw = __ldg(power + o);
        float d2 = __uint_as_float((uint32_t)(top.s[i] >> 32));
        float w = fmaxf(0.0f, -((nx * d.x + ny * d.y) + nz * d.z)) * __fdiv_rn(1.0f - __fsqrt_rn(d2), ca.exposure);
        r += pw.x * w; g += pw.y * w; b += pw.z * w; n += 1.0f;
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    r += __shfl_xor_sync(0xffffffffu, r, o); g += __shfl_xor_sync(0xffffffffu, g, o); b += __shfl_xor_sync(0xffffffffu, b, o);
    n += __shfl_xor_sync(0xffffffffu, n, o);
  }
  return make_float4(r, g, b, n);
}

// mode 0: write indices / distances / counts.  mode 1: radiance estimate (sum of powers / disc area or ball volume)
template <int KL>
__global__ void __launch_bounds__(kQueryThreads) knn_query_kernel(const __grid_constant__ TreeView tv, const float4 *__restrict__ queries,
                                                                  long long nq, int k, float max_r2, int32_t *__restrict__ out_idx,
                                                                  float *__restrict__ out_d2, int32_t *__restrict__ out_cnt,
                                                                  const float4 *__restrict__ power, int volume, float4 *__restrict__ out_rgb,
                                                                  const __grid_constant__ ConeArgs cone) {
  __shared__ __align__(128) float sbox[kMaxStagedFloats];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ u64 pend_all[kQueryThreads / 32][64];
  stage_top_levels(tv, sbox, &bar);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  u64 *pend = pend_all[w];
  const long long warps = (long long)gridDim.x * (kQueryThreads / 32);
  for (long long q = (long long)blockIdx.x * (kQueryThreads / 32) + w; q < nq; q += warps) {
    float4 qp = __ldg(queries + q);
    TopK<KL> top;
    knn_search<KL>(tv, sbox, pend, qp.x, qp.y, qp.z, k, max_r2, lane, top);
    if (out_rgb) {
      int wall = (int)qp.w;
      float4 est = cone.enabled ? (((unsigned)wall < (unsigned)PM_MAX_PLANES) ? knn_cone<KL>(top, k, power, cone, wall, lane) : make_float4(0.f, 0.f, 0.f, 0.f))
                                : knn_radiance<KL>(top, k, power, volume, lane);
      if (lane == 0) out_rgb[q] = est;
    } else {
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < KL; i++) {
        int pos = i * 32 + lane;
        if (pos < k) {
          bool have = top.s[i] != kMaxKey;
          out_idx[q * k + pos] = have ? (int32_t)(uint32_t)(top.s[i] & 0xffffffffu) : -1;
          out_d2[q * k + pos] = have ? __uint_as_float((uint32_t)(top.s[i] >> 32)) : cuda::std::numeric_limits<float>::infinity();
          cnt += have ? 1 : 0;
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (lane == 0) out_cnt[q] = cnt;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Mode B frame: one warp per pixel.  Eye ray and mirror/glass chain as in render_kernel (PMK:926-983), then the
// reference's voxel gathers are replaced by k-nearest-photon estimates: ten ray-march samples in the volume map
// (PMK:937-965) and one estimate at the wall hit in the surface map (PMK:987-999), composited as the reference does
// (media: rgb = sum of the march terms + 0.15 * wall term).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned char quantise_u8(float v) {   // PMK:1451-1453 with the device's saturating cast
  double d = (double)v * 255.0;
  d = d > 255.0 ? 255.0 : d;
  return d > 0.0 ? (unsigned char)__double2uint_rz(d) : (unsigned char)0;
}

#ifndef PM_KNN_MINBLOCKS
#define PM_KNN_MINBLOCKS 4   /* 64 registers: four 256-thread blocks per SM (uncapped the kernel takes 101 and runs 40 % slower) */
#endif
template <int KL>
__global__ void __launch_bounds__(kQueryThreads, PM_KNN_MINBLOCKS) knn_render_kernel(const __grid_constant__ DeviceScene sc, const __grid_constant__ TreeView tvs,
                                                                   const __grid_constant__ TreeView tvv, const float4 *__restrict__ pow_s,
                                                                   const float4 *__restrict__ pow_v, int k, float max_r2, float w_surf,
                                                                   float w_vol, int width, int height, int y0, int y1, int y_step, int media,
                                                                   uchar4 *__restrict__ rgba, float4 *__restrict__ rgbf,
                                                                   unsigned long long *__restrict__ work_counter, int staged_stride) {
  // dynamic shared memory: the staged top levels of the two trees (staged_stride floats each: what the trees need, not the
  // kMaxStagedFloats they may need -- 2.3 KB instead of 25 KB per tree at 16 M photons, which leaves the rest of the SM's 256 KB to L1),
  // the warps' candidate buffers, two mbarriers
  extern __shared__ __align__(128) unsigned char dyn[];
  float *sbox_s = (float *)dyn, *sbox_v = sbox_s + staged_stride;
  u64 *pend_all = (u64 *)(sbox_v + staged_stride);
  unsigned long long *bars = (unsigned long long *)(pend_all + (kQueryThreads / 32) * 64);
  stage_top_levels(tvs, sbox_s, bars + 0);
  if (media) stage_top_levels(tvv, sbox_v, bars + 1);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  u64 *pend = pend_all + w * 64;
  const int nrows = (y1 - y0 + y_step - 1) / y_step;   // rows y0, y0+y_step, ... < y1
  const float inf = cuda::std::numeric_limits<float>::infinity();
  // Work decomposition: the frame is cut into units of kRun consecutive pixels of one row; units are numbered so that
  // eight consecutive ids are vertically adjacent (a tile of 8 rows x kRun columns), and every warp fetches its next unit
  // from a global counter.  Warps that run at the same time therefore work on neighbouring pixels (shared leaves in
  // L1/L2), no warp is stuck with a long static list of expensive pixels (the gather cost varies by orders of magnitude
  // over the image: a static split left the SMs 78 % idle), and
  // a warp's consecutive pixels are neighbours, so the k-th distance found for one pixel can bound the search of the next
  // (same wall / same march step): searching within 1.2x the neighbour's radius first skips the phase in which every
  // candidate passes.  If fewer than k photons lie within the hinted radius the search is repeated without it, so the
  // result is exact either way.  Hints: lane i holds the hint of march step i, lane 10 the wall hint.
#ifndef PM_KNN_RUN
#define PM_KNN_RUN 16
#endif
  constexpr int kRun = PM_KNN_RUN, kRows = kQueryThreads / 32;
  const int tiles_x = (width + kRun - 1) / kRun, tiles_y = (nrows + kRows - 1) / kRows;
  // `step` = distance of this query from the one the hint belongs to: its k nearest photons lie within sqrt(hint) + step of this
  // query (triangle inequality), so that radius -- with a margin for the FP32 roundings of both distances -- needs no retry in exact
  // arithmetic and is usually tighter than the 1.2x guess; the retry below still covers every rounding case
  auto search = [&](const TreeView &tv, const float *sbox, float qx, float qy, float qz, float hint_r2, float step, TopK<KL> &top) -> float {
    const float tri = (__fsqrt_rn(hint_r2) + step) * 1.00001f;
    float lim = fminf(max_r2, fminf(hint_r2 * 1.44f, tri * tri));
    KSTAT(6, lim < max_r2 ? 1 : 0);
    u64 kth;
    for (;;) {
      knn_search<KL>(tv, sbox, pend, qx, qy, qz, k, lim, lane, top);
      kth = top.at(k - 1);
      if (kth != kMaxKey || !(lim < max_r2)) break;
      KSTAT(5, 1);
      lim = max_r2;                            // the hint was too tight (or the map holds fewer than k photons): unbounded retry
    }
    return kth == kMaxKey ? inf : __uint_as_float((uint32_t)(kth >> 32));
  };
  const long long units = (long long)tiles_x * tiles_y * kRows;
  for (;;) {
   long long u = 0;
   if (lane == 0) u = (long long)atomicAdd(work_counter, 1ull);
   u = __shfl_sync(0xffffffffu, u, 0);
   if (u >= units) break;
   const long long tile = u / kRows;
   const int row = (int)(tile / tiles_x) * kRows + (int)(u % kRows), x0 = (int)(tile % tiles_x) * kRun;
   if (row >= nrows) continue;
   float hint = inf;
   v3 ray_prev = V(0.0f, 0.0f, 0.0f), P_prev = ray_prev;
   for (int px = x0; px < width && px < x0 + kRun; px++) {
    const int py = y0 + row * y_step;
    const long long pix = (long long)py * width + px;
#ifdef PM_KNN_STATS
    long long t_begin = clock64();
#endif
    float x = (float)px + sc.cam_ox, y = (float)py + sc.cam_oy;
    v3 rgb = V(0.0f, 0.0f, 0.0f);
    const v3 origin = V(0.0f, 0.0f, 0.0f);
    v3 ray = V((float)((double)__fdiv_rn(x, sc.sz_img) - 0.5), (float)(-((double)__fdiv_rn(y, sc.sz_img) - 0.5)), 1.0f);
    TopK<KL> top;
    const v3 dray = sub(ray, ray_prev);
    const float dr = __fsqrt_rn(dot(dray, dray)) * 0.6001f;   // march step i of neighbouring pixels: (i + 1) * 0.6 * |ray - ray_prev| apart
    ray_prev = ray;
    if (media) {
      v3 prev = origin;
#pragma unroll 1
      for (int i = 0; i < 10; i++) {
        prev = add(mul(ray, 0.6f), prev);
        float r2 = search(tvv, sbox_v, prev.x, prev.y, prev.z, __shfl_sync(0xffffffffu, hint, i), (float)(i + 1) * dr, top);
        if (lane == i) hint = r2;
        float4 e = knn_radiance<KL>(top, k, pow_v, 1, lane);
        rgb = add(rgb, mul(V(e.x, e.y, e.z), w_vol));
      }
    }
    Hit h; h.hit = 0; h.type = 0; h.idx = 0; h.dist = -1.0f;
    raytrace(sc, ray, origin, h);
    bool wall = false;
    if (h.hit) {
      v3 P = mul(ray, h.dist);
      if (h.type == 0 && h.idx == 1) follow_specular(sc, ray, origin, h, P, 1);
      else if (h.type == 0 && h.idx == 0) follow_specular(sc, ray, origin, h, P, 0);
      if (h.hit && h.type == 1) {   // warp-uniform: every lane traced the same ray
        wall = true;
        const v3 dP = sub(P, P_prev);
        float r2 = search(tvs, sbox_s, P.x, P.y, P.z, __shfl_sync(0xffffffffu, hint, 10), __fsqrt_rn(dot(dP, dP)), top);
        P_prev = P;
        if (lane == 10) hint = r2;
        float4 e = knn_radiance<KL>(top, k, pow_s, 0, lane);
        v3 c = mul(V(e.x, e.y, e.z), w_surf);
        rgb = media ? add(rgb, mul(c, 0.15f)) : add(rgb, c);
      }
    }
    if (!wall && lane == 10) hint = inf;
    if (lane == 0) {
#ifdef PM_KNN_STATS
      if (rgbf) rgbf[pix] = make_float4(rgb.x, rgb.y, rgb.z, (float)(clock64() - t_begin));   // debug: cycles per pixel
#else
      if (rgbf) rgbf[pix] = make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
#endif
      if (rgba) rgba[pix] = make_uchar4(quantise_u8(rgb.x), quantise_u8(rgb.y), quantise_u8(rgb.z), 0);
    }
   }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Batched search (round 2): one warp = 32 NEIGHBOURING queries, one lane = one query.
//
// The warp-per-query search above spends its instructions on warp-wide machinery per candidate: 7 600 warp instructions per query
// (ncu, 16M photons, k = 50), a bitonic sort + merge network for every 32 candidates, a node step per visited child.  The queries
// of a frame are coherent -- the pixels of an 8 x 4 tile at the same march step lie within a fraction of the search radius of each
// other -- so here the 32 lanes walk the tree ONCE for their common bounding box and every lane tests every visited photon against
// its own query (8 thread instructions per test, all lanes busy):
//   * traversal: warp-uniform depth-first walk; the 32 lanes test the 32 children of a node against the batch's box expanded by the
//     largest search radius (conservative in FP32 for the same reason as above), the set of children to visit is computed once per
//     node and parked in a lane register per level; before a leaf is read, every lane tests the leaf's box against ITS sphere;
//   * a leaf = 32 photons = one coalesced 512-byte load, staged in shared memory and broadcast to the lanes;
//   * the limit of a lane comes from a neighbouring query's k-th distance (x 1.2, as above); lanes without one, or whose limit holds
//     fewer than k photons, are seeded: the warp-per-query search gives the k-th distance r of one of their points p, and
//     r_k(q) <= r + |q - p| (triangle inequality) bounds the others, so the second attempt cannot fail;
//   * selection without per-lane lists: the tree is walked TWICE.  Walk 1 only counts, per lane, the photons inside the limit in 32
//     bins of d2.  The bin j that holds the k-th nearest splits them: every photon in a lower bin is one of the k nearest and is
//     accumulated directly in walk 2 (power sum, largest distance); the photons of bin j -- a handful -- go to a short per-lane
//     list, of which the k - (count below) smallest keys are then extracted.  Both walks classify a photon with the same expression,
//     and bin(d2) is monotone in d2, so the selected set is exactly the k smallest keys (bits(d2) << 32 | original index).
//     (A first version collected every candidate in a 96-entry list per lane and selected with a heap or quickselect: data-dependent
//     loops that 32 lanes execute one after the other, 60-70 K cycles per selection, and 27 KB of shared memory per warp -- 8 warps
//     per SM.  It was 2x slower than the warp-per-query renderer.)
//
// MEASURED (B200, tools/time_knn.py, tools/knn_bstats.py): exact -- the frames agree with the warp-per-query renderer to 1e-8 -- but
// SLOWER: 135 ms vs 16.1 ms (4M photons, k = 100) and 667 ms vs 195 ms (16M photons, k = 50, 11 gathers per pixel).  The premise does
// not hold at BASELINE's densities: with 38 M wall photons a pixel's 50 nearest lie within 1.3 pixel spacings, so neighbouring
// pixels share almost no photons; a tile's common walk touches 108 leaves where one query needs 16-30, every lane tests all of
// them (a lane's test of a 32-photon leaf costs what the warp-per-query search pays per leaf: ~12 vs ~15 instructions per
// (query, leaf) pair), and 16 warps per SM sustain a third of the issue rate.  Kept behind pm_knn_set_batched (off by default) as a
// tested alternative; it wins only where the search radius spans many pixels.
// ---------------------------------------------------------------------------------------------------------
#ifdef PM_KNN_BSTATS
// development counters: 0 batches, 1 walks, 2 node visits, 3 leaves listed, 4 leaves needed, 5 boundary overflows, 6 seeds,
// 7 lanes searched one by one, 8 cycles in walks, 9 cycles selecting, 10 cycles seeding, 11 cycles one by one, 14 cycles per tile
__device__ unsigned long long g_bstats[16];
#define BSTAT(i, v) do { if (lane == 0) atomicAdd(&g_bstats[i], (unsigned long long)(v)); } while (0)
#define BCLK() clock64()
#else
#define BSTAT(i, v) do { } while (0)
#define BCLK() 0ll
#endif

constexpr int kBatchStrip = 4;           // tiles per work unit (every unit starts with one seed search per query kind)
constexpr int kBins = 32;                // bins of d2 over [0, limit]
constexpr int kBoundary = 32;            // per-lane list for the photons of the bin that holds the k-th nearest
constexpr int kBatchWarps = 16;          // warps per CTA of the batched renderer
struct BatchShared {                     // per warp: 12.6 KB
  float4 leaf[32];
  unsigned short bins[kBins][32];
  float bd2[kBoundary][32];
  uint32_t bidx[kBoundary][32];
  float hint[11][32];                    // per lane: k-th squared distance of the last search per query kind (10 march steps + wall)
  u64 pend[64];                          // scratch of the warp-per-query search (seeds, incoherent tiles)
};

__device__ __forceinline__ u64 cand_key(float d2, uint32_t idx) { return ((u64)__float_as_uint(d2) << 32) | idx; }
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One walk of the tree for the whole warp.  Lane state: query (qx, qy, qz), lim = visit photons with d2 <= lim (negative: the lane
// takes no part).  point(d2, original index) is called by a lane for every photon inside its limit.
template <class PointFn>
__device__ __forceinline__ void batch_walk(const TreeView &tv, float4 *__restrict__ leaf, float qx, float qy, float qz, float lim, int lane,
                                           PointFn &&point) {
  const float inf = cuda::std::numeric_limits<float>::infinity();
  if (tv.n <= 0) return;
  const bool act = lim >= 0.0f;
  // the batch's bounding box and largest radius (inactive lanes do not widen them)
  const float blx = warp_min_f(act ? qx : inf), bly = warp_min_f(act ? qy : inf), blz = warp_min_f(act ? qz : inf);
  const float bhx = warp_max_f(act ? qx : -inf), bhy = warp_max_f(act ? qy : -inf), bhz = warp_max_f(act ? qz : -inf);
  const float max_lim = warp_max_f(act ? lim : -1.0f);
  if (!(max_lim >= 0.0f)) return;
  BSTAT(1, 1);
  long long my_node = 0; unsigned my_pend = 0;      // lane l parks the state of level l: node index, children still to visit
  float lx = 0.f, ly = 0.f, lz = 0.f, hx = 0.f, hy = 0.f, hz = 0.f;   // this lane's child box of the node entered last
  int level = tv.levels;                            // the virtual level above the top: node 0's children are the top-level entities
  bool enter = true;
  long long j = 0;
  for (;;) {
    if (enter) {   // first visit of node j at `level`: which of its 32 children can hold a photon inside ANY lane's limit
      const int cl = level - 1;
      const long long e = j * 32 + lane;
      bool ok = false;
      if (e < tv.cnt[cl]) {
        load_box(tv.box[cl], e, true, lx, ly, lz, hx, hy, hz);
        const float dx = fmaxf(fmaxf(lx - bhx, 0.0f), blx - hx), dy = fmaxf(fmaxf(ly - bhy, 0.0f), bly - hy), dz = fmaxf(fmaxf(lz - bhz, 0.0f), blz - hz);
        ok = (dx * dx + dy * dy) + dz * dz <= max_lim;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      BSTAT(2, 1); if (level == 1) BSTAT(3, __popc(m));
      if (lane == level) { my_node = j; my_pend = m; }
      enter = false;
    }
    const unsigned pend = __shfl_sync(0xffffffffu, my_pend, level);
    if (!pend) {                                    // nothing left under this node
      level++;
      if (level > tv.levels) break;
      continue;
    }
    const int c = __ffs(pend) - 1;
    if (lane == level) my_pend = pend & (pend - 1);
    j = __shfl_sync(0xffffffffu, my_node, level);
    const long long child = j * 32 + c;
    if (level > 1) { level--; j = child; enter = true; continue; }
    // ---- leaf `child`: its box sits in lane c's registers (the node entered last is this leaf's parent) ----
    const float clx = __shfl_sync(0xffffffffu, lx, c), cly = __shfl_sync(0xffffffffu, ly, c), clz = __shfl_sync(0xffffffffu, lz, c);
    const float chx = __shfl_sync(0xffffffffu, hx, c), chy = __shfl_sync(0xffffffffu, hy, c), chz = __shfl_sync(0xffffffffu, hz, c);
    const bool need = act && box_dist2(clx, cly, clz, chx, chy, chz, qx, qy, qz) <= lim;
    if (!__ballot_sync(0xffffffffu, need)) continue;
    BSTAT(4, 1);
    const long long i = child * 32 + lane;
    float4 p = make_float4(inf, inf, inf, 0.0f);
    if (i < tv.n) p = __ldg(tv.spos + i);
    __syncwarp();
    leaf[lane] = p;
    __syncwarp();
    if (need) {
#pragma unroll 8
      for (int t = 0; t < 32; t++) {
        const float4 s = leaf[t];
        const float dx = s.x - qx, dy = s.y - qy, dz = s.z - qz;
        const float d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2 <= lim) point(d2, __float_as_uint(s.w));
      }
    }
  }
  __syncwarp();
}

// The radiance estimate over the exact k nearest photons of every lane's query (active = false: the lane has none): returns
// (sum of their powers / (pi r_k^2) or (4/3 pi r_k^3), r_k^2; r_k^2 = inf when fewer than k photons lie within max_r2 -- the estimate is
// then taken over those).  hint_r2: a neighbouring query's k-th squared distance (inf: none).
template <int KL>
__device__ __forceinline__ float4 batch_estimate(const TreeView &tv, BatchShared &sm, bool active, float qx, float qy, float qz, int k, float max_r2,
                                                 float hint_r2, const float4 *__restrict__ power, int volume, int lane) {
  const float inf = cuda::std::numeric_limits<float>::infinity();
  BSTAT(0, 1);
  bool seed = active && !(hint_r2 * 1.44f < max_r2);    // no finite limit of its own
  bool todo = active;
  float sr = 0.0f, sg = 0.0f, sb = 0.0f, rk2 = 0.0f;     // the lane's result: power sum and largest squared distance of the selected photons
  int found = 0;
  auto finish_one_by_one = [&](unsigned tm, const float lim_of_lane) {   // warp-per-query search for every lane of tm, result into that lane
    BSTAT(7, __popc(tm));
    const long long t_one = BCLK();
    while (tm) {
      const int src = __ffs(tm) - 1;
      tm &= tm - 1;
      const float sx = __shfl_sync(0xffffffffu, qx, src), sy = __shfl_sync(0xffffffffu, qy, src), sz = __shfl_sync(0xffffffffu, qz, src);
      float l = __shfl_sync(0xffffffffu, lim_of_lane, src);
      TopK<KL> top;
      for (;;) {
        knn_search<KL>(tv, nullptr, sm.pend, sx, sy, sz, k, l, lane, top);
        if (top.at(k - 1) != kMaxKey || !(l < max_r2)) break;
        l = max_r2;                                     // the limit was too tight: unbounded (nearest-first) retry
      }
      float r = 0.0f, g = 0.0f, b = 0.0f, m2 = 0.0f; int n = 0;
#pragma unroll
      for (int i = 0; i < KL; i++) {
        const int pos = i * 32 + lane;
        if (pos < k && top.s[i] != kMaxKey) {
          const float4 pw = __ldg(power + (uint32_t)(top.s[i] & 0xffffffffu));
          r += pw.x; g += pw.y; b += pw.z; n++;
          m2 = fmaxf(m2, __uint_as_float((uint32_t)(top.s[i] >> 32)));
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, o); g += __shfl_xor_sync(0xffffffffu, g, o); b += __shfl_xor_sync(0xffffffffu, b, o);
        n += __shfl_xor_sync(0xffffffffu, n, o); m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
      }
      if (lane == src) { sr = r; sg = g; sb = b; rk2 = m2; found = n; }
    }
    BSTAT(11, BCLK() - t_one);
  };
  for (int round = 0; round < 2; round++) {
    // Seeding.  The bound (r + |q - p|)^2 is only useful near p: a lane farther than r from the seed point would get a limit holding
    // many times k photons (and a walk counts all of them), so it waits for a seed of its own -- at worst every lane is searched
    // alone, which is what the warp-per-query renderer does anyway.
    float bound = inf;
    bool need_seed = seed;
    for (;;) {
      const unsigned seed_mask = __ballot_sync(0xffffffffu, need_seed);
      if (!seed_mask) break;
      BSTAT(6, 1);
      const long long t_seed = BCLK();
      const int src = __ffs(seed_mask) - 1;
      const float sx = __shfl_sync(0xffffffffu, qx, src), sy = __shfl_sync(0xffffffffu, qy, src), sz = __shfl_sync(0xffffffffu, qz, src);
      TopK<KL> top;
      knn_search<KL>(tv, nullptr, sm.pend, sx, sy, sz, k, max_r2, lane, top);
      const u64 kth = top.at(k - 1);
      if (kth == kMaxKey) {     // the map holds fewer than k photons within max_r2: no finite bound, the lanes go one by one below
        BSTAT(10, BCLK() - t_seed);
        break;
      }
      const float r2 = __uint_as_float((uint32_t)(kth >> 32));
      const float dx = qx - sx, dy = qy - sy, dz = qz - sz;
      const float dd = (dx * dx + dy * dy) + dz * dz;
      if (need_seed && (dd <= r2 || lane == src)) {   // a relative 2e-5 and a denormal cover the FP32 rounding of the bound itself
        const float rb = __fsqrt_rn(r2) + __fsqrt_rn(dd) * 1.000001f;
        bound = rb * rb * 1.00002f + 1e-30f;
        need_seed = false;
      }
      BSTAT(10, BCLK() - t_seed);
    }
    const float lim = !todo ? -1.0f : (seed ? fminf(max_r2, bound) : fminf(max_r2, hint_r2 * 1.44f));
    const bool guaranteed = seed;
    // A common walk only pays while the queries are neighbours.  The pixels of a tile that straddles a silhouette, or that see the
    // scene through the mirror / glass sphere, have wall points all over the scene: their common box would hold the whole map.
    // Such a batch (box diagonal > 8 x the largest radius), and one whose limit is not finite (a map with fewer than k photons),
    // is searched lane by lane with the warp-per-query search instead.
    {
      const float blx = warp_min_f(todo ? qx : inf), bly = warp_min_f(todo ? qy : inf), blz = warp_min_f(todo ? qz : inf);
      const float bhx = warp_max_f(todo ? qx : -inf), bhy = warp_max_f(todo ? qy : -inf), bhz = warp_max_f(todo ? qz : -inf);
      const float ex = bhx - blx, ey = bhy - bly, ez = bhz - blz;
      const float max_lim = warp_max_f(todo ? lim : -1.0f);
      const unsigned tm = __ballot_sync(0xffffffffu, todo);
      if (tm && !((ex * ex + ey * ey) + ez * ez <= 64.0f * max_lim && max_lim < inf)) {
        finish_one_by_one(tm, lim);
        break;
      }
    }
    // ---- walk 1: count the photons inside the limit per bin of d2 ----
    const long long t_walk = BCLK();
    const float inv_w = todo ? __fdiv_rn((float)kBins, lim) : 0.0f;   // bin(d2) = min(kBins - 1, (int)(d2 * inv_w)): monotone in d2
    auto bin_of = [&](float d2) { const int b = __float2int_rz(d2 * inv_w); return b < kBins - 1 ? b : kBins - 1; };
#pragma unroll
    for (int b = 0; b < kBins; b++) sm.bins[b][lane] = 0;
    __syncwarp();
    batch_walk(tv, sm.leaf, qx, qy, qz, lim, lane, [&](float d2, uint32_t) { sm.bins[bin_of(d2)][lane]++; });
    int jbin = -1, below = 0, total = 0;
#pragma unroll
    for (int b = 0; b < kBins; b++) {
      const int c = sm.bins[b][lane];
      if (jbin < 0 && total + c >= k) { jbin = b; below = total; }
      total += c;
    }
    // a lane whose limit holds fewer than k photons goes again, seeded -- unless its limit was already max_r2 or a guaranteed bound
    // (then the map simply holds fewer than k photons within max_r2 and all of them count)
    const bool short_of_k = todo && total < k;
    const bool final_short = short_of_k && (guaranteed || !(lim < max_r2));
    if (final_short) { jbin = kBins; below = total; }            // every bin is "below"
    const bool take = todo && (jbin >= 0);
    // ---- walk 2: accumulate the photons of the bins below jbin, list the photons of bin jbin ----
    // pruning limit: nothing beyond bin jbin matters; (jbin + 1) / inv_w, widened against the rounding of the product in bin_of
    const float lim2 = !take ? -1.0f : (jbin >= kBins - 1 ? lim : fminf(lim, __fdiv_rn((float)(jbin + 1), inv_w) * 1.00001f));
    int nb = 0;
    if (__ballot_sync(0xffffffffu, take))
      batch_walk(tv, sm.leaf, qx, qy, qz, lim2, lane, [&](float d2, uint32_t idx) {
        const int b = bin_of(d2);
        if (b < jbin) {
          const float4 pw = __ldg(power + idx);
          sr += pw.x; sg += pw.y; sb += pw.z; rk2 = fmaxf(rk2, d2);
        } else if (b == jbin) {
          if (nb < kBoundary) { sm.bd2[nb][lane] = d2; sm.bidx[nb][lane] = idx; }
          nb++;
        }
      });
    BSTAT(8, BCLK() - t_walk);
    // ---- the k - below smallest keys of the boundary list: repeated extraction of the smallest key above the last one ----
    const long long t_sel = BCLK();
    const bool overflow = take && nb > kBoundary;
    int want = (take && !overflow && jbin < kBins) ? k - below : 0;
    u64 last = 0ull; bool first = true;
    const int max_want = __reduce_max_sync(0xffffffffu, want), max_nb = __reduce_max_sync(0xffffffffu, overflow ? 0 : (take ? nb : 0));
    for (int r = 0; r < max_want; r++) {
      u64 best = kMaxKey; int bi = -1;
      for (int i = 0; i < max_nb; i++) {
        if (i < nb) {
          const u64 key = cand_key(sm.bd2[i][lane], sm.bidx[i][lane]);
          if ((first || key > last) && key < best) { best = key; bi = i; }
        }
      }
      if (r < want && bi >= 0) {
        const float4 pw = __ldg(power + sm.bidx[bi][lane]);
        sr += pw.x; sg += pw.y; sb += pw.z; rk2 = fmaxf(rk2, sm.bd2[bi][lane]);
        last = best; first = false;
      }
    }
    BSTAT(9, BCLK() - t_sel);
    if (take && !overflow) found = final_short ? total : k;
    // lanes whose boundary bin overflowed the list (a very uneven density inside the limit) are searched one by one
    const unsigned om = __ballot_sync(0xffffffffu, overflow);
    if (om) {
      BSTAT(5, __popc(om));
      if (overflow) { sr = sg = sb = rk2 = 0.0f; }
      finish_one_by_one(om, lim);
    }
    seed = short_of_k && !final_short;
    todo = seed;
    if (seed) { sr = sg = sb = rk2 = 0.0f; }
    if (!__ballot_sync(0xffffffffu, todo)) break;
  }
  const float PI = 3.14159265358979323846f;
  const float den = volume ? (4.0f / 3.0f) * PI * rk2 * __fsqrt_rn(rk2) : PI * rk2;
  const float inv = den > 0.0f ? __fdiv_rn(1.0f, den) : 0.0f;
  return make_float4(sr * inv, sg * inv, sb * inv, found >= k ? rk2 : inf);
}

template <int KL>
__global__ void __launch_bounds__(kBatchWarps * 32) knn_render_batched_kernel(const __grid_constant__ DeviceScene sc, const __grid_constant__ TreeView tvs,
                                                                              const __grid_constant__ TreeView tvv, const float4 *__restrict__ pow_s,
                                                                              const float4 *__restrict__ pow_v, int k, float max_r2, float w_surf,
                                                                              float w_vol, int width, int height, int y0, int y1, int y_step, int media,
                                                                              uchar4 *__restrict__ rgba, float4 *__restrict__ rgbf,
                                                                              unsigned long long *__restrict__ work_counter) {
  extern __shared__ __align__(128) unsigned char dyn[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  BatchShared &sm = reinterpret_cast<BatchShared *>(dyn)[w];
  const float inf = cuda::std::numeric_limits<float>::infinity();
  const int nrows = (y1 - y0 + y_step - 1) / y_step;   // rows y0, y0+y_step, ... < y1
  // Work unit: a strip of kBatchStrip tiles of 8 x 4 pixels along a row of tiles, fetched from a global counter (the cost of a pixel
  // varies by orders of magnitude over the image); lane = (ty, tx) of the tile.  A tile takes its limits from the tile before it
  // in the strip (the lane of the same row in its last column); the first tile of a strip has no neighbour and is seeded.
  constexpr int kTileW = 8, kTileH = 4, kStrip = kBatchStrip;
  const int tx = lane & 7, ty = lane >> 3;
  const int tiles_x = (width + kTileW - 1) / kTileW, strips_x = (tiles_x + kStrip - 1) / kStrip, tiles_y = (nrows + kTileH - 1) / kTileH;
  const long long units = (long long)strips_x * tiles_y;
  // one query kind (march step 0..9, wall point = 10) for all 32 pixels of the tile at a time
  auto run_kind = [&](const TreeView &tv, const float4 *__restrict__ power, int kind, bool first_tile, bool active, v3 q, int volume) -> v3 {
    const float prev_r2 = __shfl_sync(0xffffffffu, sm.hint[kind][lane], ty * 8 + 7);
    const float4 e = batch_estimate<KL>(tv, sm, active, q.x, q.y, q.z, k, max_r2, first_tile ? inf : prev_r2, power, volume, lane);
    __syncwarp();
    sm.hint[kind][lane] = active ? e.w : inf;
    __syncwarp();
    return V(e.x, e.y, e.z);
  };
  for (;;) {
    long long u = 0;
    if (lane == 0) u = (long long)atomicAdd(work_counter, 1ull);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= units) break;
    const int tile_row = (int)(u / strips_x), strip = (int)(u % strips_x);
    const int row = tile_row * kTileH + ty;
    for (int tile = 0; tile < kStrip; tile++) {
      if ((strip * kStrip + tile) * kTileW >= width) break;
#ifdef PM_KNN_BSTATS
      const long long t_tile = clock64();
#endif
      const int px = (strip * kStrip + tile) * kTileW + tx;
      const bool in_frame = px < width && row < nrows;
      const int py = y0 + row * y_step;
      const long long pix = (long long)py * width + px;
      const float x = (float)px + sc.cam_ox, y = (float)py + sc.cam_oy;
      v3 rgb = V(0.0f, 0.0f, 0.0f);
      const v3 origin = V(0.0f, 0.0f, 0.0f);
      v3 ray = V((float)((double)__fdiv_rn(x, sc.sz_img) - 0.5), (float)(-((double)__fdiv_rn(y, sc.sz_img) - 0.5)), 1.0f);
      if (media) {
        v3 prev = origin;
#pragma unroll 1
        for (int i = 0; i < 10; i++) {
          prev = add(mul(ray, 0.6f), prev);
          const v3 e = run_kind(tvv, pow_v, i, tile == 0, in_frame, prev, 1);
          rgb = add(rgb, mul(e, w_vol));
        }
      }
      Hit h; h.hit = 0; h.type = 0; h.idx = 0; h.dist = -1.0f;
      raytrace(sc, ray, origin, h);
      bool wall = false;
      v3 P = V(0.0f, 0.0f, 0.0f);
      if (h.hit) {
        P = mul(ray, h.dist);
        if (h.type == 0 && h.idx == 1) follow_specular(sc, ray, origin, h, P, 1);
        else if (h.type == 0 && h.idx == 0) follow_specular(sc, ray, origin, h, P, 0);
        wall = h.hit && h.type == 1;
      }
      {
        const bool active = in_frame && wall;
        const v3 e = run_kind(tvs, pow_s, 10, tile == 0, active, P, 0);
        if (active) {
          const v3 c = mul(e, w_surf);
          rgb = media ? add(rgb, mul(c, 0.15f)) : add(rgb, c);
        }
      }
      if (in_frame) {
        if (rgbf) rgbf[pix] = make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
        if (rgba) rgba[pix] = make_uchar4(quantise_u8(rgb.x), quantise_u8(rgb.y), quantise_u8(rgb.z), 0);
      }
#ifdef PM_KNN_BSTATS
      BSTAT(14, clock64() - t_tile);
#endif
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side of the build / query
// ---------------------------------------------------------------------------------------------------------
static cudaError_t ensure(void **p, size_t *cap, size_t bytes) {
  if (bytes <= *cap) return cudaSuccess;
  cudaFree(*p); *p = nullptr; *cap = 0;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaSuccess) *cap = bytes;
  return e;
}

#define KCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t knn_build(KnnMap &m, const float4 *pos, const float4 *power, long long n, int filter, int curve, cudaStream_t st, int *launches) {
  m.n = 0; m.levels = 0; m.power = power; m.src_pos = pos;
  if (n <= 0) return cudaSuccess;
  const long long tiles = (n + kSortTile - 1) / kSortTile, n_pad = tiles * kSortTile;
  KCK(ensure((void **)&m.keys[0], &m.cap_keys[0], sizeof(uint32_t) * n_pad));
  KCK(ensure((void **)&m.keys[1], &m.cap_keys[1], sizeof(uint32_t) * n_pad));
  KCK(ensure((void **)&m.vals[0], &m.cap_vals[0], sizeof(uint32_t) * n_pad));
  KCK(ensure((void **)&m.vals[1], &m.cap_vals[1], sizeof(uint32_t) * n_pad));
  KCK(ensure((void **)&m.ghist, &m.cap_ghist, sizeof(uint32_t) * 256 * tiles));
  KCK(ensure((void **)&m.scan_sums, &m.cap_scan_sums, sizeof(uint32_t) * ((256 * tiles + kScanChunk - 1) / kScanChunk + 1)));
  KCK(ensure((void **)&m.spos, &m.cap_spos, sizeof(float4) * (n + 2048)));   // whole (padded) leaves: rows past the last photon are NaN
  if (!m.d_count) KCK(cudaMalloc(&m.d_count, sizeof(unsigned long long)));
  KCK(cudaMemsetAsync(m.d_count, 0, sizeof(unsigned long long), st));
  morton_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(pos, n, n_pad, filter, curve, m.keys[0], m.vals[0], m.d_count);
  (*launches)++;
  int cur = 0;
  for (int pass = 0; pass < 4; pass++) {
    radix_hist_kernel<<<(unsigned)tiles, kSortThreads, 0, st>>>(m.keys[cur], pass * 8, (uint32_t)tiles, m.ghist);
    {
      size_t cnt = (size_t)256 * tiles;
      int chunks = (int)((cnt + kScanChunk - 1) / kScanChunk);
      scan_sums_kernel<<<chunks, kScanThreads, 0, st>>>(m.ghist, cnt, m.scan_sums);
      scan_top_kernel<<<1, 1024, 0, st>>>(m.scan_sums, chunks);
      scan_apply_kernel<<<chunks, kScanThreads, 0, st>>>(m.ghist, cnt, m.scan_sums);
    }
    radix_scatter_kernel<<<(unsigned)tiles, kSortThreads, 0, st>>>(m.keys[cur], m.vals[cur], m.keys[cur ^ 1], m.vals[cur ^ 1], pass * 8,
                                                                    (uint32_t)tiles, m.ghist);
    cur ^= 1;
    *launches += 5;
  }
  m.sorted = cur;
  KCK(cudaGetLastError());
  unsigned long long nv = (unsigned long long)n;   // an unfiltered build keeps every row: nothing to wait for
  if (filter) {
    KCK(cudaMemcpyAsync(&nv, m.d_count, sizeof(nv), cudaMemcpyDeviceToHost, st));
    KCK(cudaStreamSynchronize(st));   // the number of kept points sizes the tree
  }
  m.n = (long long)nv; m.n_sorted_pad = n_pad;
  if (m.n == 0) return cudaSuccess;
  // level geometry
  long long cnt = (m.n + 31) / 32; int L = 0; size_t total = 0;
  for (;;) {
    m.cnt[L] = cnt; m.pad[L] = (cnt + 31) / 32 * 32; m.off[L] = total; total += 6 * (size_t)m.pad[L];
    L++;
    if (cnt <= 32) break;
    cnt = (cnt + 31) / 32;
  }
  m.levels = L;
  KCK(ensure((void **)&m.boxes, &m.cap_boxes, sizeof(float) * total));
  permute_kernel<<<(unsigned)((m.pad[0] * 32 + 255) / 256), 256, 0, st>>>(pos, m.vals[cur], m.n, m.spos, m.cnt[0], m.pad[0], m.boxes + m.off[0]);
  (*launches)++;
  for (int l = 1; l < L; l++) {
    node_box_kernel<<<(unsigned)((m.pad[l] * 32 + 255) / 256), 256, 0, st>>>(m.boxes + m.off[l - 1], m.cnt[l - 1], m.pad[l - 1], m.cnt[l],
                                                                             m.pad[l], m.boxes + m.off[l]);
    (*launches)++;
  }
  return cudaGetLastError();
}

static TreeView make_view(const KnnMap &m) {
  TreeView tv;
  memset(&tv, 0, sizeof(tv));
  tv.spos = m.spos; tv.n = m.n; tv.levels = m.levels;
  for (int l = 0; l < m.levels; l++) { tv.cnt[l] = m.cnt[l]; tv.pad[l] = m.pad[l]; tv.box[l] = m.boxes + m.off[l]; }
  // stage the top two levels (they are contiguous at the end of the box allocation) if they fit
  tv.staged_from = m.levels; tv.staged_floats = 0;
  for (int l = m.levels - 1; l >= 0 && l >= m.levels - 2; l--) {
    long long f = tv.staged_floats + 6 * m.pad[l];
    if (f > kMaxStagedFloats) break;
    tv.staged_floats = f; tv.staged_from = l;
  }
  {
    int off = 0;
    for (int l = tv.staged_from; l < m.levels; l++) { tv.staged_off[l] = off; off += 6 * (int)m.pad[l]; }
  }
  return tv;
}

cudaError_t knn_query(const KnnMap &m, const float4 *queries, long long nq, int k, float max_r2, int32_t *idx, float *d2, int32_t *cnt,
                      int volume, float4 *rgb, int num_sms, cudaStream_t st, const float4 *cone_pos_meta, const float4 *cone_dir,
                      const float *cone_normals15, float exposure) {
  if (nq <= 0) return cudaSuccess;
  TreeView tv = make_view(m);
  long long want = (nq + kQueryThreads / 32 - 1) / (kQueryThreads / 32), cap = (long long)num_sms * 16;
  unsigned grid = (unsigned)(want < cap ? want : cap);
  const float4 *pw = m.power;
  ConeArgs ca;
  memset(&ca, 0, sizeof(ca));
  if (cone_dir) {
    ca.pos_meta = cone_pos_meta; ca.dir = cone_dir; ca.exposure = exposure; ca.enabled = 1;
    memcpy(ca.normal, cone_normals15, sizeof(ca.normal));
  }
  if (k <= 32) knn_query_kernel<1><<<grid, kQueryThreads, 0, st>>>(tv, queries, nq, k, max_r2, idx, d2, cnt, pw, volume, rgb, ca);
  else if (k <= 64) knn_query_kernel<2><<<grid, kQueryThreads, 0, st>>>(tv, queries, nq, k, max_r2, idx, d2, cnt, pw, volume, rgb, ca);
  else knn_query_kernel<4><<<grid, kQueryThreads, 0, st>>>(tv, queries, nq, k, max_r2, idx, d2, cnt, pw, volume, rgb, ca);
  return cudaGetLastError();
}

static TreeView make_view_unstaged(const KnnMap &m) {
  TreeView tv = make_view(m);
  tv.staged_from = m.levels; tv.staged_floats = 0;   // every level is read through L1 / L2
  return tv;
}

cudaError_t knn_render(const DeviceScene &sc, const KnnMap &ms, const KnnMap &mv, int k, float max_r2, float w_surf, float w_vol, int width,
                       int height, int y0, int y1, int y_step, bool media, unsigned long long *work_counter, uchar4 *rgba, float4 *rgbf, int num_sms,
                       cudaStream_t st, bool batched) {
  if (y_step < 1) y_step = 1;
  long long n = (long long)((y1 - y0 + y_step - 1) / y_step) * width;
  if (n <= 0) return cudaSuccess;
  if (batched) {   // one warp = one 8 x 4 pixel tile, one lane = one pixel (see "Batched search")
    TreeView tvs = make_view_unstaged(ms), tvv = make_view_unstaged(mv);
    KCK(cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), st));
    const int nrows = (y1 - y0 + y_step - 1) / y_step;
    const long long units = (long long)(((width + 7) / 8 + kBatchStrip - 1) / kBatchStrip) * ((nrows + 3) / 4);
#define LAUNCH_BATCHED(KL)                                                                                                      \
    do {                                                                                                                        \
      const size_t smem = sizeof(BatchShared) * kBatchWarps;                                                                    \
      KCK(cudaFuncSetAttribute(knn_render_batched_kernel<KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
      long long want = (units + kBatchWarps - 1) / kBatchWarps;                                                                 \
      unsigned grid = (unsigned)(want < (long long)num_sms ? want : (long long)num_sms);                                        \
      knn_render_batched_kernel<KL><<<grid, kBatchWarps * 32, smem, st>>>(sc, tvs, tvv, ms.power, mv.power, k, max_r2, w_surf, w_vol, width,   \
                                                                           height, y0, y1, y_step, media ? 1 : 0, rgba, rgbf, work_counter);    \
    } while (0)
    if (k <= 32) LAUNCH_BATCHED(1); else if (k <= 64) LAUNCH_BATCHED(2); else LAUNCH_BATCHED(4);
#undef LAUNCH_BATCHED
    return cudaGetLastError();
  }
  TreeView tvs = make_view(ms), tvv = make_view(mv);
  long long want = (long long)((width + 15) / 16) * (((y1 - y0 + y_step - 1) / y_step + 7) / 8), cap = (long long)num_sms * 16;   // 8x16-pixel tiles
  unsigned grid = (unsigned)(want < cap ? want : cap);
  long long staged = tvs.staged_floats > tvv.staged_floats ? tvs.staged_floats : tvv.staged_floats;
#ifdef PM_KNN_STATIC_SMEM
  staged = kMaxStagedFloats;   // the A/B baseline: room for the largest staged levels whatever the trees need (54.8 KB per CTA)
#endif
  const int stride = (int)((staged + 31) / 32 * 32 > 32 ? (staged + 31) / 32 * 32 : 32);   // 128-byte multiples (bulk-copy alignment)
  size_t smem = sizeof(float) * 2 * (size_t)stride + sizeof(u64) * (kQueryThreads / 32) * 64 + 2 * sizeof(unsigned long long);
  KCK(cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), st));
#define LAUNCH_RENDER(KL)                                                                                                  \
  do {                                                                                                                     \
    KCK(cudaFuncSetAttribute(knn_render_kernel<KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    knn_render_kernel<KL><<<grid, kQueryThreads, smem, st>>>(sc, tvs, tvv, ms.power, mv.power, k, max_r2, w_surf, w_vol, width, height, y0, \
                                                              y1, y_step, media ? 1 : 0, rgba, rgbf, work_counter, stride);                       \
  } while (0)
  if (k <= 32) LAUNCH_RENDER(1); else if (k <= 64) LAUNCH_RENDER(2); else LAUNCH_RENDER(4);
#undef LAUNCH_RENDER
  return cudaGetLastError();
}

#ifdef PM_KNN_BSTATS
extern "C" void pm_debug_knn_bstats(unsigned long long *out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_bstats, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_bstats, z, sizeof(z)); }
}
#endif

#ifdef PM_KNN_STATS
extern "C" void pm_debug_knn_stats(unsigned long long *out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_knn_stats, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_knn_stats, z, sizeof(z)); }
}
#endif

void knn_free(KnnMap &m) {
  cudaFree(m.keys[0]); cudaFree(m.keys[1]); cudaFree(m.vals[0]); cudaFree(m.vals[1]); cudaFree(m.ghist); cudaFree(m.scan_sums); cudaFree(m.spos);
  cudaFree(m.boxes); cudaFree(m.d_count);
  m = KnnMap();
}

}  // namespace pm
