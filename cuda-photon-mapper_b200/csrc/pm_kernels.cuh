// pm_kernels.cuh -- launcher declarations shared between the kernel TUs and the C-ABI host layer (pm_api.cu).
#pragma once

#include <cuda_runtime.h>
#include "../../include/pmb200.h"
#include "pm_layout.h"
#include "pm_math.cuh"

namespace pm {

// host-side MWC arithmetic: one MWC lane is x -> x * a mod (a*2^16 - 1)  (see pm_math.cuh MwcJump)
inline uint32_t host_mulmod(uint32_t a, uint32_t b, uint32_t m) { return (uint32_t)(((unsigned long long)a * b) % m); }
inline uint32_t host_powmod(uint32_t a, unsigned long long e, uint32_t m) {
  uint32_t r = 1;
  while (e) { if (e & 1) r = host_mulmod(r, a, m); a = host_mulmod(a, a, m); e >>= 1; }
  return r;
}

// pm_trace.cu
cudaError_t launch_mwc_table(float4 *table, long long first, long long last, long long n, uint32_t w0, uint32_t z0, const MwcJump *J, cudaStream_t st);
cudaError_t launch_philox_table(float4 *table, long long n, unsigned long long seed, cudaStream_t st);
cudaError_t launch_selftest_fdiv(unsigned long long n, uint32_t seed, unsigned long long *out, int num_sms, cudaStream_t st);
cudaError_t launch_table_norm(float4 *table, long long n, cudaStream_t st);   // w = 1/sqrt(x*x+y*y+z*z) of rows [0, n)
// each returns the number of kernels launched; *err receives the CUDA status
int launch_trace_volume(const DeviceScene &sc, const float4 *table, long long first, long long last, unsigned flags, uint32_t w0,
                        uint32_t z0, const MwcJump *J, unsigned long long *acc, uint32_t *vol_cnt, float4 *rec_pos, float4 *rec_pow,
                        float4 *rec_dir, unsigned long long *rec_count, long long rec_cap, int num_sms, cudaStream_t st,
                        cudaError_t *err);
int launch_trace(const DeviceScene &sc, const float4 *table, long long first, long long last, unsigned flags, int vol_warps, uint32_t w0,
                 uint32_t z0, const MwcJump *J, unsigned long long *acc, uint32_t *vol_cnt, float4 *rec_pos, float4 *rec_pow,
                 float4 *rec_dir, float4 *vrec_pos, float4 *vrec_pow, long long vrec_cap, unsigned long long *rec_count, long long rec_cap,
                 int num_sms, cudaStream_t st, cudaError_t *err, unsigned long long *dbg = nullptr, uint32_t *vox_touched = nullptr,
                 uint32_t *queue = nullptr);
// per-warp queues of trace_kernel's two-phase surface walk: two queues of kTraceQueueCap entries of 8 words, structure of arrays
constexpr int kTraceQueueCap = 128, kTraceQueueWords = 2 * 8 * kTraceQueueCap;
constexpr size_t kTraceQueueWordsPerCta = (size_t)32 * kTraceQueueWords;   // scratch the caller provides: this many words per SM
constexpr int kTraceDbgWords = 48;   // pm_trace_profile: per CTA [0] start, [1] accumulators zeroed, [2] all warps done, [3] flushed, [8+w] warp w done (ns)

// pm_map.cu
cudaError_t launch_build_map(const long long *acc, float energy_scale, float *grid, cudaStream_t st);
cudaError_t launch_build_tables(const float *grid, float4 *vol_table, float4 *surf_table, cudaStream_t st);

// Force-load the kernels of a TU (cudaFuncGetAttributes).  With CUDA's lazy module loading the FIRST launch of a kernel loads its
// code, which synchronises with work already running on the device: a rank spinning inside peer_reduce_kernel for a peer whose
// trace kernel is not loaded yet would wait forever (measured: the same-device group test timed out).  pm_peer_connect* preloads.
cudaError_t preload_trace_kernels();
cudaError_t preload_map_kernels();
cudaError_t preload_render_kernels();
cudaError_t preload_peer_kernels();

// pm_peer.cu -- multi-GPU exchange over peer memory
struct PeerView {
  int world, rank;
  ExchangeHeader *hdr[kMaxPeers];      // every rank's exchange block (own entry: the local one)
  const long long *acc[kMaxPeers];     // = (long long *)(hdr[p] + 1): two buffers, kAccStride apart
  unsigned long long timeout_ns;
};
cudaError_t launch_peer_reduce(const PeerView &pv, int buf, uint32_t seq, long long *out, int blocks, cudaStream_t st);
cudaError_t launch_peer_barrier(const PeerView &pv, uint32_t seq, cudaStream_t st);

// pm_render.cu
cudaError_t launch_render(const DeviceScene &sc, const float4 *vol_table, const float4 *surf_table, int width, int height,
                          int y0, int y1, bool interp, bool media, uchar4 *rgba, float4 *rgbf, cudaStream_t st);

cudaError_t launch_present_float3(const float *rgb, long long pixels, uchar4 *pos, cudaStream_t st);

// pm_knn.cu -- Mode B photon map: sorted points + implicit 32-wide LBVH (see the file header)
struct KnnMap {
  long long n = 0;                 // points in the map (after filtering)
  int levels = 0;                  // box levels, level 0 = leaves of 32 points
  long long cnt[8] = {0}, pad[8] = {0};
  size_t off[8] = {0};             // float offset of each level's boxes (6 floats per entity) inside `boxes`
  float *boxes = nullptr; size_t cap_boxes = 0;
  uint32_t *keys[2] = {nullptr, nullptr}, *vals[2] = {nullptr, nullptr};
  size_t cap_keys[2] = {0, 0}, cap_vals[2] = {0, 0};
  uint32_t *ghist = nullptr; size_t cap_ghist = 0;
  uint32_t *scan_sums = nullptr; size_t cap_scan_sums = 0;
  float4 *spos = nullptr; size_t cap_spos = 0;
  unsigned long long *d_count = nullptr;
  int sorted = 0;                  // which of keys[]/vals[] holds the sorted pairs
  long long n_sorted_pad = 0;
  const float4 *power = nullptr;   // rgb power per ORIGINAL index (not owned)
  const float4 *src_pos = nullptr; // positions per ORIGINAL index (not owned)
};
cudaError_t knn_build(KnnMap &m, const float4 *pos, const float4 *power, long long n, int filter, int curve, cudaStream_t st, int *launches);
cudaError_t knn_query(const KnnMap &m, const float4 *queries, long long nq, int k, float max_r2, int32_t *idx, float *d2, int32_t *cnt,
                      int volume, float4 *rgb, int num_sms, cudaStream_t st, const float4 *cone_pos_meta = nullptr,
                      const float4 *cone_dir = nullptr, const float *cone_normals15 = nullptr, float exposure = 1.0f);
cudaError_t knn_render(const DeviceScene &sc, const KnnMap &ms, const KnnMap &mv, int k, float max_r2, float w_surf, float w_vol, int width,
                       int height, int y0, int y1, int y_step, bool media, unsigned long long *work_counter, uchar4 *rgba, float4 *rgbf, int num_sms,
                       cudaStream_t st, bool batched = false);
void knn_free(KnnMap &m);

}  // namespace pm
