/* include/pmb200.h -- the C-ABI of libpmb200.so, the B200-native drop-in for the CUDA hot path of
 * gamalik/cuda-photon-mapper (PMK = /root/reference/photonMappingKernel.cu).
 *
 * Two layers:
 *   (1) the reference's three extern "C" launchers, same names, argument meaning and error behaviour
 *       (PMK:1523, :1549, :1569; declared by their caller at simplePBO.cpp:27-29).  A host program that
 *       links photonMappingKernel.cu can link libpmb200.so instead.
 *   (2) an extended, handle-based API (pm_*) that exposes what the reference hides in __device__ globals
 *       and #defines (scene, photon count, RNG table, photon map), adds a headless float framebuffer,
 *       photon-range / row-band sharding for multi-GPU, and the k-NN photon-map pipeline (Mode B).
 *
 * Plain pointers and sizes only; no CUDA or torch types in any signature.  Pointers named dev_* are device
 * pointers owned by the caller; host_* are host pointers.  All pm_* calls return 0 on success or a negative
 * pm_status; pm_last_error() gives the text.  pm_* launches are asynchronous on the context's stream unless
 * the name ends in _host or says otherwise.  There is no CPU fallback: without a CUDA device pm_create fails.
 * Every pm_* call makes its context's device the calling thread's current CUDA device (cudaSetDevice) and leaves it so.
 */
#ifndef PMB200_H
#define PMB200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include "pmb200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* layout-compatible with CUDA's uchar4 (the mapped GL pixel-buffer object of simplePBO.cpp:137) */
typedef struct pm_uchar4 { unsigned char x, y, z, w; } pm_uchar4;

/* ------------------------------------------------------------------------------------------------
 * (1) Legacy drop-in symbols.  Process-global default context, default stream, fully synchronous; on a
 * CUDA error they print "Cuda error: <msg>: <str>." to stderr and exit(EXIT_FAILURE), as checkCUDAError
 * does (PMK:49-55).  Photon count = 10 000 (PMK:29) unless the environment variable PMB200_NR_PHOTONS is set.
 * ---------------------------------------------------------------------------------------------- */

/* replaces PMK:1569 launch_init_random_numbers_kernel: fills the random-direction table from the MWC
 * generator (seeds 6548/316), bit-identical to the reference's serial <<<1,1>>> loop, but in parallel. */
void launch_init_random_numbers_kernel(void);

/* replaces PMK:1523 launch_emit_photons_kernel: clear the photon map, trace all photons at animTime.
 * pos / image_width / image_height are ignored, as in the reference. */
void launch_emit_photons_kernel(pm_uchar4 *pos, unsigned int image_width, unsigned int image_height,
                                float animTime, bool interpolateFlag, bool participatingMediaFlag);

/* replaces PMK:1549 launch_photon_mapping_kernel: render image_width x image_height pixels into the DEVICE
 * buffer pos ({r,g,b,0}, row 0 = top). */
void launch_photon_mapping_kernel(pm_uchar4 *pos, unsigned int image_width, unsigned int image_height,
                                  float animTime, bool interpolateFlag, bool participatingMediaFlag);

/* The two launchers of the older variant (kernelPBO.cu).  They are dead in the reference: the caller's declarations
 * (simplePBO.cpp:21-24) sit inside a comment block (:20-25) and every call site is commented out (:118, :158, :176).  Exported anyway
 * (round 1) so that they would resolve if re-enabled.
 * launch_render_kernel (kernelPBO.cu:295-313 + render_kernel :268-291): uploads a HOST array of image_width*image_height
 * float3 pixels and writes pos[i] = {(unsigned char)r, (unsigned char)g, (unsigned char)b, 0} into the DEVICE buffer pos --
 * no scaling, conversion as nvcc compiles it (cvt.rzi.u32.f32, low byte: negatives and NaN give 0, 300.0f gives 44).
 * launch_kernel (kernelPBO.cu:317-359): its body is commented out in the reference; what remains -- synchronise, check
 * for errors, leave pos untouched -- is what this does.  numPhotons / photons are never read. */
void launch_render_kernel(pm_uchar4 *pos, unsigned int image_width, unsigned int image_height, float time,
                          const float *pixelData /* float3[image_width * image_height], host */);
void launch_kernel(pm_uchar4 *pos, unsigned int image_width, unsigned int image_height, float time,
                   void *numPhotons /* int [][5] */, void *photons /* float *[2][5][5000][3] */);

/* ------------------------------------------------------------------------------------------------
 * (2) Extended API
 * ---------------------------------------------------------------------------------------------- */
typedef struct pm_context pm_context;

typedef enum pm_status {
  PM_OK = 0,
  PM_ERR_CUDA = -1,        /* a CUDA runtime call failed (text in pm_last_error) */
  PM_ERR_ARG = -2,         /* invalid argument */
  PM_ERR_STATE = -3,       /* call order violated (e.g. trace before a table exists) */
  PM_ERR_NO_DEVICE = -4    /* no CUDA device: there is no CPU fallback */
} pm_status;

int         pm_create(pm_context **out, int device /* -1: current device */);
int         pm_destroy(pm_context *ctx);
const char *pm_last_error(const pm_context *ctx);
const char *pm_version(void);
pm_context *pm_default_context(void);                 /* the context behind the legacy symbols */
int         pm_set_stream(pm_context *ctx, void *cuda_stream /* cudaStream_t; NULL = default stream */);
int         pm_sync(pm_context *ctx);

/* scene / parameters (reference: __device__ globals PMK:59-73, #define nrPhotons PMK:29) */
void pm_scene_default(pm_scene *scene);
int  pm_set_scene(pm_context *ctx, const pm_scene *scene);
int  pm_get_scene(const pm_context *ctx, pm_scene *scene);
int  pm_position_objects(const pm_scene *in, float animTime, pm_scene *out);   /* PMK:1380-1404, host side */
/* What a Mode A trace of this scene at animTime will do (host side, no device needed): *two_phase = 1 when the scene meets the conditions
 * of the lock-step surface path (reference object layout, everything within [-8, 8], light >= 0.05 from every wall plane and outside the
 * spheres; otherwise every photon goes through the general state machine -- same results); shadow_need5[w] bit i = the shadow ray behind
 * wall w still has to test sphere i (0: the sphere lies wholly on the light's side of the wall, with margin, and is skipped) */
int  pm_trace_plan(const pm_scene *in, float animTime, int32_t *two_phase, uint32_t *shadow_need5);
int  pm_set_photon_count(pm_context *ctx, int64_t n_photons);                  /* table size, nrPhotons */
int  pm_set_photon_range(pm_context *ctx, int64_t first, int64_t last);        /* this GPU traces [first,last) */
int  pm_set_energy_scale(pm_context *ctx, float scale);   /* photon map is multiplied by this when built (1 = reference) */

/* random-direction table T2 (PMK:80) */
int pm_init_random_table(pm_context *ctx);       /* MWC stream, == launch_init_random_numbers_kernel.  Generates the rows of the
                                                    current photon range (+ rows 0..2); the stream advances by 3*n_photons */
/* counter-based alternative: row i = Philox4x32-10(counter i, key seed) mapped like randFloat(1.0); the MWC state (medium
 * draws) is untouched.  The oracle consumes the same table (oracle/pm_oracle.c pmo_philox_table). */
int pm_init_random_table_philox(pm_context *ctx, uint64_t seed);
int pm_set_random_table_host(pm_context *ctx, const float *host_xyz, int64_t n);
int pm_get_random_table_host(pm_context *ctx, float *host_xyz, int64_t n);
int pm_set_mwc_state(pm_context *ctx, uint32_t w, uint32_t z);   /* where the medium-scatter draws continue from */
int pm_get_mwc_state(const pm_context *ctx, uint32_t *w, uint32_t *z);

/* stage 1: photon emission + tracing (PMK:1215-1375 emitPhotons; F9-F17 of SURVEY.md 8(a)) */
#define PM_TRACE_MEDIA    1u   /* participatingMediaFlag */
#define PM_TRACE_RECORDS  2u   /* also append photon records to the SoA buffers (Mode B input) */
#define PM_TRACE_NO_MAP   4u   /* skip the voxel-map accumulation (records only) */
#define PM_TRACE_SPLIT    8u   /* run the medium walk and the surface walk as two launches instead of the fused,
                                  warp-specialised one (same results; kept for measurement and as a cross-check) */
#define PM_TRACE_EXACT_MEDIUM 16u /* medium walk: every deposit point with the reference's exact arithmetic.  By default a Mode A trace
                                  (counts only) computes steps 1-2 approximately and redoes, exactly, the photons that come near a voxel
                                  boundary (csrc/pm_trace.cu volume_photon_fast) -- same counts; kept as the cross-check */
#define PM_TRACE_ONE_PHASE 32u    /* surface walk: every photon through the general state machine.  By default a Mode A trace of a scene
                                  with the reference's object layout takes fresh photons through the common path in lock-step and hands
                                  the rest to the machine through per-warp queues (csrc/pm_trace.cu) -- same deposits; the cross-check */
int pm_clear_map(pm_context *ctx);                                 /* init_photons_kernel, PMK:1503-1521 */
int pm_trace(pm_context *ctx, float animTime, unsigned flags);
/* tuning: how many of a CTA's 32 warps run the medium walk in the fused trace kernel (default 7; 1..16) */
int pm_set_volume_warps(pm_context *ctx, int warps);
/* tuning: CTAs of the persistent trace kernel (default 0 = one per SM).  Fewer leaves SMs free for the previous frames' exchange +
 * map build and render, which the pipelined frame calls run on two further streams (worth it when those are a large part of the
 * frame, i.e. at 2-8 GPUs: 140 of 148 at 2, 132 at 4 and 8 measured best) */
int pm_set_trace_sms(pm_context *ctx, int sms);
/* the exact (int64 fixed-point) accumulators of the CURRENT frame (they rotate through three buffers, one per pm_clear_map).  Multi-GPU:
 * connect the ranks (pm_peer_* / pm_group_*, below) and pm_build_map sums them itself over NVLink; summing them by hand between
 * pm_trace and pm_build_map (e.g. ncclAllReduce, ncclInt64 / ncclSum -- round 1's path) still works when no peers are connected */
int pm_accumulators(pm_context *ctx, void **dev_ptr, size_t *n_int64);
int pm_get_accumulators_host(pm_context *ctx, int64_t *host_out /* n_int64 entries */);
int pm_build_map(pm_context *ctx);                                 /* accumulators -> float photon map + gather tables */
int pm_get_map_host(pm_context *ctx, float *host_grid /* 32*32*32*3 */);
int pm_set_map_host(pm_context *ctx, const float *host_grid);      /* inject a photon map (parity tests) */
int pm_map_device(pm_context *ctx, float **dev_grid);

/* photon records (valid after pm_trace with PM_TRACE_RECORDS).  Two SoA float4 buffer sets: PM_MAP_SURFACE holds the
 * storePhoton calls (surface + shadow photons, appended with warp-aggregated atomics, capacity = max_records),
 * PM_MAP_VOLUME the storeVolumePhoton calls (fixed slot 3*(index-first)+step, sized by the library). */
#define PM_MAP_SURFACE 0
#define PM_MAP_VOLUME  1
int pm_set_record_capacity(pm_context *ctx, int64_t max_records);
int pm_record_count(pm_context *ctx, int64_t *n);                  /* both sets; synchronises */
int pm_get_records_host(pm_context *ctx, pm_record *host_out, int64_t max_records);   /* canonical (index, call) order */
int pm_record_buffers(pm_context *ctx, int which, float **dev_pos_meta /* float4[] */, float **dev_power_index /* float4[] */,
                      float **dev_dir /* float4[] or NULL */, int64_t *count /* synchronises */);

/* stage 2 (Mode B): photon-map build = Morton keys + radix sort + implicit 32-wide LBVH, and stage 4: k-nearest-photon
 * search.  The reference has no counterpart; the definition is oracle/knn_oracle.c (brute force, (d2, index) order).
 * pm_knn_build uses the record buffers of the last PM_TRACE_RECORDS trace (surface map: wall hits only);
 * pm_knn_build_points builds over any DEVICE point set (float4 xyz+ignored, float4 power rgb+ignored; the arrays must
 * stay alive while the map is used) -- e.g. all-gathered records of several GPUs. */
#define PM_CURVE_MORTON  0   /* 30-bit Morton (Z-order) key of the 10-bit cell coordinates */
#define PM_CURVE_HILBERT 1   /* the same cell along the 3-D Hilbert curve (default: tighter leaf / node boxes) */
int pm_knn_set_curve(pm_context *ctx, int curve);
/* Mode B frames (pm_render_knn*): false (default) = one warp per pixel; true = one warp per 8 x 4 pixel tile, one lane per pixel, the
 * tree walked once per tile (experimental: same k-nearest sets, currently slower -- DESIGN.md 8). */
int pm_knn_set_batched(pm_context *ctx, bool on);
int pm_knn_build(pm_context *ctx, int which);
int pm_knn_build_points(pm_context *ctx, int which, const float *dev_pos4, const float *dev_power4, int64_t n,
                        bool records /* true: rows are photon records (w = meta); the surface map keeps wall hits only */);
int pm_knn_size(pm_context *ctx, int which, int64_t *n_points, int32_t *n_levels);
/* nq DEVICE queries (float4, w ignored); per query up to k photons with d2 <= max_r2 (INFINITY: pure k-NN), ascending
 * (d2, index): dev_idx[nq*k] (original indices, -1 = none), dev_d2[nq*k], dev_cnt[nq].  k <= 128. */
int pm_knn_query(pm_context *ctx, int which, const float *dev_queries4, int64_t nq, int k, float max_r2,
                 int32_t *dev_idx, float *dev_d2, int32_t *dev_cnt);
/* radiance estimate per query: dev_rgb4[q] = (sum of the found powers / (pi r_k^2) [surface] or (4/3 pi r_k^3) [volume], r_k^2) */
int pm_knn_radiance(pm_context *ctx, int which, const float *dev_queries4, int64_t nq, int k, float max_r2, float *dev_rgb4);
/* the legacy estimator (the only point-based one the reference's author wrote, "photonMappingKernel - Copy.cu":191-208):
 * fixed squared radius (0.7 there), cone filter, per-object photon lists.  Queries: float4 (x, y, z, wall id 0..4);
 * dev_rgb4[q] = (sum over the <= k nearest surface photons within sq_radius that hit the same wall of
 * power * max(0, -N.dir) * (1 - sqrt(d2)) / exposure, number of contributing photons).  Needs pm_knn_build(PM_MAP_SURFACE). */
int pm_knn_radiance_cone(pm_context *ctx, const float *dev_queries4, int64_t nq, int k, float sq_radius, float exposure, float *dev_rgb4);
/* stages 3+4+5, Mode B: rows [y0,y1) of a frame whose wall term is a k-NN estimate in the surface map and whose ten
 * ray-march terms are k-NN estimates in the volume map (media only), weighted by w_surface / w_volume and composited
 * like the reference (media: march sum + 0.15 * wall term).  Both maps must have been built with powers. */
int pm_render_knn(pm_context *ctx, float animTime, bool participatingMediaFlag, int width, int height, int y0, int y1, int k,
                  float max_r2, float w_surface, float w_volume, pm_uchar4 *dev_rgba, float *dev_rgbf);
/* rows y0, y0+y_step, y0+2*y_step, ... < y1 only: with y0 = rank and y_step = number of GPUs the (strongly position
 * dependent) gather cost is balanced across GPUs */
int pm_render_knn_rows(pm_context *ctx, float animTime, bool participatingMediaFlag, int width, int height, int y0, int y1, int y_step,
                       int k, float max_r2, float w_surface, float w_volume, pm_uchar4 *dev_rgba, float *dev_rgbf);
/* the same through HOST buffers (context-owned device frame buffers, synchronous copy-back) */
int pm_render_knn_host(pm_context *ctx, float animTime, bool participatingMediaFlag, int width, int height, int k, float max_r2,
                       float w_surface, float w_volume, pm_uchar4 *host_rgba, float *host_rgbf);
/* build products for the parity tests: sorted Morton keys + permutation (host), box arrays of one level (host) */
int pm_knn_sorted_host(pm_context *ctx, int which, uint32_t *host_keys, uint32_t *host_perm, int64_t n);
int pm_knn_level_host(pm_context *ctx, int which, int level, int64_t *count, float *host_boxes6 /* [6][count] or NULL */);

/* stages 3-5: eye rays, photon-map gather, volumetric ray-march (PMK:926-1017, :1409-1462).
 * Renders rows [y0,y1) of a width x height frame.  dev_rgba (uchar4, may be NULL) and dev_rgbf (float4 =
 * pre-quantisation rgb + 1, may be NULL) are whole-frame DEVICE buffers. */
int pm_render(pm_context *ctx, float animTime, bool interpolateFlag, bool participatingMediaFlag,
              int width, int height, int y0, int y1, pm_uchar4 *dev_rgba, float *dev_rgbf);
/* the same through HOST buffers: device frame buffers are owned by the context, results are copied back
 * (synchronous). */
int pm_render_host(pm_context *ctx, float animTime, bool interpolateFlag, bool participatingMediaFlag,
                   int width, int height, pm_uchar4 *host_rgba, float *host_rgbf);
/* one whole reference frame through host buffers: (optional emit) + render + copy-back; what display() does
 * (callbacksPBO.cpp:47-101) minus OpenGL */
int pm_frame_host(pm_context *ctx, float animTime, bool emitFlag, bool interpolateFlag, bool participatingMediaFlag,
                  int width, int height, pm_uchar4 *host_rgba, float *host_rgbf);
/* the same frame, pipelined: everything is enqueued and the call returns at once with a ticket.  The uchar4 frame (the
 * reference's out_data) is copied to host_rgba -- pinned memory, or the copy is not asynchronous -- on the context's
 * copy stream from a ring of three device frame buffers, so the copy of frame f runs under the trace of frame f+1.
 * pm_frame_wait blocks until the frame of that ticket is in host memory; only the three most recent tickets are valid
 * (three device frame buffers), so wait for ticket f-2 at the latest before submitting frame f+1 -- and give every frame in
 * flight its own host buffer. */
int pm_frame_host_async(pm_context *ctx, float animTime, bool emitFlag, bool interpolateFlag, bool participatingMediaFlag,
                        int width, int height, pm_uchar4 *host_rgba, int64_t *ticket);
int pm_frame_wait(pm_context *ctx, int64_t ticket);
/* the same pipeline with the frame left in DEVICE memory (dev_rgba / dev_rgbf: whole-frame buffers, either may be NULL; the
 * rows of pm_set_row_band are written).  Everything is enqueued: (emit: clear + trace) on the context's stream, exchange +
 * map build on a second stream, render (+ barrier) on a third, so up to three frames are in flight (the accumulators rotate
 * through three buffers and the gather tables through two for that).  With peers connected, dev_* may be another rank's
 * memory (pm_shared_open): every rank renders its band straight into rank 0's frame over NVLink, and a device-side barrier
 * follows -- when a rank's pm_sync returns, every band of that frame has landed. */
int pm_frame_device(pm_context *ctx, float animTime, bool emitFlag, bool interpolateFlag, bool participatingMediaFlag,
                    int width, int height, pm_uchar4 *dev_rgba, float *dev_rgbf);
/* allocate the context-owned device frame buffers of the *_host calls up front (they are otherwise sized on first use) */
int pm_reserve_frame(pm_context *ctx, int width, int height);

/* ------------------------------------------------------------------------------------------------
 * multi-GPU (the reference is single-device, simplePBO.cpp:191-195).  The path shards as: rank r traces the
 * photon range pm_set_photon_range gives it, the exact accumulators of all ranks are SUMMED (the one exchange), every rank
 * builds the same photon map and renders its own row band.  Once peers are connected pm_build_map performs the exchange
 * itself: one kernel pulls the other ranks' accumulators over NVLink peer memory, with the rank synchronisation inside it
 * (csrc/pm_peer.cu) -- no library collective.  All ranks must then call pm_clear_map / pm_trace / pm_build_map in lockstep,
 * frame by frame.  Integer sums: the map is bit-identical for every number of ranks.
 *   - ranks as PROCESSES (one per GPU, e.g. under torchrun): pm_peer_export on every rank, exchange the 64-byte handles by
 *     any means, pm_peer_connect with all of them (CUDA IPC);
 *   - ranks as contexts of ONE process: pm_peer_connect_local with the member contexts (peer access), or simply the
 *     pm_group_* calls below, which own the contexts and one host thread per GPU.
 * ---------------------------------------------------------------------------------------------- */
#define PM_PEER_HANDLE_BYTES 64
#define PM_MAX_RANKS 16
int pm_peer_export(pm_context *ctx, void *handle /* PM_PEER_HANDLE_BYTES */);
int pm_peer_connect(pm_context *ctx, int rank, int world, const void *handles /* world x PM_PEER_HANDLE_BYTES; entry [rank] unused */);
int pm_peer_connect_local(pm_context *ctx, int rank, int world, pm_context *const *members /* [world], members[rank] == ctx */);
int pm_peer_disconnect(pm_context *ctx);
int pm_peer_info(const pm_context *ctx, int *rank, int *world);
int pm_peer_barrier(pm_context *ctx);              /* device-side barrier of all ranks, enqueued on the stream */
int pm_peer_status(pm_context *ctx);               /* synchronises; PM_ERR_STATE if a peer wait timed out (default 5 s) */
int pm_peer_set_timeout(pm_context *ctx, double seconds);
/* device memory that another rank can map (CUDA IPC), e.g. the frame buffer every rank renders its band into */
int pm_shared_alloc(pm_context *ctx, size_t bytes, void **dev_ptr, void *handle /* PM_PEER_HANDLE_BYTES, or NULL */);
int pm_shared_open(pm_context *ctx, const void *handle, void **dev_ptr);
int pm_shared_close(pm_context *ctx, void *dev_ptr, bool opened /* true: from pm_shared_open; false: from pm_shared_alloc */);
/* the rows [y0, y1) this rank renders and copies in pm_render_host / pm_frame_host / pm_frame_host_async (the host pointer
 * stays the base of the WHOLE frame: every rank copies its own band over its own PCIe link); (-1, -1) = all rows */
int pm_set_row_band(pm_context *ctx, int y0, int y1);

/* The same as one object, for a single-process host program (the reference's model: one host thread calling display(),
 * callbacksPBO.cpp:47-101): n contexts on n devices, one worker thread per GPU inside the library, peers connected.
 * pm_group_frame_host* is display() for the group: (emit: clear + trace 1/n of the photons + exchange + map build) + every
 * GPU renders its row band and copies it into the caller's HOST frame.  devices may repeat a device (tests). */
typedef struct pm_group pm_group;
int         pm_group_create(pm_group **out, const int *devices, int n);
int         pm_group_destroy(pm_group *g);
int         pm_group_size(const pm_group *g);
pm_context *pm_group_context(pm_group *g, int rank);          /* for inspection (maps, accumulators, timings) */
const char *pm_group_last_error(const pm_group *g);
int pm_group_set_scene(pm_group *g, const pm_scene *scene);
int pm_group_set_photon_count(pm_group *g, int64_t n_photons);   /* also shards the photon range across the ranks */
int pm_group_set_energy_scale(pm_group *g, float scale);
int pm_group_init_random_table(pm_group *g);
int pm_group_frame_host(pm_group *g, float animTime, bool emitFlag, bool interpolateFlag, bool participatingMediaFlag,
                        int width, int height, pm_uchar4 *host_rgba, float *host_rgbf);
int pm_group_frame_host_async(pm_group *g, float animTime, bool emitFlag, bool interpolateFlag, bool participatingMediaFlag,
                              int width, int height, pm_uchar4 *host_rgba /* pinned */, int64_t *ticket);
int pm_group_frame_wait(pm_group *g, int64_t ticket);

/* instrumentation: number of kernels this context has launched so far, and optional per-kernel CUDA-event timing
 * (event pairs recorded on the context's stream around each launch; pm_get_timings synchronises, fills
 * total_ms[k] / launches[k] for k < pm_kernel_count() since the last call, and resets) */
int64_t     pm_launch_count(const pm_context *ctx);
int         pm_enable_timing(pm_context *ctx, bool on);
int         pm_kernel_count(void);
const char *pm_kernel_name(int kind);
int         pm_get_timings(pm_context *ctx, double *total_ms, int64_t *launches);
/* development aid: per-CTA %globaltimer stamps (ns) of the next trace launches -- 48 words per CTA: [0] start,
 * [1] shared accumulators zeroed, [2] every warp out of photons, [3] accumulators flushed, [8+w] warp w out of photons */
int         pm_trace_profile(pm_context *ctx, bool on);
int         pm_get_trace_profile_host(pm_context *ctx, uint64_t *host_out, int64_t max_words, int64_t *words);
/* self-test of the trace kernel's branch-free wall division (csrc/pm_math.cuh fdiv_fastpath) against the IEEE division on `pairs`
 * pseudo-random operand pairs of the domain the kernel argues about: *violations = pairs where an acceptable quotient
 * (0 < q < 999999.9) differs in any bit, or a rejected one comes out acceptable; *accepted = pairs with an acceptable quotient */
int         pm_selftest_fdiv(pm_context *ctx, uint64_t pairs, uint32_t seed, uint64_t *violations, uint64_t *accepted);

#ifdef __cplusplus
}
#endif
#endif
