/* oracle/ref_cuda_harness.cu -- TEST / BASELINE INFRASTRUCTURE (never linked into the product).
 *
 * Compiles the reference's OWN CUDA file (photonMappingKernel.cu, unmodified except the nrPhotons #define,
 * see oracle/build_ref.sh) for sm_100a and exposes handles to its __device__ globals, so that bench.py can
 * time "the reference CUDA kernel itself on one B200" (BASELINE.md 3, B-CUDA) and the GPU tests can compare
 * images statistically.  It is a SPEED baseline and a statistical check only: the reference kernels are racy
 * (non-atomic += on the grid, shared MWC state), compiled with FMA contraction and rsqrt.approx, hence not a
 * bit oracle.  Output: oracle/_ref/libpmref_cuda_<capacity>.so.
 *
 * nrPhotons is a compile-time constant of the reference: this library always traces PM_REF_CAPACITY photons.
 */
#include <cuda_runtime.h>
#include <stdio.h>

#ifdef PM_REF_ATOMIC   /* build_ref.sh P4: the deposit sites call this instead of the racy `+=` */
static __device__ __forceinline__ void pm_atomic_add3(float3 *p, float3 v) {
  atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); atomicAdd(&p->z, v.z);
}
#endif

#include PM_REF_STAGED   /* the whole reference file, streamed by build_ref.sh */

#define RCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "refcu: %s: %s\n", #x, cudaGetErrorString(e_)); return -1; } } while (0)

extern "C" {

int refcu_capacity(void) { return nrPhotons; }

int refcu_set_table(const float *host_xyz, int n) {
  if (n > nrPhotons) return -2;
  RCK(cudaMemcpyToSymbol(randomNumbers, host_xyz, sizeof(float) * 3 * (size_t)n));
  return 0;
}
int refcu_set_table_dev(const float *dev_xyz, int n) {
  if (n > nrPhotons) return -2;
  RCK(cudaMemcpyToSymbol(randomNumbers, dev_xyz, sizeof(float) * 3 * (size_t)n, 0, cudaMemcpyDeviceToDevice));
  return 0;
}
int refcu_set_szimg(int sz) { RCK(cudaMemcpyToSymbol(szImg, &sz, sizeof(int))); return 0; }
int refcu_set_scene(const int *nrObj, const float *pl, const float *sp, const float *light) {
  RCK(cudaMemcpyToSymbol(nrObjects, nrObj, sizeof(int) * 2));
  RCK(cudaMemcpyToSymbol(planes, pl, sizeof(float) * 10));
  RCK(cudaMemcpyToSymbol(spheres, sp, sizeof(float) * 12));
  RCK(cudaMemcpyToSymbol(Light, light, sizeof(float) * 3));
  return 0;
}
int refcu_get_grid(float *host) { RCK(cudaMemcpyFromSymbol(host, photons, sizeof(float) * 3 * NR_PHOTONS_X * NR_PHOTONS_Y * NR_PHOTONS_Z)); return 0; }
int refcu_set_grid(const float *host) { RCK(cudaMemcpyToSymbol(photons, host, sizeof(float) * 3 * NR_PHOTONS_X * NR_PHOTONS_Y * NR_PHOTONS_Z)); return 0; }

/* the reference's own three launchers (PMK:1523, :1549, :1569); each synchronises internally */
void refcu_init_random(void) { launch_init_random_numbers_kernel(); }
void refcu_emit(float t, int interp, int media) { launch_emit_photons_kernel(0, 0, 0, t, interp != 0, media != 0); }
void refcu_render(void *dev_rgba, unsigned w, unsigned h, float t, int interp, int media) {
  launch_photon_mapping_kernel((uchar4 *)dev_rgba, w, h, t, interp != 0, media != 0);
}

/* frame loop timed inside with CUDA events on the default stream (the stream the reference launches on) */
int refcu_time_frames(void *dev_rgba, unsigned w, unsigned h, float t, int interp, int media, int warmup, int steps,
                      float *emit_ms, float *render_ms) {
  cudaEvent_t e0, e1, e2;
  RCK(cudaEventCreate(&e0)); RCK(cudaEventCreate(&e1)); RCK(cudaEventCreate(&e2));
  double se = 0.0, sr = 0.0;
  for (int i = 0; i < warmup + steps; i++) {
    RCK(cudaEventRecord(e0, 0));
    launch_emit_photons_kernel(0, 0, 0, t, interp != 0, media != 0);
    RCK(cudaEventRecord(e1, 0));
    launch_photon_mapping_kernel((uchar4 *)dev_rgba, w, h, t, interp != 0, media != 0);
    RCK(cudaEventRecord(e2, 0));
    RCK(cudaEventSynchronize(e2));
    float a = 0, b = 0;
    RCK(cudaEventElapsedTime(&a, e0, e1)); RCK(cudaEventElapsedTime(&b, e1, e2));
    if (i >= warmup) { se += a; sr += b; }
  }
  *emit_ms = (float)(se / steps); *render_ms = (float)(sr / steps);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  return 0;
}

} /* extern "C" */
