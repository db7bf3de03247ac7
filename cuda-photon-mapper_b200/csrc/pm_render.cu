// pm_render.cu -- stages 3+4+5: eye rays, wall gather, volumetric ray-march (sm_100a).
//
// Replaces photon_mapping_kernel / computePixelColor (PMK:1409-1462, :926-1017).  One thread per pixel.
// The per-pixel voxel loops of the reference (integrate: <=36 reads, x4 with interpolation; integrateVolumePhotons:
// <=216 reads x 10 march steps) are replaced by lookups in the tables pm_map.cu built in the reference's own
// summation order, so the float framebuffer is bit-identical to the sequential oracle's for the same photon map.
// Writes the reference's uchar4 {r,g,b,0} (row 0 = top) and/or a float4 framebuffer (pre-quantisation rgb, w = 1).
#include "pm_kernels.cuh"

namespace pm {

__device__ __forceinline__ v3 vol_lookup(const float4 *__restrict__ vol_table, v3 p) {
  int wx = voxel_x_wide(p.x) - kVolLo, wy = voxel_x_wide(p.y) - kVolLo, wz = voxel_z_wide(p.z) - kVolLo;   // exact in -2..34
  if ((unsigned)wx >= (unsigned)kVolN || (unsigned)wy >= (unsigned)kVolN || (unsigned)wz >= (unsigned)kVolN) return V(0.0f, 0.0f, 0.0f);
  float4 t = __ldg(vol_table + (wx * kVolN + wy) * kVolN + wz);
  return V(t.x, t.y, t.z);
}

// integrate() for the point p on wall `id` (PMK:314-389); spheres and unknown ids gather nothing
__device__ __forceinline__ v3 surf_lookup(const float4 *__restrict__ surf_table, v3 p, int type, int id) {
  if (type != 1 || (unsigned)id >= (unsigned)PM_MAX_PLANES) return V(0.0f, 0.0f, 0.0f);
  int wx = voxel_x_wide(p.x), wy = voxel_x_wide(p.y), wz = voxel_z_wide(p.z);
  // an empty window on the wall's own axis cannot happen (the slab index is forced), only in-plane coords matter
  int a = (id == 0 || id == 2) ? wy : wx;
  int b = (id == 4) ? wy : wz;
  a -= kSurfLo; b -= kSurfLo;
  if ((unsigned)a >= (unsigned)kSurfN || (unsigned)b >= (unsigned)kSurfN) return V(0.0f, 0.0f, 0.0f);
  float4 t = __ldg(surf_table + (id * kSurfN + a) * kSurfN + b);
  return V(t.x, t.y, t.z);
}

// centerPoint / getWorldCoordinates, PMK:392-400, :269-274 (double scale and shift)
__device__ __forceinline__ v3 center_point(v3 p) {
  int wx = voxel_x(p.x), wy = voxel_x(p.y), wz = voxel_z(p.z);
  v3 c;
  c.x = (float)((double)__fdiv_rn((float)wx, 32.0f) * 3.0 - 1.5);
  c.y = (float)((double)__fdiv_rn((float)wy, 32.0f) * 3.0 - 1.5);
  c.z = (float)((double)__fdiv_rn((float)wz, 32.0f) * 6.0);
  return add(c, V(0.046875f, 0.046875f, 0.09375f));
}
__device__ __forceinline__ float alpha_of(v3 p, v3 a, v3 b, int axis) {
  return __fdiv_rn(comp(p, axis) - comp(a, axis), comp(b, axis) - comp(a, axis));
}
__device__ __forceinline__ v3 lerp3(v3 a, v3 b, float alfa) {   // (1.0 - alfa) in double, narrowed: PMK:424-426
  return add(mul(a, (float)(1.0 - (double)alfa)), mul(b, alfa));
}

// interpolateEnergy + bilinearInterpolate, PMK:442-583
static __device__ __noinline__ v3 interpolate_energy(const float4 *__restrict__ surf_table, v3 p, int type, int id) {
  if (type != 1 || (unsigned)id >= (unsigned)PM_MAX_PLANES) return V(0.0f, 0.0f, 0.0f);
  int a1, a2;
  if (id == 0 || id == 2) { a1 = 2; a2 = 1; } else if (id == 1 || id == 3) { a1 = 0; a2 = 2; } else { a1 = 0; a2 = 1; }
  float d1 = (a1 == 2) ? 0.1875f : 0.09375f, d2 = (a2 == 2) ? 0.1875f : 0.09375f;
  v3 p1 = center_point(p), p2 = V(0.0f, 0.0f, 0.0f), p3 = p2, p4 = p2;
  float s1 = comp(p, a1) > comp(p1, a1) ? d1 : -d1;
  set_comp(p2, a1, comp(p1, a1) + s1); set_comp(p3, a1, comp(p1, a1) + s1); set_comp(p4, a1, comp(p1, a1));
  float s2 = comp(p, a2) > comp(p1, a2) ? d2 : -d2;
  set_comp(p2, a2, comp(p1, a2)); set_comp(p3, a2, comp(p1, a2) + s2); set_comp(p4, a2, comp(p1, a2) + s2);
  v3 c1 = surf_lookup(surf_table, p1, type, id), c2 = surf_lookup(surf_table, p2, type, id);
  v3 c3 = surf_lookup(surf_table, p3, type, id), c4 = surf_lookup(surf_table, p4, type, id);
  float alfa = alpha_of(p, p1, p2, a1);
  v3 p12 = lerp3(p1, p2, alfa), c12 = lerp3(c1, c2, alfa);
  alfa = alpha_of(p, p3, p4, a1);
  v3 p34 = lerp3(p3, p4, alfa), c34 = lerp3(c3, c4, alfa);
  alfa = alpha_of(p, p12, p34, a2);
  return lerp3(c12, c34, alfa);
}

// quantisation of photon_mapping_kernel, PMK:1451-1453, with the device's saturating cast (NaN, negatives -> 0)
__device__ __forceinline__ unsigned char quantise(float v) {
  double d = (double)v * 255.0;
  d = d > 255.0 ? 255.0 : d;
  return d > 0.0 ? (unsigned char)__double2uint_rz(d) : (unsigned char)0;
}

__global__ void __launch_bounds__(256) render_kernel(const __grid_constant__ DeviceScene sc, const float4 *__restrict__ vol_table,
                                                     const float4 *__restrict__ surf_table, int width, int height, int y0, int y1,
                                                     int interp, int media, uchar4 *__restrict__ rgba, float4 *__restrict__ rgbf) {
  // 32-bit index arithmetic (launch_render checks that the band has fewer than 2^32 pixels): the 64-bit % and / were 8 % of the samples
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)(y1 - y0) * (unsigned)width) return;
  const unsigned row = i / (unsigned)width;
  const int px = (int)(i - row * (unsigned)width), py = y0 + (int)row;
  const long long pix = (long long)py * width + px;
  float x = (float)px + sc.cam_ox, y = (float)py + sc.cam_oy;

  v3 rgb = V(0.0f, 0.0f, 0.0f);
  const v3 origin = V(0.0f, 0.0f, 0.0f);
  v3 ray = V((float)((double)__fdiv_rn(x, sc.sz_img) - 0.5), (float)(-((double)__fdiv_rn(y, sc.sz_img) - 0.5)), 1.0f);
  Hit h; h.hit = 0; h.type = 0; h.idx = 0; h.dist = -1.0f;
  if (media) {   // PMK:937-965: 10 steps of 0.6 along the unnormalised ray, raw box sums
    // unrolled: ray.z is the literal 1.0f, so the z coordinate of every march step -- and its voxel, and that voxel's range test --
    // is a compile-time constant (the same IEEE operations, folded by the compiler)
    v3 prev = origin;
#pragma unroll
    for (int i = 0; i < 10; i++) { prev = add(mul(ray, 0.6f), prev); rgb = add(rgb, vol_lookup(vol_table, prev)); }
  }
  raytrace(sc, ray, origin, h);
  if (h.hit) {
    v3 P = mul(ray, h.dist);
    if (h.type == 0 && h.idx == 1) follow_specular(sc, ray, origin, h, P, 1);
    else if (h.type == 0 && h.idx == 0) follow_specular(sc, ray, origin, h, P, 0);
    if (h.hit) {
      v3 c = interp ? interpolate_energy(surf_table, P, h.type, h.idx) : surf_lookup(surf_table, P, h.type, h.idx);
      rgb = media ? add(rgb, mul(c, 0.15f)) : add(rgb, c);
    }
  }
  if (rgbf) rgbf[pix] = make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
  if (rgba) rgba[pix] = make_uchar4(quantise(rgb.x), quantise(rgb.y), quantise(rgb.z), 0);
}

cudaError_t launch_render(const DeviceScene &sc, const float4 *vol_table, const float4 *surf_table, int width, int height,
                          int y0, int y1, bool interp, bool media, uchar4 *rgba, float4 *rgbf, cudaStream_t st) {
  long long n = (long long)(y1 - y0) * width;
  if (n <= 0) return cudaSuccess;
  if (n >= (1ll << 32) - 256) return cudaErrorInvalidValue;   // the kernel indexes the band's pixels with 32 bits
  unsigned blocks = (unsigned)((n + 255) / 256);
  render_kernel<<<blocks, 256, 0, st>>>(sc, vol_table, surf_table, width, height, y0, y1, interp ? 1 : 0, media ? 1 : 0, rgba, rgbf);
  return cudaGetLastError();
}

// render_kernel of the older variant (kernelPBO.cu:268-291): float3 pixels -> uchar4, unscaled, device conversion semantics
__global__ void __launch_bounds__(256) present_float3_kernel(const float *__restrict__ rgb, long long pixels, uchar4 *__restrict__ pos) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  uchar4 o;
  o.x = (unsigned char)__float2uint_rz(rgb[3 * i + 0]);
  o.y = (unsigned char)__float2uint_rz(rgb[3 * i + 1]);
  o.z = (unsigned char)__float2uint_rz(rgb[3 * i + 2]);
  o.w = 0;
  pos[i] = o;
}

cudaError_t preload_render_kernels() {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, render_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, present_float3_kernel);
  return e;
}

cudaError_t launch_present_float3(const float *rgb, long long pixels, uchar4 *pos, cudaStream_t st) {
  if (pixels <= 0) return cudaSuccess;
  present_float3_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, st>>>(rgb, pixels, pos);
  return cudaGetLastError();
}

}  // namespace pm
