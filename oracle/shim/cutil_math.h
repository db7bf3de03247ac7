/* oracle/shim/cutil_math.h -- TEST INFRASTRUCTURE (not product code).
 *
 * Stand-in for NVIDIA GPU Computing SDK 3.x/4.x <cutil_math.h>, which the reference
 * includes (photonMappingKernel.cu:4) but does not vendor.  Only the operators the
 * reference's hot file actually uses are provided.  Pinned semantics (SURVEY.md 8(c),
 * "parity unpinned" by the reference itself -- these are OUR documented choices):
 *   - dot(a,b)      = a.x*b.x + a.y*b.y + a.z*b.z, left to right, FP32
 *   - normalize(v)  = v * rsqrtf(dot(v,v));  on the host rsqrtf(x) = 1.0f/sqrtf(x)
 *                     (what cutil's host path did); on the device it is the rsqrt.approx
 *                     intrinsic, as in the original SDK header
 *   - a / s         = a * (1.0f/s)   (SDK-era cutil_math multiplied by the reciprocal)
 * Used by: oracle/ref_host_harness.cpp (host build) and oracle/ref_cuda_harness.cu (sm_100a build).
 */
#ifndef PMB200_ORACLE_SHIM_CUTIL_MATH_H
#define PMB200_ORACLE_SHIM_CUTIL_MATH_H

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define PM_HD __host__ __device__
#else
#include "host_cuda_shim.h"
#define PM_HD
#endif

typedef unsigned int uint;

#ifndef __CUDACC__
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#endif

inline PM_HD float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline PM_HD float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline PM_HD float3 operator-(float3 a, float b)  { return make_float3(a.x - b, a.y - b, a.z - b); }
inline PM_HD float3 operator*(float3 a, float s)  { return make_float3(a.x * s, a.y * s, a.z * s); }
inline PM_HD float3 operator*(float s, float3 a)  { return make_float3(a.x * s, a.y * s, a.z * s); }
inline PM_HD float3 operator/(float3 a, float s)  { float inv = 1.0f / s; return a * inv; }
inline PM_HD void operator+=(float3 &a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
inline PM_HD float dot(float3 a, float3 b)        { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline PM_HD float3 normalize(float3 v)           { float invLen = rsqrtf(dot(v, v)); return v * invLen; }

#endif
