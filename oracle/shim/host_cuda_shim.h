/* oracle/shim/host_cuda_shim.h -- TEST INFRASTRUCTURE (not product code).
 *
 * Lets the device routines of the reference's hot file (photonMappingKernel.cu:1-1521) compile
 * as ordinary host C++ so that they can serve as the sequential parity oracle and the OpenMP CPU
 * baseline.  Provides: empty __device__/__global__, CUDA vector
 * structs, thread_local block/thread ids, a no-op __threadfence, the mixed double/float min/max
 * overloads CUDA has and <algorithm> lacks (needed at photonMappingKernel.cu:213), and the
 * cudaError_t stubs used by checkCUDAError (photonMappingKernel.cu:49-55).
 */
#ifndef PMB200_ORACLE_HOST_CUDA_SHIM_H
#define PMB200_ORACLE_HOST_CUDA_SHIM_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define __device__
#define __global__
#define __host__

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct int3   { int x, y, z; };
struct uchar4 { unsigned char x, y, z, w; };
struct dim3   { unsigned int x, y, z; };

static inline float3 make_float3(float x, float y, float z) { float3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline float2 make_float2(float x, float y)          { float2 r; r.x = x; r.y = y; return r; }
static inline int3   make_int3(int x, int y, int z)         { int3 r; r.x = x; r.y = y; r.z = z; return r; }

extern thread_local dim3 blockIdx, blockDim, threadIdx;
static inline void __threadfence(void) {}

/* CUDA's device min/max: float pairs follow fminf/fmaxf; mixed pairs promote to double. */
static inline float  min(float a, float b)   { return fminf(a, b); }
static inline float  max(float a, float b)   { return fmaxf(a, b); }
static inline double min(double a, float b)  { return fmin(a, (double)b); }
static inline double min(float a, double b)  { return fmin((double)a, b); }
static inline double max(double a, float b)  { return fmax(a, (double)b); }
static inline double max(float a, double b)  { return fmax((double)a, b); }

typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline cudaError_t cudaGetLastError(void) { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "host build"; }

#endif
