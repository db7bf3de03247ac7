// pm_kernels.cuh -- launcher declarations shared between the kernel TUs and the C-ABI host layer (pm_api.cu).
#pragma once

#include <cuda_runtime.h>
#include "../../include/pmb200.h"
#include "pm_layout.h"
#include "pm_math.cuh"

namespace pm {

// pm_trace.cu
cudaError_t launch_mwc_table(float *table, long long n, uint32_t w0, uint32_t z0, const MwcJump *J, cudaStream_t st);
cudaError_t launch_trace(const DeviceScene &sc, const float *table, long long first, long long last, unsigned flags,
                         uint32_t w0, uint32_t z0, const MwcJump *J, unsigned long long *acc, float4 *rec_pos,
                         float4 *rec_pow, float4 *rec_dir, unsigned long long *rec_count, long long rec_cap,
                         cudaStream_t st);

// pm_map.cu
cudaError_t launch_build_map(const long long *acc, float energy_scale, float *grid, cudaStream_t st);
cudaError_t launch_build_tables(const float *grid, float4 *vol_table, float4 *surf_table, cudaStream_t st);

// pm_render.cu
cudaError_t launch_render(const DeviceScene &sc, const float4 *vol_table, const float4 *surf_table, int width, int height,
                          int y0, int y1, bool interp, bool media, uchar4 *rgba, float4 *rgbf, cudaStream_t st);

}  // namespace pm
