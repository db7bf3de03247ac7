/* oracle/pm_oracle.h -- TEST INFRASTRUCTURE.  C API of the CPU restatement (oracle/pm_oracle.c) of the
 * reference's photon-mapping hot path.  See pm_oracle.c for the arithmetic contract and the citations. */
#ifndef PMB200_ORACLE_H
#define PMB200_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/pmb200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

uint32_t pmo_mwc_next(uint32_t *w, uint32_t *z);
float    pmo_rand_float(uint32_t *w, uint32_t *z, float max);
void     pmo_mwc_table(uint32_t *w, uint32_t *z, float *xyz, int n);
void     pmo_mwc_skip(uint32_t *w, uint32_t *z, long n);

void     pmo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void     pmo_philox_table(uint64_t seed, float *xyz, int n);

void pmo_scene_default(pm_scene *sc);
void pmo_position_objects(pm_scene *sc, float t);
void pmo_voxel(const float p[3], int v[3]);

/* Sequential emission of photon indices [n0,n1) (index order) into `grid` (32*32*32*3 floats, x-major, NOT
 * cleared; may be NULL) and/or `rec` (first max_rec store calls; may be NULL).  `table` is the whole random
 * table (entries 0..2 are also read by the medium scattering).  (w,z) is the MWC state the medium draws
 * continue from (the state left by pmo_mwc_table).  Returns the number of store calls. */
/* test aid: also sum the (FP32) deposit values in double into grid64[32*32*32*3] (NULL switches it off) */
void pmo_set_shadow_grid64(double *grid64);
long pmo_emit(const pm_scene *scene, float t, const float *table, int n0, int n1, int media,
              uint32_t *w, uint32_t *z, float *grid, pm_record *rec, long max_rec);

/* Rows [y0,y1) of the frame: float3 framebuffer `rgb` (may be NULL) and uchar4 `rgba` (may be NULL). */
void pmo_render(const pm_scene *scene, float t, const float *grid, int width, int height, int y0, int y1,
                int interp, int media, float *rgb, uint8_t *rgba);

void pmo_eye_geometry(const pm_scene *scene, float t, int width, int height, int y0, int y1, float *out_hit, float *out_march);

int  pmo_raytrace(const pm_scene *sc, const float ray[3], const float org[3], float *dist, int *type, int *idx);
void pmo_integrate_volume(const float *grid, const float p[3], float rgb[3]);
void pmo_gather(const float *grid, const float p[3], int type, int id, int interp, float rgb[3]);

#ifdef __cplusplus
}
#endif
#endif
