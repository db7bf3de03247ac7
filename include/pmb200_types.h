/* include/pmb200_types.h -- plain-data types shared by the C-ABI (include/pmb200.h), the CUDA kernels
 * and the CPU oracle (oracle/pm_oracle.c).  No CUDA, no C++: C99 PODs only.
 *
 * The reference keeps all of this in __device__ globals and compile-time #defines
 * (photonMappingKernel.cu:9-46, :59-98); the legacy three-symbol ABI never passes them.  The extended
 * ABI passes them explicitly, with defaults (pm_scene_default) equal to the reference's globals.
 */
#ifndef PMB200_TYPES_H
#define PMB200_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* photon map dimensions: NR_PHOTONS_X/Y/Z (photonMappingKernel.cu:15-17), world box (:21-23) */
#define PM_GRID_N        32
#define PM_GRID_VOXELS   (PM_GRID_N * PM_GRID_N * PM_GRID_N)
#define PM_GRID_FLOATS   (PM_GRID_VOXELS * 3)
#define PM_MAX_SPHERES   3
#define PM_MAX_PLANES    5

/* Scene description == reference globals szImg, nrObjects, Light, spheres[], planes[]
 * (photonMappingKernel.cu:59-73) plus a camera pixel offset (0,0 == reference). */
typedef struct pm_scene {
  int32_t n_spheres;                /* nrObjects[0]: [0]=glass (refracts), [1]=mirror (reflects), [2]=unused */
  int32_t n_planes;                 /* nrObjects[1]: ids 0..4 = x=+1.5, y=-1.5, x=-1.5, y=+1.5, z=6 by default */
  float   spheres[PM_MAX_SPHERES][4];   /* centre xyz, radius */
  float   planes[PM_MAX_PLANES][2];     /* axis (0,1,2 stored as float, as in the reference), offset */
  float   light[3];                 /* point light, photon origin */
  int32_t sz_img;                   /* image-plane scale szImg: ray = (x/szImg-0.5, -(y/szImg-0.5), 1) */
  float   cam_ox, cam_oy;           /* added to the pixel coordinates before projection; 0 == reference */
  int32_t animate;                  /* 1 (reference behaviour): positionObjects(animTime) overwrites
                                       spheres[0..1].xyz on every launch (photonMappingKernel.cu:1380-1404);
                                       0: spheres are used as given */
} pm_scene;

/* One photon interaction, as recorded by the trace stage (Mode B input) and by the oracle's hooks on
 * storePhoton / storeVolumePhoton (photonMappingKernel.cu:1164, :1147).  52 bytes, packed ints first. */
typedef struct pm_record {
  int32_t type;        /* object type of the surface hit (0 sphere, 1 plane); -1 for a volume deposit */
  int32_t id;          /* object index;                                        -1 for a volume deposit */
  int32_t index;       /* photon index */
  int32_t kind;        /* 0 = storePhoton call (surface or shadow photon), 1 = storeVolumePhoton call */
  float   loc[3];
  float   dir[3];      /* incoming direction (zero for volume deposits) */
  float   energy[3];
} pm_record;

/* Marsaglia multiply-with-carry generator of the reference (photonMappingKernel.cu:1026-1037):
 * seeds m_w = 6548, m_z = 316. */
#define PM_MWC_SEED_W 6548u
#define PM_MWC_SEED_Z 316u

#ifdef __cplusplus
}
#endif
#endif
