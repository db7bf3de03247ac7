#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.  Builds the reference's OWN hot file into oracle/_ref/:
#   oracle/_ref/libpmref_host.so   photonMappingKernel.cu:1-1521 as host C++ (sequential oracle + OpenMP CPU baseline)
#   oracle/_ref/libpmref_cuda_<N>.so  the whole photonMappingKernel.cu for sm_100a with nrPhotons = N (reference CUDA kernel speed baseline)
#
# The reference source is compiled from where it lies (/root/reference, read-only).  It is streamed
# through sed into a mktemp staging directory OUTSIDE the repository, compiled, and the staging
# directory is deleted: no reference source is ever written into this repository.  oracle/_ref/ is
# git-ignored but NOT gpurun-ignored, so the built .so files travel to the GPU box.
#
# Build-time patches applied to the stream (SURVEY.md 8(c) "harness trick"):
#   P1  '#define nrPhotons 10000'  ->  '#define nrPhotons PM_REF_CAPACITY'   (it is an in-file #define, -D cannot override it)
#   P2  host build only: the photon grid and the MWC state become thread_local (private per OpenMP thread)
#   P3  host build only: a record hook as the first statement of storePhoton / storeVolumePhoton
#   P4  CUDA "atomic" flavour only: the three racy grid `+=` become atomicAdd (see below)
# Nothing else is touched; the reference's own build system (none exists) is not used.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${PM_REFERENCE_DIR:-/root/reference}"
SRC="$REF/photonMappingKernel.cu"
OUT="$HERE/_ref"
CAP_HOST="${PM_REF_CAPACITY_HOST:-16777216}"
if [ ! -f "$SRC" ]; then
  echo "build_ref.sh: $SRC not present (GPU box?) -- keeping prebuilt oracle/_ref/" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT

# ---- host build -------------------------------------------------------------------------------
sed -n '1,1521p' "$SRC" \
 | sed -e 's/^#define nrPhotons 10000/#define nrPhotons PM_REF_CAPACITY/' \
       -e 's/__device__ float3 photons\[/thread_local float3 photons[/' \
       -e 's/^__device__ uint m_w = /thread_local uint m_w = /' \
       -e 's/^__device__ uint m_z = /thread_local uint m_z = /' \
       -e '/^__device__ void storePhoton(int type, int id, float3 location, float3 direction, float3 energy, int index){/a PM_HOOK_STORE(type, id, location, direction, energy, index);' \
       -e '/^__device__ void storeVolumePhoton(float3 location, float3 energy) {/a PM_HOOK_VOLUME(location, energy);' \
 > "$TMP/pmk_host.inc"
grep -q 'PM_REF_CAPACITY' "$TMP/pmk_host.inc"
[ "$(grep -c 'PM_HOOK_' "$TMP/pmk_host.inc")" = "2" ]
[ "$(grep -c 'thread_local' "$TMP/pmk_host.inc")" = "3" ]
g++ -O2 -fopenmp -ffp-contract=off -fPIC -shared -std=c++17 -w \
    -I"$HERE" -I"$HERE/shim" -DPM_REF_CAPACITY="$CAP_HOST" -DPM_REF_STAGED="\"$TMP/pmk_host.inc\"" \
    "$HERE/ref_host_harness.cpp" -o "$OUT/libpmref_host.so"
echo "built $OUT/libpmref_host.so"

# ---- CUDA builds (sm_100a): one library per photon count, because nrPhotons is compile-time in the reference ----
if [ -f "$HERE/ref_cuda_harness.cu" ] && command -v nvcc >/dev/null 2>&1; then
  sed -e 's/^#define nrPhotons 10000/#define nrPhotons PM_REF_CAPACITY/' "$SRC" > "$TMP/pmk_cuda.inc"
  grep -q 'PM_REF_CAPACITY' "$TMP/pmk_cuda.inc"
  for CAP in ${PM_REF_CAPACITIES_CUDA:-10000 1048576 16777216}; do
    nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -w -Xlinker -Bsymbolic \
         -I"$HERE" -I"$HERE/shim" -DPM_REF_CAPACITY="$CAP" -DPM_REF_STAGED="\"$TMP/pmk_cuda.inc\"" \
         "$HERE/ref_cuda_harness.cu" -o "$OUT/libpmref_cuda_$CAP.so" &
  done
  wait
  # ---- deterministic-deposit flavour: the reference kernel with its three racy `photons[..] += ..` sites (PMK:1068, :1158, :1177)
  #      turned into atomicAdd and compiled -fmad=false.  What is left between it and the product / the oracle is the DEVICE
  #      arithmetic of the reference -- rsqrtf in normalize (rsqrt.approx, 2 ulp) where the host build and the oracle use 1/sqrtf --
  #      and the order of the float sums.  tests/test_gpu_fullsize.py measures the image distance against this build (P4).
  sed -e 's/^#define nrPhotons 10000/#define nrPhotons PM_REF_CAPACITY/' \
      -e 's/photons\[i\]\[j\]\[k\] += \(SPLAT_ENERGY_WEIGHT\*energy\/dist\);/pm_atomic_add3(\&photons[i][j][k], \1);/' \
      -e 's/photons\[voxelPoint.x\]\[voxelPoint.y\]\[voxelPoint.z\] += energy;/pm_atomic_add3(\&photons[voxelPoint.x][voxelPoint.y][voxelPoint.z], energy);/' \
      "$SRC" > "$TMP/pmk_cuda_atomic.inc"
  [ "$(grep -c 'pm_atomic_add3' "$TMP/pmk_cuda_atomic.inc")" = "3" ]
  for CAP in ${PM_REF_CAPACITIES_CUDA_ATOMIC:-65536 1048576}; do
    nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared -w -Xlinker -Bsymbolic \
         -I"$HERE" -I"$HERE/shim" -DPM_REF_ATOMIC -DPM_REF_CAPACITY="$CAP" -DPM_REF_STAGED="\"$TMP/pmk_cuda_atomic.inc\"" \
         "$HERE/ref_cuda_harness.cu" -o "$OUT/libpmref_cuda_atomic_$CAP.so" &
  done
  wait
  ls "$OUT"/libpmref_cuda_*.so
fi
