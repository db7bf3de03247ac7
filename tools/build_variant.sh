# tools/build_variant.sh NAME "extra nvcc flags" -- builds variants/libpmb200_NAME.so (A/B timing on the GPU box via PMB200_LIB)
set -e
cd "$(dirname "$0")/../cuda-photon-mapper_b200"
mkdir -p ../variants/build_$1
for f in csrc/*.cu; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-O2 $2 -c $f -o ../variants/build_$1/$(basename $f .cu).o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libpmb200_$1.so ../variants/build_$1/*.o -lcudart
echo built variants/libpmb200_$1.so
