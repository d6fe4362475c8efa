#!/bin/bash
# A/B builds of the library with extra -D flags (loaded through NGM_B200_LIB):  tools/build_variant.sh wa -DNGM_WARP_ARRIVE
set -e
tag=$1; shift
cd "$(dirname "$0")/../neural_graph_mapping_b200"
mkdir -p build_$tag
for f in abi sampler composite composite_bwd encode adam targets field_simt field_tc field_tc_bwd knn; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
       -I ../include -c csrc/$f.cu -o build_$tag/$f.o &
done
wait
nvcc -shared -o libngm_b200_$tag.so build_$tag/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
ls -la libngm_b200_$tag.so
