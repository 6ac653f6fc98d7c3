#!/bin/bash
# build a kernel variant: tools/build_variant.sh NAME "-DFOO=1 -DBAR=2"  ->  gpurun_variants/libnav24orb_NAME.so
set -e
cd "$(dirname "$0")/../nav24_b200/csrc"
out=../../variants; mkdir -p $out/_b_$1
for f in capi orb_kernels match_kernels; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $2 -c $f.cu -o $out/_b_$1/$f.o &
done; wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libnav24orb_$1.so $out/_b_$1/*.o
rm -rf $out/_b_$1
