#!/bin/bash
# A/B builds of the K0 v2 kernel (tuning knobs in csrc/k0_vit2.cu) into build_alt/: the rest of the library is reused from build/.
# usage: tools/build_k0_variants.sh name "-DK0V2_WARPS=4 -DK0V2_CTAS=4 ..."   ->  build_alt/librgbnm_<name>.so  (select with RGBNM_LIB=...)
set -e
cd "$(dirname "$0")/.."
mkdir -p build_alt
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC,-O3,-pthread --expt-relaxed-constexpr \
     -I include $@ -Xptxas -v -c rgb_no_more_b200/csrc/k0_vit2.cu -o build_alt/k0_vit2_$name.o 2>&1 | grep -E "ILi1|registers|spill" | sed -n 4,6p
objs=$(ls build/*.o | grep -v k0_vit2)
nvcc -shared -o build_alt/librgbnm_$name.so $objs build_alt/k0_vit2_$name.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-pthread -lpthread
echo "built build_alt/librgbnm_$name.so"
