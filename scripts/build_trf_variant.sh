#!/bin/bash
# Build libpnb200_<tag>.so with extra -D flags for ONE TRF kernel (MODEL=<id>, default 3 = bi-exponential S0) (dev tool):
#   scripts/build_trf_variant.sh <tag> "<flags>"       then run with PNB_LIB=pyneapple_b200/csrc/libpnb200_<tag>.so
set -e
cd "$(dirname "$0")/../pyneapple_b200/csrc"
tag=$1; flags=$2
mkdir -p _obj/var_$tag
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
  -Xptxas -v -DPNB_MODEL_ID=${MODEL:-3} -DPNB_T1MODE=0 $flags -c pnb_trf_inst.cu -o _obj/var_$tag/trf_${MODEL:-3}_0.o 2> _obj/var_$tag/ptxas.log
grep -h "registers\|spill" _obj/var_$tag/ptxas.log | head -2
objs=$(ls _obj/*.o | grep -v "/trf_${MODEL:-3}_0.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libpnb200_$tag.so $objs _obj/var_$tag/trf_${MODEL:-3}_0.o -lcudart
echo "built libpnb200_$tag.so"
