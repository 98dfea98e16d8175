#!/bin/bash
# Build libpnb200_<tag>.so with extra -D flags for the (16, 2) NNLS v3 kernel only (dev tool):
#   scripts/build_nnls_variant.sh <tag> "<flags>"       then run with PNB_LIB=pyneapple_b200/csrc/libpnb200_<tag>.so
set -e
cd "$(dirname "$0")/../pyneapple_b200/csrc"
tag=$1; flags=$2
mkdir -p _obj/var_$tag
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
  -Xptxas -v -DPNB_V3_MT=16 -DPNB_V3_WK=2 $flags -c pnb_nnls_v3_inst.cu -o _obj/var_$tag/nnlsv3_16_2.o 2> _obj/var_$tag/ptxas.log
grep -h "registers" _obj/var_$tag/ptxas.log | tail -1
objs=$(ls _obj/*.o | grep -v nnlsv3_16_2.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libpnb200_$tag.so $objs _obj/var_$tag/nnlsv3_16_2.o -lcudart
cuobjdump -xelf all _obj/var_$tag/nnlsv3_16_2.o > /dev/null; n=$(nvdisasm -c pnb_nnls_v3_inst.sm_100a.cubin | grep -c "^\s*/\*[0-9a-f]*\*/"); rm -f pnb_nnls_v3_inst.sm_100a.cubin
echo "variant $tag: $n instructions"
