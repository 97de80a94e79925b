#!/bin/bash
# Developer tool: build a variant of libngp_b200.so with extra -D flags for A/B timing (only the field kernels are recompiled).
#   tools/ab_variant.sh <tag> [-DNAME=VALUE ...]   ->  raw_ngp_b200/lib/variants/libngp_b200_<tag>.so   (select with NGP_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
out=raw_ngp_b200/lib/variants; tmp=/tmp/ngp_ab_$tag
mkdir -p $out $tmp
python -m raw_ngp_b200.build >/dev/null
for f in field field_ws field_bwd_ws; do  # the kernels that include field_core.cuh
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -diag-suppress 177 -Xcompiler -fPIC "$@" \
       -c raw_ngp_b200/csrc/$f.cu -o $tmp/$f.o &
done
wait
objs=$(ls raw_ngp_b200/_build/*.o | grep -v -e /field.o -e /field_ws.o -e /field_bwd_ws.o)
nvcc -shared -o $out/libngp_b200_$tag.so $objs $tmp/field.o $tmp/field_ws.o $tmp/field_bwd_ws.o -gencode arch=compute_100a,code=sm_100a
echo "built $out/libngp_b200_$tag.so"
