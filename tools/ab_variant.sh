#!/bin/bash
# Developer tool: build a variant of libngp_b200.so with extra -D flags for A/B timing.  Only the translation units named in
# $FILES (default: the kernels that include field_core.cuh) are recompiled, the rest comes from raw_ngp_b200/_build.
#   [FILES="optim raymarch"] tools/ab_variant.sh <tag> [-DNAME=VALUE ...]  ->  raw_ngp_b200/lib/variants/libngp_b200_<tag>.so
# Select it with NGP_B200_LIB=<path> (tools/ab_time.py, tools/ab_fwd.py).
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
files=${FILES:-"field field_ws field_bwd_ws"}
out=raw_ngp_b200/lib/variants; tmp=raw_ngp_b200/_build/ab_$tag
mkdir -p $out $tmp
python -m raw_ngp_b200.build >/dev/null
objs=""
for f in $files; do
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -diag-suppress 177 -Xcompiler -fPIC "$@" \
       -c raw_ngp_b200/csrc/$f.cu -o $tmp/$f.o &
  objs="$objs $tmp/$f.o"
done
wait
keep=$(for o in raw_ngp_b200/_build/*.o; do b=$(basename $o .o); case " $files " in *" $b "*) ;; *) echo $o;; esac; done)
nvcc -shared -o $out/libngp_b200_$tag.so $keep $objs -gencode arch=compute_100a,code=sm_100a
echo "built $out/libngp_b200_$tag.so"
