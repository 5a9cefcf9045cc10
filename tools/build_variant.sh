#!/bin/bash
# usage: tools/build_variant.sh NAME "<extra nvcc -D flags>" [file.cu ...]   (default file: spmm.cu)
# Recompiles the given csrc files with the extra flags and links them with the cached objects of the regular build into
# seoul_tourism_recommendation_ngcf_b200/variants/libngcf_NAME.so; select it at run time with NGCF_B200_LIB=<path>.
set -e
NAME=$1; FLAGS=$2; shift 2
FILES=${@:-spmm.cu}
PKG=$(dirname "$0")/../seoul_tourism_recommendation_ngcf_b200
mkdir -p $PKG/variants /tmp/variant_$NAME
OBJS=""
for f in $PKG/csrc/*.cu; do
  b=$(basename $f .cu)
  if echo " $FILES " | grep -q " $b.cu "; then
    nvcc -c -Xcompiler -fPIC -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a $FLAGS -I $PKG/../include -I $PKG/csrc $f -o /tmp/variant_$NAME/$b.o
    OBJS="$OBJS /tmp/variant_$NAME/$b.o"
  else
    OBJS="$OBJS $PKG/_build/$b.o"
  fi
done
nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a $OBJS -o $PKG/variants/libngcf_$NAME.so
echo $PKG/variants/libngcf_$NAME.so
