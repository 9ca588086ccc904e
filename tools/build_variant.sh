#!/bin/bash
# usage: tools/build_variant.sh NAME [extra nvcc flags...]  -> variants/NAME.so (bc7.cu rebuilt with the flags, other objects as built)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants /tmp/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC "$@" \
  -c fastc_b200/csrc/bc7.cu -o /tmp/variants/bc7_$name.o
objs=""
for f in capi dxt etc1 decode pvrtc; do objs="$objs fastc_b200/csrc/$f.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/$name.so $objs /tmp/variants/bc7_$name.o -lcudart
echo built variants/$name.so
