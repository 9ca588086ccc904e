#!/bin/bash
# usage: tools/build_variant_file.sh FILE NAME [extra nvcc flags...]  -> variants/NAME.so with csrc/FILE.cu rebuilt with the flags
set -e
cd "$(dirname "$0")/.."
file=$1; name=$2; shift 2
mkdir -p variants /tmp/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC "$@" \
  -c fastc_b200/csrc/$file.cu -o /tmp/variants/${file}_$name.o
objs=""
for f in capi dxt etc1 bc7 decode pvrtc; do if [ $f != $file ]; then objs="$objs fastc_b200/csrc/$f.o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/$name.so $objs /tmp/variants/${file}_$name.o -lcudart
echo built variants/$name.so
