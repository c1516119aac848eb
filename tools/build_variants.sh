#!/bin/bash
# Build A/B variants of libgaisb200.so into gnuais_b200/lib/variants/ (git-ignored, travel to the GPU box):
#   tools/build_variants.sh name "EXTRA_NVFLAGS" [git-rev]     (git-rev: build that revision's sources instead of the tree)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; flags=$2; rev=$3
out=$ROOT/gnuais_b200/lib/variants; mkdir -p $out
src=$ROOT
if [ -n "$rev" ]; then
  src=$(mktemp -d); git -C $ROOT archive $rev gnuais_b200/csrc include | tar -x -C $src
fi
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$src/include -I$src/gnuais_b200/csrc $flags \
  -shared -o $out/$name.so $src/gnuais_b200/csrc/gais_api.cu $src/gnuais_b200/csrc/gais_synth.cu $src/gnuais_b200/csrc/gais_text.cpp
echo built $out/$name.so
