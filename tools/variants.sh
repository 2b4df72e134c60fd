#!/bin/bash
# Build tuning variants of the engine into build/variants/ (measurement aid): tools/variants.sh name "-DX=.. -DY=.." ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -gt 1 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared $2 \
       -o build/variants/$1.so reina_b200/csrc/engine.cu &
  shift 2
done
wait
ls -la build/variants
