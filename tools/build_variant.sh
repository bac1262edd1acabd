#!/bin/bash
# Development: builds a variant of the library with extra -D flags into build/variants/ (git-ignored,
# shipped by gpurun).  usage: tools/build_variant.sh name -DFLAG=... ; run with ACES_B200_LIB=build/variants/libaces_name.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants build/obj
# objects of the files the flags do not touch are shared between variants
for f in context dense design fit project; do
  if [ ! -f build/obj/$f.o ] || [ quantumaces.jl_b200/csrc/$f.cu -nt build/obj/$f.o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -c quantumaces.jl_b200/csrc/$f.cu -o build/obj/$f.o &
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c quantumaces.jl_b200/csrc/parity.cu -o build/obj/parity_$name.o &
wait
nvcc -shared -o build/variants/libaces_$name.so build/obj/context.o build/obj/dense.o build/obj/design.o build/obj/fit.o build/obj/project.o build/obj/parity_$name.o -lcudart
echo built build/variants/libaces_$name.so
