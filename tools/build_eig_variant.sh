#!/bin/bash
# Development: variant of the library with extra -D flags for eigen.cu -> build/variants/libaces_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants build/obj
for f in context dense design fit project parity; do
  if [ ! -f build/obj/$f.o ] || [ quantumaces.jl_b200/csrc/$f.cu -nt build/obj/$f.o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -c quantumaces.jl_b200/csrc/$f.cu -o build/obj/$f.o &
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c quantumaces.jl_b200/csrc/eigen.cu -o build/obj/eigen_$name.o &
wait
nvcc -shared -o build/variants/libaces_$name.so build/obj/context.o build/obj/dense.o build/obj/design.o build/obj/fit.o build/obj/project.o build/obj/parity.o build/obj/eigen_$name.o -lcudart
echo built build/variants/libaces_$name.so
