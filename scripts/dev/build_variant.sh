#!/bin/bash
# Dev tool: build a compile-time variant of the library.  Usage: scripts/dev/build_variant.sh NAME "-DFLAG ..."
# -> scripts/dev/_variants/lib_NAME.so (git-ignored; select it with RN_LIB_PATH)
set -e
cd "$(dirname "$0")/../.."
name=$1; flags=$2
out=scripts/dev/_variants; mkdir -p $out/obj_$name
for f in rec_now_b200/csrc/*.cu; do
  o=$out/obj_$name/$(basename ${f%.cu}).o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c $f -o $o &
done
wait
nvcc -shared -o $out/lib_$name.so $out/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf $out/obj_$name
echo $out/lib_$name.so
