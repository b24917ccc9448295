#!/bin/bash
# tools/build_variant.sh NAME "<extra nvcc flags>" -> /tmp/veon_variants/NAME/libveonlift.so
# Experimental builds of the library (e.g. -DVEON_FWD_TRACE, -DVEON_FWD_NO_LOCKSTEP) for the
# tools/ scripts, which load them through VEON_LIB.  The shipped library is built by
# veon_b200/csrc/Makefile alone and carries none of these flags.
set -e
cd "$(dirname "$0")/.."
out=${VEON_VARIANT_DIR:-/tmp/veon_variants}/$1
mkdir -p $out
for f in veon_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --fmad=true $2 -c $f -o $out/$b.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libveonlift.so $out/*.o -lcudart
echo $out/libveonlift.so
