#!/bin/bash
# A/B builds of libdgprf.so that differ only in compile-time switches of one kernel file:
#   scripts/build_variants.sh k10_step_cluster.cu K10_VAR 0 1 2 3   ->  dgp-rf-mcmc_b200/lib/libdgprf_K10_VAR<v>.so
# (select at run time with DGPRF_LIB_PATH=<path>)
set -e
cd "$(dirname "$0")/../dgp-rf-mcmc_b200/csrc"
FILE=$1; MACRO=$2; shift 2
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC"
mkdir -p /tmp/dgprf_obj ../lib
for f in *.cu; do
  if [ "$f" != "$FILE" ] && [ ! -f /tmp/dgprf_obj/${f%.cu}.o -o "$f" -nt /tmp/dgprf_obj/${f%.cu}.o ]; then nvcc $FLAGS -c $f -o /tmp/dgprf_obj/${f%.cu}.o & fi
done
wait
for v in "$@"; do
  nvcc $FLAGS -D$MACRO=$v -c $FILE -o /tmp/dgprf_obj/var_$v.o &
done
wait
for v in "$@"; do
  OBJS=$(ls /tmp/dgprf_obj/*.o | grep -v "/var_" | grep -v "/${FILE%.cu}.o")
  nvcc -shared -o ../lib/libdgprf_${MACRO}${v}.so $OBJS /tmp/dgprf_obj/var_$v.o
  echo built ../lib/libdgprf_${MACRO}${v}.so
done
