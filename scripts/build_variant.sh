#!/bin/bash
# scripts/build_variant.sh NAME [extra nvcc flags...]  ->  pgopt_b200/lib/libv_NAME.so  (A/B experiments, see ab_variants.py)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xcompiler -fPIC \
  -Xcompiler -ffp-contract=off -shared "$@" pgopt_b200/csrc/pgb_capi.cu -o pgopt_b200/lib/libv_$name.so
echo built libv_$name.so
