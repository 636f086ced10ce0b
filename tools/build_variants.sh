#!/bin/bash
# Builds kernel launch-configuration variants of libbppp.so for on-GPU tuning (tools/variant_bench.py).
set -e
cd "$(dirname "$0")/../bp_pp_b200/csrc"
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 550"
mkdir -p _obj/var ../variants
build() { # name  var_flags  core_flags
  name=$1
  ( $NV $2 -c -o _obj/var/${name}_var.o engine_var.cu ) &
  ( $NV $3 -c -o _obj/var/${name}_core.o engine_core.cu ) &
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libbppp_${name}.so _obj/var/${name}_var.o _obj/var/${name}_core.o _obj/engine_verify.o _obj/engine_prove.o _obj/engine_bench.o
  echo built $name
}
rm -f ../variants/*.so
build v64x5_l2   "-DBPPP_VAR_BLOCK=64 -DBPPP_VAR_MINBLOCKS=5"  "-DBPPP_MSM_LANES=2" &
build v64x6_l4b64  "-DBPPP_VAR_BLOCK=64 -DBPPP_VAR_MINBLOCKS=6" "-DBPPP_MSM_LANES=4 -DBPPP_MSM_BLOCK=64 -DBPPP_MSM_MINBLOCKS=5" &
wait
