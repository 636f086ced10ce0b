#!/bin/bash
# Builds kernel launch-configuration variants of libbppp.so for on-GPU tuning (tools/variant_bench.py).
set -e
cd "$(dirname "$0")/../bp_pp_b200/csrc"
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 550"
mkdir -p _obj/var ../variants
build() { # name  var_flags  core_flags
  name=$1
  ( $NV $2 -Xptxas -v -c -o _obj/var/${name}_var.o engine_var.cu 2> _obj/var/${name}_var.log ) &
  ( $NV $3 -Xptxas -v -c -o _obj/var/${name}_core.o engine_core.cu 2> _obj/var/${name}_core.log ) &
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libbppp_${name}.so _obj/var/${name}_var.o _obj/var/${name}_core.o _obj/engine_verify.o _obj/engine_prove.o _obj/engine_bench.o _obj/engine_msm.o _obj/engine_wnla.o _obj/engine_circuit.o
  echo built $name
  grep -h -A2 "k_v_var2\|k_msm_fixed" _obj/var/${name}_var.log _obj/var/${name}_core.log | grep -E "Used|spill" | paste - - | sed 's/ptxas info    ://g' | cut -c1-200
}
rm -f ../variants/*.so
build dblc   "-DBPPP_PTJ_DBL_NOINLINE"  "" &
wait
