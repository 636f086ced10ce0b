#!/bin/bash
# Builds a complete variant of libbppp.so with extra nvcc flags: tools/build_variant_full.sh <name> "<flags>"
# -> bp_pp_b200/variants/libbppp_<name>.so (tools/run_variants.sh / tools/variant_bench.py time it on the GPU box).
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../bp_pp_b200/csrc"
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 550 $flags"
mkdir -p _obj/var/$name ../variants
for f in engine_core engine_verify engine_prove engine_var engine_var_lat engine_bench engine_msm engine_wnla engine_circuit engine_multi engine_peer; do
  ( $NV -Xptxas -v -c -o _obj/var/$name/$f.o $f.cu 2> _obj/var/$name/$f.log || (cat _obj/var/$name/$f.log; false) ) &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libbppp_$name.so _obj/var/$name/*.o
echo built $name
grep -h -A2 "k_v_var2N\|k_v_var5N\|k_msm_fixedILi4" _obj/var/$name/engine_var.log _obj/var/$name/engine_core.log | grep -E "Used|spill" | paste - - | sed 's/ptxas info    ://g' | cut -c1-220
