#!/bin/bash
# Ladder-table construction sweep (run under gpurun on one B200): affine levels against the projective build, items per
# inversion by level, small batches.  One JSON line per configuration in gpurun_out/tab_sweep.jsonl.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/tab_sweep.jsonl; : > $out
run() { n=$1; shift; echo "# n=$n $*" >> $out; env "$@" BPPP_PROFILE=1 python tools/variant_bench.py $n >> $out 2>> gpurun_out/tab_sweep.err; }
run 65536 BPPP_TAB_AFFINE=1
run 65536 BPPP_TAB_K=8,16,32
run 65536 BPPP_TAB_K=4,8,16
run 65536 BPPP_TAB_K=16,16,32
run 65536 BPPP_TAB_K=32,64,64
run 8192 BPPP_TAB_AFFINE=1
run 2048 BPPP_TAB_AFFINE=1
cat $out
