#!/bin/bash
# Runs tools/variant_bench.py for the default library and every library under bp_pp_b200/variants/.
cd "$(dirname "$0")/.."
export BPPP_W=${BPPP_W:-20} BPPP_PROFILE=1
timeout 300 python tools/variant_bench.py
for so in bp_pp_b200/variants/*.so; do
  BPPP_LIB=$PWD/$so timeout 300 python tools/variant_bench.py
done
