for ns in 1 2 4 8; do BPPP_NSUB=$ns timeout 200 python tools/variant_bench.py; done
for v in v64x5_l2 v64x6_l4b64; do for ns in 1 4; do BPPP_LIB=$PWD/bp_pp_b200/variants/libbppp_$v.so BPPP_NSUB=$ns BPPP_PROFILE=1 timeout 200 python tools/variant_bench.py; done; done
