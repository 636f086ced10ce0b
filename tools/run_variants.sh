BPPP_W=20 BPPP_NSUB=2 BPPP_PROFILE=1 timeout 300 python tools/variant_bench.py
