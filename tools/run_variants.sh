python -c "
import sys; sys.path.insert(0,'.')
import bp_pp_b200 as B, json; print(json.dumps(B.microbench(0)))"
BPPP_NSUB=2 BPPP_PROFILE=1 timeout 200 python tools/variant_bench.py
BPPP_W=20 BPPP_NSUB=2 BPPP_PROFILE=1 timeout 300 python tools/variant_bench.py
