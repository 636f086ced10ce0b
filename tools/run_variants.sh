timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_generic.py -x -q -m gpu -k "not config4 and not config5" 2>&1 | tail -3
BPPP_W=20 BPPP_NSUB=2 BPPP_PROFILE=1 timeout 300 python tools/variant_bench.py
timeout 300 python tools/msm_once.py 21
