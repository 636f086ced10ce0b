#!/bin/bash
# Round-2 profiles (run under gpurun on one B200; outputs land in gpurun_out/, summaries are copied to profiles/):
#   launch list of the bench command, and one `ncu --set full` capture per dominant kernel (full-batch launches: BPPP_NSUB=1).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --quick --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.json 2> gpurun_out/r2_bench_under_ncu.err
export BPPP_NSUB=1 BPPP_NSUB_HOST=1 BPPP_W=20
cap() { # name  kernel-regex  skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/r2_$1 python tools/variant_bench.py > gpurun_out/r2_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_$1.ncu-rep gpurun_out/r2_ncu_full_$1.txt
}
cap k_v_var2 '^k_v_var2$' 9
cap k_v_var5 '^k_v_var5$' 2
cap k_msm_fixed_verify49 k_msm_fixed 17
ncu --set full --clock-control none --import-source on -k "regex:k_msm_slices" -s 2 -c 1 -f -o gpurun_out/r2_k_msm_slices python tools/msm_once.py 21 > gpurun_out/r2_k_msm_slices.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_k_msm_slices.ncu-rep gpurun_out/r2_ncu_full_k_msm_slices.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_msm21.csv python tools/msm_once.py 21 > /dev/null 2>&1
ls -la gpurun_out/r2_*
