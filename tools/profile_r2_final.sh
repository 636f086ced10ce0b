#!/bin/bash
# Final round-2 records (run under gpurun on one B200; outputs in gpurun_out/, summaries copied to profiles/ afterwards):
# ncu --set full capture of the segmented two-point ladder, the bench lines, the launch list of the bench command.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export BPPP_W=20
ncu --set full --clock-control none --import-source on -k "regex:k_v_var_seg" -s 11 -c 1 -f -o gpurun_out/r2_k_v_var_seg timeout 300 python tools/variant_bench.py 65536 > gpurun_out/r2_k_v_var_seg.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_k_v_var_seg.ncu-rep gpurun_out/r2_ncu_full_k_v_var_seg.txt
unset BPPP_W
cp gpurun_out/r2_ncu_full_k_v_var_seg.txt profiles/ 2>/dev/null     # bench.py reads traffic / pipe occupancy of the dominant kernel from profiles/
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --quick --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.json 2> gpurun_out/r2_bench_under_ncu.err
grep -E "Kernel Name|Grid Size|gpu__time_duration.sum|warps_active.avg.per_cycle|fmaheavy_cycles_active.avg.pct|dram__bytes_read.sum \[" gpurun_out/r2_ncu_full_k_v_var_seg.txt | cut -c1-160
