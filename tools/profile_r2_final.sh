#!/bin/bash
# Final round-2 records (run under gpurun on one B200; outputs in gpurun_out/, summaries copied to profiles/ afterwards):
# the bench lines, the launch list of the bench command, and ncu --set full captures of the kernels that changed late in the
# round (k_v_tables_affine) and of the 4-lane ladder a strong-scaling rank runs at 8,192 proofs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
python bench.py --metric prove --quick --steps 3 --warmup 3 > gpurun_out/r2_bench_prove_1gpu.json 2> gpurun_out/r2_bench_prove_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --quick --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.json 2> gpurun_out/r2_bench_under_ncu.err
export BPPP_NSUB=1 BPPP_NSUB_HOST=1 BPPP_W=20
cap() { # name  kernel-regex  skip  batch
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/r2_$1 python tools/variant_bench.py $4 > gpurun_out/r2_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_$1.ncu-rep gpurun_out/r2_ncu_full_$1.txt
}
cap k_v_tables_affine_level3 'k_v_tables_affine' 8 65536
cap k_v_var2_lanes_8192 'k_v_var2_lanes' 9 8192
cap k_v_var2 '^k_v_var2$' 9 65536
ls -la gpurun_out/r2_*
