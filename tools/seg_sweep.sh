#!/bin/bash
# Segmented-ladder sweep (run under gpurun on one B200): BPPP_VAR_SEG = 1 (whole ladders) against 4 / 8 / 16 segments, warps per launch.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/seg_sweep.jsonl; : > $out
run() { n=$1; shift; echo "# n=$n $*" >> $out; env "$@" BPPP_PROFILE=1 timeout 150 python tools/variant_bench.py $n >> $out 2>> gpurun_out/seg_sweep.err; }
for cfg in "$@"; do run 65536 $cfg; done
python - <<'P'
import json
cfg=None
for l in open('gpurun_out/seg_sweep.jsonl'):
    l=l.strip()
    if l.startswith('#'): cfg=l; continue
    if l.startswith('{'):
        d=json.loads(l); kv=d.get('kernels_verify',{})
        print(cfg, '| verify', d['verify_ms'], 'prove', d['prove_ms'], 'ok', d['ok'], {k:v for k,v in kv.items() if 'var' in k})
P
