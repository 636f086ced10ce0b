#!/bin/bash
# End-to-end (host-buffer) throughput against sub-batch count and stream priorities (run under gpurun on one B200).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/e2e_sweep.jsonl; : > $out
run() { echo "# $*" >> $out; env "$@" python bench.py --quick --no-cpu-baseline --metric prove --steps 3 --warmup 3 >> $out 2>> gpurun_out/e2e_sweep.err; }
run BPPP_STREAM_PRIO=0
run BPPP_STREAM_PRIO=1
run BPPP_STREAM_PRIO=1 BPPP_NSUB_HOST=2
run BPPP_STREAM_PRIO=0 BPPP_NSUB_HOST=2
run BPPP_STREAM_PRIO=1 BPPP_NSUB_HOST=8
run BPPP_STREAM_PRIO=1 BPPP_NSUB_HOST=3
python - <<'P'
import json
cfg=None
for l in open('gpurun_out/e2e_sweep.jsonl'):
    l=l.strip()
    if l.startswith('#'): cfg=l; continue
    if l.startswith('{'):
        d=json.loads(l)
        print(cfg, 'prove dev', round(d['value']), 'e2e', round(d['e2e']['value']), '| verify dev', round(d['verify']['value']), 'e2e', round(d['verify']['e2e']['value']), 'ok', d['outputs_ok'])
P
