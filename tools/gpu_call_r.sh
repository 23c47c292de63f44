#!/bin/bash
# Round-2 GPU call R (1 GPU): fused point rendering (parity + timing), full bench line.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/r_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r_pytest.log
echo "== bench"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "rc=$?"; tail -c 300 gpurun_out/r_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['modes']);
for k,v in d['other_workloads'].items(): print(k, v.get('ms_per_step'), v.get('frames_per_s'), (v.get('roofline') or {}).get('frac'))
print(d['reference_cuda']); print(d['cpu_baseline'])"
echo "== done"
