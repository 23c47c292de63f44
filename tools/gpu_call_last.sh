#!/bin/bash
# Round-2 last GPU call (1 GPU): GPU suite, smoke, full bench line and the --impl reference arm on the final tree.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/last_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/last_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/last_bench_n1.json 2> gpurun_out/last_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/last_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/last_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['modes'], d['roofline']['ms_per_launch'], d['roofline']['frac'])
for k,v in d['other_workloads'].items(): print(k, round(v['ms_per_step'],4), round(v['frames_per_s'],1))"
echo "== done"
