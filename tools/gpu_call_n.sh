#!/bin/bash
# Round-2 GPU call N (1 GPU): torch-extension binding: parity (whole suite), smoke, bench with e2e modes.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/n_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/n_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"
timeout 900 python bench.py --steps 50 --warmup 5 --skip-others --skip-c4 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "rc=$?"; tail -c 300 gpurun_out/n_bench.err; python -c "
import json; d=json.load(open('gpurun_out/n_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['modes'], d['e2e_host_abi']['value'], d['roofline']['ms_per_launch'], d['roofline']['other_kernels'])"
echo "== done"
