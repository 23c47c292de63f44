#!/bin/bash
# Round-2 GPU call H (1 GPU): lane-independent mask walk + PDL chain: timing A/B, parity.
set -u
mkdir -p gpurun_out
echo "== phase times (pdl off / on)"
timeout 600 python tools/phase_times.py --pdl both ns c2 ns_blur c3 c5 > gpurun_out/h_phase.log 2>&1; cat gpurun_out/h_phase.log | tail -12
if [ -f tools/_variants/lib_oldwalk.so ]; then
  timeout 300 python tools/phase_times.py --lib tools/_variants/lib_oldwalk.so --pdl 1 ns c2 > gpurun_out/h_phase_oldwalk.log 2>&1; tail -3 gpurun_out/h_phase_oldwalk.log
fi
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/h_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/h_pytest.log
echo "== bench"
timeout 600 python bench.py --steps 50 --warmup 5 --skip-others --skip-cpu --skip-host-abi --skip-c4 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/h_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['modes'], d['roofline']['ms_per_launch'], d['roofline']['other_kernels'])"
echo "== done"
