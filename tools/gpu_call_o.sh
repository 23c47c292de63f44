#!/bin/bash
# Round-2 GPU call O (1 GPU): binding overhead A/B, hash-accumulated points backward (timing + parity + sanitizer).
set -u
mkdir -p gpurun_out
echo "== binding overhead"
timeout 300 python tools/binding_overhead.py 2>&1 | tail -4
echo "== phase times"
timeout 600 python tools/phase_times.py c3 ns > gpurun_out/o_phase.log 2>&1; tail -3 gpurun_out/o_phase.log
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/o_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/o_pytest.log
echo "== sanitizer"
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/o_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 1 gpurun_out/o_memcheck.log
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_step.py > gpurun_out/o_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 1 gpurun_out/o_racecheck.log
echo "== bench e2e only"
timeout 900 python bench.py --steps 50 --warmup 5 --skip-others --skip-c4 --skip-cpu --skip-host-abi > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/o_bench.json')); print(d['value'], d['e2e']['modes'])"
echo "== done"
