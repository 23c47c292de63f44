#!/bin/bash
# Round-2 GPU call L (1 GPU): fix of the empty-tile mask read; pre-rejection by approximate depth; parity + timing.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/l_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/l_pytest.log
echo "== phase times"
timeout 600 python tools/phase_times.py ns c2 ns_blur c5 c3 > gpurun_out/l_phase.log 2>&1; tail -6 gpurun_out/l_phase.log
echo "== sanitizer"
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/l_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 1 gpurun_out/l_memcheck.log
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_step.py > gpurun_out/l_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 1 gpurun_out/l_racecheck.log
timeout 700 compute-sanitizer --tool initcheck python tools/sanitize_step.py > gpurun_out/l_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -n 1 gpurun_out/l_initcheck.log
echo "== done"
