#!/bin/bash
# Round-2 GPU call W (1 GPU): vector reductions in the compositing backward kernels: timing + parity.
set -u
mkdir -p gpurun_out
echo "== compositing pieces"
timeout 300 python tools/time_composite.py 2>&1 | tail -13
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/w_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/w_pytest.log
echo "== sanitizer"
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/w_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 1 gpurun_out/w_memcheck.log
echo "== done"
