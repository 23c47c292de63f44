#!/bin/bash
# Round-2 GPU call Y (1 GPU): 16x8 forward tiles (128-thread CTAs) A/B + parity; setup occupancy and backward prefetch variants.
set -u
mkdir -p gpurun_out
echo "== default"
timeout 300 python tools/phase_times.py ns c2 ns_blur ns_k16 c5 > gpurun_out/y_phase.log 2>&1; tail -5 gpurun_out/y_phase.log
for v in tile8 setup6 setup8 bwdpf4; do
  echo "== $v"
  timeout 300 python tools/phase_times.py --lib tools/_variants/lib_$v.so ns c2 ns_blur ns_k16 c5 > gpurun_out/y_phase_$v.log 2>&1; tail -5 gpurun_out/y_phase_$v.log
done
echo "== pytest gpu (default build)"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/y_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/y_pytest.log
echo "== pytest gpu (16x8 tiles, ctypes binding)"
B200R_LIB=tools/_variants/lib_tile8.so timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/y_pytest_tile8.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/y_pytest_tile8.log
echo "== done"
