#!/bin/bash
# Round-2 GPU call X (1 GPU): setup-kernel occupancy variants.
set -u
mkdir -p gpurun_out
timeout 300 python tools/phase_times.py ns c2 c5 > gpurun_out/x_phase.log 2>&1; tail -3 gpurun_out/x_phase.log
for v in setup6 setup8; do
  timeout 300 python tools/phase_times.py --lib tools/_variants/lib_$v.so ns c2 c5 > gpurun_out/x_phase_$v.log 2>&1; tail -3 gpurun_out/x_phase_$v.log
done
echo "== done"
