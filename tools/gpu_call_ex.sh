#!/bin/bash
# Round-2 GPU call (1 GPU): the frame-exchange kernels in isolation (pack + push to local regions, expansion) + ncu capture.
set -u
mkdir -p gpurun_out
timeout 300 python tools/time_exchange_kernels.py 2 > gpurun_out/ex_times.log 2>&1; tail -1 gpurun_out/ex_times.log
timeout 300 python tools/time_exchange_kernels.py 8 >> gpurun_out/ex_times.log 2>&1; tail -1 gpurun_out/ex_times.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fragments_" -s 8 -c 3 -o gpurun_out/ex_prof -f python tools/time_exchange_kernels.py 2 2 > gpurun_out/ex_ncu.log 2>&1
ls -la gpurun_out/ex_prof.ncu-rep
echo "== done"
