#!/bin/bash
# Round-2 GPU call F (1 GPU): kernel variants of the fine pass, exchange kernels in isolation (+ ncu), parity.
set -u
mkdir -p gpurun_out
echo "== variants"
timeout 600 python tools/variant_time.py > gpurun_out/f_variants.log 2>&1; cat gpurun_out/f_variants.log | tail -8
echo "== exchange kernels"
timeout 300 python tools/time_exchange_kernels.py 2 > gpurun_out/f_exchange.log 2>&1; tail -2 gpurun_out/f_exchange.log
timeout 300 python tools/time_exchange_kernels.py 8 >> gpurun_out/f_exchange.log 2>&1; tail -1 gpurun_out/f_exchange.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fragments_" -s 8 -c 3 -o gpurun_out/f_prof_exchange -f python tools/time_exchange_kernels.py 2 2 > gpurun_out/f_ncu_exchange.log 2>&1
echo "== pytest gpu (peer + parity subset)"
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/f_pytest.log
echo "== done"
