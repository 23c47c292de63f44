#!/bin/bash
# Round-2 GPU call Q (1 GPU): coalesced record write-out A/B, K = 16 on the north-star batch, parity.
set -u
mkdir -p gpurun_out
echo "== phase times"
timeout 600 python tools/phase_times.py ns ns_k16 c2 c3 > gpurun_out/q_phase.log 2>&1; tail -4 gpurun_out/q_phase.log
timeout 300 python tools/phase_times.py --lib tools/_variants/lib_directrec.so ns c2 > gpurun_out/q_phase_directrec.log 2>&1; tail -2 gpurun_out/q_phase_directrec.log
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/q_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/q_pytest.log
echo "== done"
