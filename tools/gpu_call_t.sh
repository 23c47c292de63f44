#!/bin/bash
# Round-2 GPU call T (1 GPU): compositing pieces, backward prefetch at 4 CTAs/SM, parity after disabling the experiments.
set -u
mkdir -p gpurun_out
echo "== compositing pieces"
timeout 300 python tools/time_composite.py 2>&1 | tail -13
echo "== phase times"
timeout 600 python tools/phase_times.py ns c2 ns_blur > gpurun_out/t_phase.log 2>&1; tail -3 gpurun_out/t_phase.log
timeout 300 python tools/phase_times.py --lib tools/_variants/lib_bwdpf4.so ns c2 ns_blur > gpurun_out/t_phase_bwdpf4.log 2>&1; tail -3 gpurun_out/t_phase_bwdpf4.log
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/t_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_pytest.log
echo "== done"
