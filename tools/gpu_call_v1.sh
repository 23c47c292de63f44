#!/bin/bash
# Round-2 re-entry call V1 (1 GPU): GPU suite + smoke on the restored tree, phase baseline of every workload.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/v1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v1_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/phase_times.py ns c2 ns_blur ns_k16 c5 c3 > gpurun_out/v1_phase.log 2>&1; tail -8 gpurun_out/v1_phase.log
echo "== done"
