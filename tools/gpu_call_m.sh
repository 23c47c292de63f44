#!/bin/bash
# Round-2 GPU call M (1 GPU): config 2 profile, timing after the latest changes, parity.
set -u
mkdir -p gpurun_out
echo "== phase times"
timeout 600 python tools/phase_times.py ns c2 ns_blur c5 c3 > gpurun_out/m_phase.log 2>&1; tail -6 gpurun_out/m_phase.log
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/m_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/m_pytest.log
echo "== ncu c2"
OURS='regex:b200r|mesh_|tile_|points_'
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 9 -c 5 -o gpurun_out/m_prof_c2 -f python tools/profile_step.py c2 3 > gpurun_out/m_ncu_c2.log 2>&1
ls -la gpurun_out/m_*.ncu-rep
echo "== done"
