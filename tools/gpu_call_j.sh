#!/bin/bash
# Round-2 GPU call J (1 GPU): depth culling in the blur path (A/B against a build without it), ncu of config 3 and ns_blur.
set -u
mkdir -p gpurun_out
echo "== phase times"
timeout 600 python tools/phase_times.py ns c2 ns_blur c5 c3 > gpurun_out/j_phase.log 2>&1; tail -6 gpurun_out/j_phase.log
timeout 600 python tools/phase_times.py --lib tools/_variants/lib_nocull.so c2 ns_blur c5 > gpurun_out/j_phase_nocull.log 2>&1; tail -4 gpurun_out/j_phase_nocull.log
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/j_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/j_pytest.log
echo "== ncu"
OURS='regex:b200r|mesh_|tile_|points_'
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 9 -c 5 -o gpurun_out/j_prof_c3 -f python tools/profile_step.py c3 3 > gpurun_out/j_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mesh_fine -s 2 -c 1 -o gpurun_out/j_prof_ns_blur -f python tools/profile_step.py ns_blur 3 > gpurun_out/j_ncu_ns_blur.log 2>&1
ls -la gpurun_out/j_*.ncu-rep
echo "== done"
