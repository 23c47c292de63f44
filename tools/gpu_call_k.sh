#!/bin/bash
# Round-2 GPU call K (1 GPU): points scatter path, vectorised scan kernel, scan-conversion lanes-per-face variants.
set -u
mkdir -p gpurun_out
echo "== phase times"
timeout 600 python tools/phase_times.py ns c2 ns_blur c5 c3 > gpurun_out/k_phase.log 2>&1; tail -6 gpurun_out/k_phase.log
for v in scan1 scan2; do
  timeout 300 python tools/phase_times.py --lib tools/_variants/lib_$v.so ns > gpurun_out/k_phase_$v.log 2>&1; tail -1 gpurun_out/k_phase_$v.log
done
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/k_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/k_pytest.log
echo "== sanitizer"
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/k_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 1 gpurun_out/k_memcheck.log
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_step.py > gpurun_out/k_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 1 gpurun_out/k_racecheck.log
echo "== ncu"
OURS='regex:b200r|mesh_|tile_|points_'
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 9 -c 5 -o gpurun_out/k_prof_c3 -f python tools/profile_step.py c3 3 > gpurun_out/k_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 9 -c 5 -o gpurun_out/k_prof_ns -f python tools/profile_step.py ns 3 > gpurun_out/k_ncu_ns.log 2>&1
ls -la gpurun_out/k_*.ncu-rep
echo "== done"
