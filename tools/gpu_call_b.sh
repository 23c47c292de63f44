#!/bin/bash
# Round-2 GPU call B (1 GPU): parity suite after the optimistic-walk change, bench, launch lists, ncu, sanitizer.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 30 gpurun_out/b_pytest.log
echo "== bench"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/b_bench.err
echo "== launch lists"
for w in ns c3; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/b_launches_$w.csv python tools/profile_step.py $w 2 > gpurun_out/b_l_$w.log 2>&1
done
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mesh_fine_kernel|mesh_backward_kernel" -s 2 -c 2 -o gpurun_out/b_prof_ns -f python tools/profile_step.py ns 3 > gpurun_out/b_ncu_ns.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mesh_fine_smemq" -s 0 -c 1 -o gpurun_out/b_prof_c5 -f python tools/profile_step.py c5 1 > gpurun_out/b_ncu_c5.log 2>&1
echo "== sanitizer"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_step.py > gpurun_out/b_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 4 gpurun_out/b_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_step.py > gpurun_out/b_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 4 gpurun_out/b_racecheck.log
echo "== done"
