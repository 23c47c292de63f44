#!/bin/bash
# Round-2 GPU call G (1 GPU): full GPU test suite, bench N=1, launch lists and ncu --set full captures of every
# workload's kernels (north-star, blur band, config 5, config 3), compute-sanitizer memcheck + racecheck.
set -u
mkdir -p gpurun_out
OURS='regex:b200r|mesh_|tile_|points_'
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/g_pytest.log
echo "== bench N=1"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/g_bench_n1.json 2> gpurun_out/g_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/g_bench_n1.err
echo "== launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g_launches_ns.csv python bench.py --steps 2 --warmup 3 --skip-others --skip-cpu --skip-host-abi --skip-c4 > gpurun_out/g_launches_ns.log 2>&1
for w in ns_blur c5 c3; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/g_launches_$w.csv python tools/profile_step.py $w 2 > gpurun_out/g_launches_$w.log 2>&1
done
echo "== ncu full: ns (one whole step), ns_blur, c5, c3"
# profile_step: one warm-up forward (4 of our kernels: setup, scan, fill, fine) then `steps` x 5 (... + backward)
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 9 -c 5 -o gpurun_out/g_prof_ns -f python tools/profile_step.py ns 3 > gpurun_out/g_ncu_ns.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 9 -c 5 -o gpurun_out/g_prof_ns_blur -f python tools/profile_step.py ns_blur 3 > gpurun_out/g_ncu_ns_blur.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 4 -c 5 -o gpurun_out/g_prof_c5 -f python tools/profile_step.py c5 2 > gpurun_out/g_ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 9 -c 5 -o gpurun_out/g_prof_c3 -f python tools/profile_step.py c3 3 > gpurun_out/g_ncu_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo "== sanitizer"
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/g_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 3 gpurun_out/g_memcheck.log
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_step.py > gpurun_out/g_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 3 gpurun_out/g_racecheck.log
echo "== done"
