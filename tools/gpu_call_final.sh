#!/bin/bash
# Round-2 final GPU call (1 GPU): whole GPU suite, smoke, full bench line, launch lists and ncu --set full captures of every
# workload (the files kept under profiles/ as *_r02_final_*), compute-sanitizer.
set -u
mkdir -p gpurun_out
OURS='regex:b200r|mesh_|tile_|points_'
echo "== pytest gpu"
timeout 1200 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/final_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench N=1"
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "rc=$?"; tail -c 300 gpurun_out/final_bench_n1.err
echo "== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "rc=$?"; head -c 400 gpurun_out/final_bench_ref.json
echo "== launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_ns.csv python bench.py --steps 2 --warmup 3 --skip-others --skip-cpu --skip-host-abi --skip-c4 > gpurun_out/final_launches_ns.log 2>&1
for w in ns_blur c2 c5 c3; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/final_launches_$w.csv python tools/profile_step.py $w 2 > /dev/null 2>&1
done
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 9 -c 5 -o gpurun_out/final_prof_ns -f python tools/profile_step.py ns 3 > gpurun_out/final_ncu_ns.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 9 -c 5 -o gpurun_out/final_prof_ns_blur -f python tools/profile_step.py ns_blur 3 > gpurun_out/final_ncu_ns_blur.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 9 -c 5 -o gpurun_out/final_prof_c2 -f python tools/profile_step.py c2 3 > gpurun_out/final_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 4 -c 5 -o gpurun_out/final_prof_c5 -f python tools/profile_step.py c5 2 > gpurun_out/final_ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 9 -c 5 -o gpurun_out/final_prof_c3 -f python tools/profile_step.py c3 3 > gpurun_out/final_ncu_c3.log 2>&1
ls -la gpurun_out/final_*.ncu-rep
echo "== sanitizer"
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/final_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 1 gpurun_out/final_memcheck.log
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_step.py > gpurun_out/final_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 1 gpurun_out/final_racecheck.log
echo "== done"
