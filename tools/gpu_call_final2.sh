#!/bin/bash
# Round-2 final GPU call of the last session (1 GPU): whole GPU suite, smoke, full bench line, launch lists and ncu --set full
# captures of every workload (kept under profiles/ as *_r02_final2_*).
set -u
mkdir -p gpurun_out
OURS='regex:b200r|mesh_|tile_|points_|zero_ints'
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/final2_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/final2_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench N=1"
date +%s
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/final2_bench_n1.json 2> gpurun_out/final2_bench_n1.err; echo "rc=$?"; tail -c 300 gpurun_out/final2_bench_n1.err
date +%s
python -c "
import json; d=json.load(open('gpurun_out/final2_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['modes'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['gpu_launches'])
for k,v in d['other_workloads'].items(): print(k, round(v['ms_per_step'],4), round(v['frames_per_s'],1))"
echo "== launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final2_launches_ns.csv python bench.py --steps 2 --warmup 3 --skip-others --skip-cpu --skip-host-abi --skip-c4 > gpurun_out/final2_launches_ns.log 2>&1
for w in ns_blur c2 c5 c3; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/final2_launches_$w.csv python tools/profile_step.py $w 2 > /dev/null 2>&1
done
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 12 -c 6 -o gpurun_out/final2_prof_ns -f python tools/profile_step.py ns 3 > gpurun_out/final2_ncu_ns.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 12 -c 6 -o gpurun_out/final2_prof_ns_blur -f python tools/profile_step.py ns_blur 3 > gpurun_out/final2_ncu_ns_blur.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 12 -c 6 -o gpurun_out/final2_prof_c3 -f python tools/profile_step.py c3 3 > gpurun_out/final2_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$OURS" -s 6 -c 6 -o gpurun_out/final2_prof_c5 -f python tools/profile_step.py c5 2 > gpurun_out/final2_ncu_c5.log 2>&1
ls -la gpurun_out/final2_*.ncu-rep
echo "== done"
