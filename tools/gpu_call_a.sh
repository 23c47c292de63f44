#!/bin/bash
# Round-2 GPU call A (1 GPU): parity suite, bench, launch lists, ncu captures of the changed kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.csv 2>&1
echo "== pytest gpu" 
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/a_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/a_smoke.log
echo "== bench"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/a_bench.err
echo "== launch lists"
for w in ns ns_blur c3 c5; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/a_launches_$w.csv python tools/profile_step.py $w 2 > gpurun_out/a_l_$w.log 2>&1
done
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mesh_fine_kernel|mesh_backward_kernel|mesh_setup|tile_scan|tile_fill" -s 5 -c 5 -o gpurun_out/a_prof_ns -f python tools/profile_step.py ns 3 > gpurun_out/a_ncu_ns.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"points_fine|points_backward" -s 2 -c 2 -o gpurun_out/a_prof_c3 -f python tools/profile_step.py c3 3 > gpurun_out/a_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mesh_fine_kernel" -s 1 -c 1 -o gpurun_out/a_prof_nsblur -f python tools/profile_step.py ns_blur 2 > gpurun_out/a_ncu_nsblur.log 2>&1
ls -la gpurun_out | tail -n 20
echo "== done"
