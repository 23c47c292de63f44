#!/bin/bash
# Round-2 call V3 (1 GPU): binning experiments (owner lookup per block, 128-face setup CTAs, fill unroll, points chunk size),
# ncu --set full of the binning kernels (meshes + points) with source lines.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rs -x -p no:cacheprovider > gpurun_out/v3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v3_pytest.log
echo "== default"
timeout 200 python tools/phase_times.py --lib pytorch3d_b200/lib/libb200raster.so ns c2 ns_blur c5 c3 2>&1 | tail -5
for v in noowner0 setup128 fill8; do
  echo "== $v"
  timeout 200 python tools/phase_times.py --lib tools/_variants/lib_$v.so ns c2 ns_blur c5 2>&1 | tail -4
done
for v in bin1024 bin512 bin256; do
  echo "== $v"
  timeout 200 python tools/phase_times.py --lib tools/_variants/lib_$v.so c3 2>&1 | tail -1
done
echo "== ncu binning kernels"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"setup|scan|fill" -s 3 -c 3 -o gpurun_out/v3_prof_bin_ns -f python tools/profile_step.py ns 3 > gpurun_out/v3_ncu_ns.log 2>&1; tail -1 gpurun_out/v3_ncu_ns.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"setup|scan|fill" -s 3 -c 3 -o gpurun_out/v3_prof_bin_c3 -f python tools/profile_step.py c3 3 > gpurun_out/v3_ncu_c3.log 2>&1; tail -1 gpurun_out/v3_ncu_c3.log
ls -la gpurun_out/*.ncu-rep
echo "== done"
