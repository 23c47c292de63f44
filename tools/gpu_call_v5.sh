#!/bin/bash
# Round-2 call V5 (1 GPU): fewer load/store instructions in the backward kernels (faces as three 16-byte pieces, 8-byte vector
# reductions), point setup with all loads up front, run-length aggregation instead of MATCH: parity, A/B timings, sanitizer.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rs -x -p no:cacheprovider > gpurun_out/v5_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v5_pytest.log
echo "== default"
timeout 200 python tools/phase_times.py --lib pytorch3d_b200/lib/libb200raster.so ns c2 ns_blur ns_k16 c5 c3 2>&1 | tail -6
for v in scalar_face scalar_red gb64 rle; do
  echo "== $v"
  timeout 200 python tools/phase_times.py --lib tools/_variants/lib_$v.so ns c2 ns_blur ns_k16 2>&1 | tail -4
done
for v in scalar_red nohoist pfill_serial; do
  echo "== $v"
  timeout 200 python tools/phase_times.py --lib tools/_variants/lib_$v.so c3 2>&1 | tail -1
done
echo "== indexed"
timeout 200 python tools/time_indexed.py 2>&1 | tail -4
echo "== sanitizer"
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/v5_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 1 gpurun_out/v5_memcheck.log
echo "== done"
