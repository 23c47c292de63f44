#!/bin/bash
# Round-2 call V6 (1 GPU): run-length aggregation in the binning passes and 16-byte reductions in the mesh backward as
# defaults: parity, A/B against the 8-byte-only / MATCH builds, sanitizer (memcheck + racecheck).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rs -x -p no:cacheprovider > gpurun_out/v6_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v6_pytest.log
echo "== default"
timeout 200 python tools/phase_times.py --lib pytorch3d_b200/lib/libb200raster.so ns c2 ns_blur ns_k16 c5 c3 2>&1 | tail -6
for v in v2only match; do
  echo "== $v"
  timeout 200 python tools/phase_times.py --lib tools/_variants/lib_$v.so ns c2 ns_blur ns_k16 c5 2>&1 | tail -5
done
echo "== indexed"
timeout 200 python tools/time_indexed.py 2>&1 | tail -4
echo "== sanitizer"
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/v6_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 1 gpurun_out/v6_memcheck.log
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_step.py > gpurun_out/v6_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 1 gpurun_out/v6_racecheck.log
echo "== done"
