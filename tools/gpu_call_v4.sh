#!/bin/bash
# Round-2 call V4 (1 GPU): points chunk 1024 as the default (parity), backward experiments: next-wave index prefetch without
# the 64-bit divisions of the first attempt, register budget for 5 / 6 CTAs per SM.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rs -x -p no:cacheprovider > gpurun_out/v4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v4_pytest.log
echo "== default"
timeout 200 python tools/phase_times.py --lib pytorch3d_b200/lib/libb200raster.so ns c2 ns_blur ns_k16 c3 2>&1 | tail -5
for v in aheadr aheadr2 aheadr1184 aheadr296 bwd5 bwd6 bwd5_aheadr; do
  echo "== $v"
  timeout 200 python tools/phase_times.py --lib tools/_variants/lib_$v.so ns c2 ns_blur ns_k16 2>&1 | tail -4
done
echo "== done"
