#!/bin/bash
# Round-2 call V7 (1 GPU): A/B of the setup pass with the owner looked up per block, the counters zeroed by a chained kernel
# instead of a memset node, and run-length merging in the mesh backward.
set -u
mkdir -p gpurun_out
for v in DEFAULT owner0 zerok zerok_owner0 bwd_rle DEFAULT; do
  echo "== $v"
  lib=tools/_variants/lib_$v.so; [ $v = DEFAULT ] && lib=pytorch3d_b200/lib/libb200raster.so
  timeout 200 python tools/phase_times.py --lib $lib ns c2 ns_blur ns_k16 2>&1 | tail -4
done
echo "== done"
