#!/bin/bash
# Round-2 call V10 (1 GPU): mesh backward with two tiles per CTA and the second tile's indices prefetched by cp.async:
# timing, then parity with the variant library swapped in on the box.
set -u
mkdir -p gpurun_out
echo "== default"
timeout 200 python tools/phase_times.py --lib pytorch3d_b200/lib/libb200raster.so ns c2 ns_blur 2>&1 | tail -3
echo "== pair"
timeout 200 python tools/phase_times.py --lib tools/_variants/lib_pair.so ns c2 ns_blur 2>&1 | tail -3
cp tools/_variants/lib_pair.so pytorch3d_b200/lib/libb200raster.so
touch pytorch3d_b200/lib/libb200raster.so pytorch3d_b200/lib/_b200_ext*.so
timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "backward or full_size or autograd or indexed or config or peer or module" > gpurun_out/v10_pytest_pair.log 2>&1; echo "pytest(pair) rc=$?"; tail -3 gpurun_out/v10_pytest_pair.log
echo "== done"
