#!/bin/bash
# Round-2 call V2 (1 GPU): persistent double-buffered setup kernel, fill with four atomics in flight, points backward with
# the next slot's gather in flight: parity, then A/B timings against variant builds (tools/variant_time.py build ...).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rs -x -p no:cacheprovider > gpurun_out/v2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v2_pytest.log
echo "== default (ctypes binding, like the variants)"
timeout 200 python tools/phase_times.py --lib pytorch3d_b200/lib/libb200raster.so ns c2 ns_blur c3 2>&1 | tail -4
for v in oneshot fill1 setup5 ahead592 ahead1184 facepf facepf_ahead; do
  echo "== $v"
  timeout 200 python tools/phase_times.py --lib tools/_variants/lib_$v.so ns c2 ns_blur 2>&1 | tail -3
done
for v in bin1024 pbwd_nopipe; do
  echo "== $v"
  timeout 200 python tools/phase_times.py --lib tools/_variants/lib_$v.so c3 2>&1 | tail -1
done
echo "== indexed"
timeout 200 python tools/time_indexed.py 2>&1 | tail -4
echo "== done"
