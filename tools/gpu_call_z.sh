#!/bin/bash
# Round-2 GPU call Z (1 GPU): tile schedule (heavy tiles first) A/B + parity + sanitizer.
set -u
mkdir -p gpurun_out
echo "== default (tile order on)"
timeout 300 python tools/phase_times.py ns c2 ns_blur ns_k16 c5 > gpurun_out/z_phase.log 2>&1; tail -5 gpurun_out/z_phase.log
echo "== notileorder"
timeout 300 python tools/phase_times.py --lib tools/_variants/lib_notileorder.so ns c2 ns_blur ns_k16 c5 > gpurun_out/z_phase_notileorder.log 2>&1; tail -5 gpurun_out/z_phase_notileorder.log
echo "== default through ctypes (same binding as the variant)"
B200R_LIB=pytorch3d_b200/lib/libb200raster.so timeout 300 python tools/phase_times.py ns c2 > gpurun_out/z_phase_ctypes.log 2>&1; tail -2 gpurun_out/z_phase_ctypes.log
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/z_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/z_pytest.log
echo "== sanitizer"
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/z_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 1 gpurun_out/z_memcheck.log
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_step.py > gpurun_out/z_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 1 gpurun_out/z_racecheck.log
echo "== done"
