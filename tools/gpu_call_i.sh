#!/bin/bash
# Round-2 GPU call I (1 GPU): points -- depth-ordered walk, private-histogram binning: timing, parity, sanitizer.
set -u
mkdir -p gpurun_out
echo "== phase times"
timeout 600 python tools/phase_times.py c3 ns > gpurun_out/i_phase.log 2>&1; tail -4 gpurun_out/i_phase.log
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/i_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/i_pytest.log
echo "== sanitizer"
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/i_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 2 gpurun_out/i_memcheck.log
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_step.py > gpurun_out/i_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 2 gpurun_out/i_racecheck.log
echo "== launches c3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/i_launches_c3.csv python tools/profile_step.py c3 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/i_launches_c3.csv | tail -8
echo "== done"
