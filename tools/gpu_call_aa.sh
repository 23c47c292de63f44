#!/bin/bash
# Round-2 GPU call AA (1 GPU): indexed backward scattering straight into the vertices (parity + e2e), four-class schedule variant.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/aa_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/aa_pytest.log
echo "== phase: default (3 classes) vs order4"
timeout 300 python tools/phase_times.py c2 ns_blur c5 > gpurun_out/aa_phase.log 2>&1; tail -3 gpurun_out/aa_phase.log
timeout 300 python tools/phase_times.py --lib tools/_variants/lib_order4.so c2 ns_blur c5 > gpurun_out/aa_phase_order4.log 2>&1; tail -3 gpurun_out/aa_phase_order4.log
echo "== bench (e2e)"
timeout 900 python bench.py --steps 50 --warmup 5 --skip-others --skip-c4 --skip-cpu > gpurun_out/aa_bench.json 2> gpurun_out/aa_bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/aa_bench.json')); print(d['value'], d['e2e']['modes'], d['e2e_host_abi']['value'])"
echo "== sanitizer"
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/aa_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 1 gpurun_out/aa_memcheck.log
echo "== done"
