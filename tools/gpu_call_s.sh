#!/bin/bash
# Round-2 GPU call S (1 GPU): fused point rendering with point-major features; backward with up-front gradient loads (A/B).
set -u
mkdir -p gpurun_out
echo "== phase times (bwd prefetch: in-tree = 3 CTAs/SM; variants: 4 CTAs/SM, no prefetch)"
timeout 600 python tools/phase_times.py ns c2 ns_blur ns_k16 c5 > gpurun_out/s_phase.log 2>&1; tail -5 gpurun_out/s_phase.log
for v in bwdpf4 nobwdpf nodepth; do
  timeout 300 python tools/phase_times.py --lib tools/_variants/lib_$v.so ns c2 ns_blur c5 > gpurun_out/s_phase_$v.log 2>&1; tail -4 gpurun_out/s_phase_$v.log
done
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/s_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/s_pytest.log
echo "== bench others"
timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-host-abi --skip-c4 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s_bench.json')); print(d['value'], d['ms_per_step']);
for k,v in d['other_workloads'].items(): print(k, v.get('ms_per_step'), v.get('frames_per_s'))"
echo "== done"
